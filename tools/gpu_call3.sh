#!/bin/bash
# round-2 call 3: full GPU suite on the reworked library, keyed kernel after the latency fixes, K1 variants sustained
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02c.txt 2>&1
tail -15 gpurun_out/pytest_gpu_r02c.txt
timeout 600 python tools/keyed_sweep.py 1000000000 1024 > gpurun_out/keyed_sweep_r02c.txt 2>&1
grep -E "chunk=(8388608|16777216|33554432)" gpurun_out/keyed_sweep_r02c.txt | head -40
timeout 600 python tools/k1_sustained.py 1000000000 1.5 0,1,5,6,3,4 U,N > gpurun_out/k1_sustained_r02c.txt 2>&1
cat gpurun_out/k1_sustained_r02c.txt
timeout 400 python bench.py --no-cpu-baseline --no-e2e --debug-steps > gpurun_out/bench_c2_r02c.json 2> gpurun_out/bench_c2_r02c.err
head -c 400 gpurun_out/bench_c2_r02c.json; echo; cat gpurun_out/bench_c2_r02c.err | tail -5
timeout 300 python bench.py --no-cpu-baseline --no-e2e --debug-steps --no-api --k1-variant 5 > gpurun_out/bench_c2v5_r02c.json 2> gpurun_out/bench_c2v5_r02c.err
head -c 400 gpurun_out/bench_c2v5_r02c.json; echo; cat gpurun_out/bench_c2v5_r02c.err | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02c \
    python tools/keyed_sweep.py 500000000 1024 quick > gpurun_out/ncu_kwc_r02c.log 2>&1
tail -2 gpurun_out/ncu_kwc_r02c.log
