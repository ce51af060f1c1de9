#!/bin/bash
# round-2 call 2: the refactored library (precision, flags, peer comm, write-combining keyed kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02b.txt 2>&1
tail -5 gpurun_out/pytest_gpu_r02b.txt
timeout 900 python tools/keyed_sweep.py 1000000000 1024 > gpurun_out/keyed_sweep_r02b.txt 2>&1
tail -30 gpurun_out/keyed_sweep_r02b.txt
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_r02b.json 2> gpurun_out/bench_c2_r02b.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --stream S --sustain-seconds 0 > gpurun_out/bench_c2S_r02b.json 2> gpurun_out/bench_c2S_r02b.err
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02b.json 2> gpurun_out/bench_c3_r02b.err
timeout 300 python bench.py --workload c5 --steps 10 > gpurun_out/bench_c5_r02b.json 2> gpurun_out/bench_c5_r02b.err
for f in c2 c2S c3 c5; do head -c 300 gpurun_out/bench_${f}_r02b.json; echo; tail -3 gpurun_out/bench_${f}_r02b.err; done
# ncu: the new keyed kernel, full set, one launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02b \
    python tools/keyed_sweep.py 500000000 1024 quick > gpurun_out/ncu_kwc_r02b.log 2>&1
tail -3 gpurun_out/ncu_kwc_r02b.log
