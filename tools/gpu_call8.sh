#!/bin/bash
# 1 GPU: full GPU test suite, keyed sweep over the flush interval, per-call API probe (threads x staging slot size),
# c3 / c5 bench, ncu capture of the write-combining kernel at the bench's batch size, racecheck
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02h.txt 2>&1
tail -6 gpurun_out/pytest_gpu_r02h.txt
timeout 400 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02j.txt 2>&1
cut -c1-170 gpurun_out/keyed_sweep_r02j.txt
PROBE_LOCK_AB=1 timeout 700 python tools/api_probe.py 1,16,32,64,128 4194304,262144 1 > gpurun_out/api_probe_r02h.txt 2>&1
cat gpurun_out/api_probe_r02h.txt
timeout 200 python tools/api_probe.py 32,128 262144 1024 >> gpurun_out/api_probe_r02h.txt 2>&1
tail -2 gpurun_out/api_probe_r02h.txt
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e --no-api > gpurun_out/bench_c3_r02h.json 2> gpurun_out/bench_c3_r02h.err
head -c 300 gpurun_out/bench_c3_r02h.json; echo; tail -3 gpurun_out/bench_c3_r02h.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02h \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-parity --no-e2e --no-cpu-baseline --no-api > gpurun_out/ncu_kwc_r02h.log 2>&1
tail -2 gpurun_out/ncu_kwc_r02h.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitize_racecheck_r02h.txt 2>&1
tail -3 gpurun_out/sanitize_racecheck_r02h.txt
