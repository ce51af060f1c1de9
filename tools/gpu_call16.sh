#!/bin/bash
# 1 GPU, final: larger chunks for the record, full GPU tests and the keyed / mixed / default bench lines with the shipped defaults
mkdir -p gpurun_out
timeout 100 python tools/keyed_pf_probe.py kp_chunk=67108864,134217728,268435456 > gpurun_out/keyed_pf_probe_r02r.txt 2>&1
cat gpurun_out/keyed_pf_probe_r02r.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02r.txt 2>&1
tail -3 gpurun_out/pytest_gpu_r02r.txt
timeout 200 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02r.json 2> gpurun_out/bench_c3_r02r.err
head -c 250 gpurun_out/bench_c3_r02r.json; echo; tail -2 gpurun_out/bench_c3_r02r.err
timeout 200 python bench.py --workload c5 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5_r02r.json 2> gpurun_out/bench_c5_r02r.err
head -c 250 gpurun_out/bench_c5_r02r.json; echo; tail -2 gpurun_out/bench_c5_r02r.err
