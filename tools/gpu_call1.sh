#!/bin/bash
# round-2 call 1: primitive costs, full-size parity bench lines, the existing GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc >> gpurun_out/c1_smi.txt; free -g >> gpurun_out/c1_smi.txt
timeout 300 tools/ubench > gpurun_out/ubench_r02.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_c2_r02a.json 2> gpurun_out/bench_c2_r02a.err
timeout 600 python bench.py --workload c4x1 --steps 10 --no-cpu-baseline > gpurun_out/bench_c4x1_r02a.json 2> gpurun_out/bench_c4x1_r02a.err
timeout 600 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02a.json 2> gpurun_out/bench_c3_r02a.err
timeout 600 python bench.py --workload c5 --steps 10 > gpurun_out/bench_c5_r02a.json 2> gpurun_out/bench_c5_r02a.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.txt 2>&1
tail -3 gpurun_out/pytest_gpu_r02a.txt
head -c 600 gpurun_out/bench_c2_r02a.json
