#!/bin/bash
# 1 GPU: the 256 M-sample chunk default -- full GPU tests and the keyed bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02s.txt 2>&1
tail -3 gpurun_out/pytest_gpu_r02s.txt
timeout 200 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02s.json 2> gpurun_out/bench_c3_r02s.err
head -c 250 gpurun_out/bench_c3_r02s.json; echo; tail -2 gpurun_out/bench_c3_r02s.err
