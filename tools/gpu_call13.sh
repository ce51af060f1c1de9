#!/bin/bash
# 1 GPU: last check of the committed state -- smoke, full GPU test suite, default bench line
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_r02o.txt 2>&1
tail -1 gpurun_out/smoke_r02o.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02o.txt 2>&1
tail -3 gpurun_out/pytest_gpu_r02o.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_r02o.json 2> gpurun_out/bench_c2_r02o.err
head -c 250 gpurun_out/bench_c2_r02o.json; echo
