#!/bin/bash
# 1 GPU, end-of-round state: smoke, full GPU test suite, the bench lines of every workload, launch list of c5, memcheck
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_r02m.txt 2>&1
tail -2 gpurun_out/smoke_r02m.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02m.txt 2>&1
tail -4 gpurun_out/pytest_gpu_r02m.txt
timeout 600 python bench.py > gpurun_out/bench_c2_r02m.json 2> gpurun_out/bench_c2_r02m.err
head -c 300 gpurun_out/bench_c2_r02m.json; echo; tail -2 gpurun_out/bench_c2_r02m.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02m.json 2> gpurun_out/bench_ref_r02m.err
head -c 300 gpurun_out/bench_ref_r02m.json; echo
timeout 300 python bench.py --workload c3 --steps 5 > gpurun_out/bench_c3_r02m.json 2> gpurun_out/bench_c3_r02m.err
head -c 300 gpurun_out/bench_c3_r02m.json; echo; tail -2 gpurun_out/bench_c3_r02m.err
timeout 400 python bench.py --workload c5 --steps 10 > gpurun_out/bench_c5_r02m.json 2> gpurun_out/bench_c5_r02m.err
head -c 300 gpurun_out/bench_c5_r02m.json; echo; tail -2 gpurun_out/bench_c5_r02m.err
timeout 300 python tools/keyed_batch_probe.py > gpurun_out/keyed_batch_probe_r02m.txt 2>&1
grep "shape 6" gpurun_out/keyed_batch_probe_r02m.txt | head -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_c5_r02m.csv \
    python bench.py --workload c5 --steps 6 --no-parity --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_c5_r02m.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitize_memcheck_r02m.txt 2>&1
tail -2 gpurun_out/sanitize_memcheck_r02m.txt
