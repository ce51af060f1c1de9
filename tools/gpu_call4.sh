#!/bin/bash
# round-2 call 4: suite; keyed kernel with 16384-sample flush period; benches incl. per-call API leg; launch lists
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02d.txt 2>&1
tail -12 gpurun_out/pytest_gpu_r02d.txt
timeout 600 python tools/keyed_sweep.py 1000000000 1024 > gpurun_out/keyed_sweep_r02d.txt 2>&1
grep -E "vec|chunk=(8388608|16777216|33554432)" gpurun_out/keyed_sweep_r02d.txt | head -40
timeout 500 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_r02d.json 2> gpurun_out/bench_c2_r02d.err
head -c 300 gpurun_out/bench_c2_r02d.json; echo; tail -3 gpurun_out/bench_c2_r02d.err
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02d.json 2> gpurun_out/bench_c3_r02d.err
head -c 300 gpurun_out/bench_c3_r02d.json; echo; tail -3 gpurun_out/bench_c3_r02d.err
timeout 300 python bench.py --workload c5 --steps 10 > gpurun_out/bench_c5_r02d.json 2> gpurun_out/bench_c5_r02d.err
head -c 300 gpurun_out/bench_c5_r02d.json; echo; tail -3 gpurun_out/bench_c5_r02d.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_c5_r02d.csv \
    python bench.py --workload c5 --steps 6 --no-parity > /dev/null 2> gpurun_out/ncu_c5_r02d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02d \
    python tools/keyed_sweep.py 500000000 1024 quick > gpurun_out/ncu_kwc_r02d.log 2>&1
tail -2 gpurun_out/ncu_kwc_r02d.log
