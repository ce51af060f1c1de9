#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02f.txt 2>&1
tail -8 gpurun_out/pytest_gpu_r02f.txt
timeout 600 python tools/keyed_sweep.py 1000000000 1024 > gpurun_out/keyed_sweep_r02f.txt 2>&1
grep -E "vec|spt=4 " gpurun_out/keyed_sweep_r02f.txt | cut -c1-150
timeout 300 python tools/api_probe.py > gpurun_out/api_probe_r02f.txt 2>&1
cat gpurun_out/api_probe_r02f.txt
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline > gpurun_out/bench_c3_r02f.json 2> gpurun_out/bench_c3_r02f.err
head -c 300 gpurun_out/bench_c3_r02f.json; echo; tail -3 gpurun_out/bench_c3_r02f.err
timeout 400 python bench.py --workload c5 --steps 10 > gpurun_out/bench_c5_r02f.json 2> gpurun_out/bench_c5_r02f.err
head -c 300 gpurun_out/bench_c5_r02f.json; echo; tail -3 gpurun_out/bench_c5_r02f.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_c5_r02f.csv \
    python bench.py --workload c5 --steps 6 --no-parity --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_c5_r02f.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_wc -s 1 -c 1 -o gpurun_out/prof_kwc_r02f \
    python tools/keyed_sweep.py 500000000 1024 quick > gpurun_out/ncu_kwc_r02f.log 2>&1
tail -2 gpurun_out/ncu_kwc_r02f.log
