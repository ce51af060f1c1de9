#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02g.txt 2>&1
tail -8 gpurun_out/pytest_gpu_r02g.txt
timeout 300 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02g.txt 2>&1
cat gpurun_out/keyed_sweep_r02g.txt | cut -c1-150
timeout 200 python tools/api_probe.py > gpurun_out/api_probe_r02g.txt 2>&1
cat gpurun_out/api_probe_r02g.txt
timeout 400 python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/bench_c5_r02g.json 2> gpurun_out/bench_c5_r02g.err
head -c 300 gpurun_out/bench_c5_r02g.json; echo; tail -3 gpurun_out/bench_c5_r02g.err
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02g.json 2> gpurun_out/bench_c3_r02g.err
head -c 300 gpurun_out/bench_c3_r02g.json; echo; tail -3 gpurun_out/bench_c3_r02g.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_c5_r02g.csv \
    python bench.py --workload c5 --steps 6 --no-parity --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_c5_r02g.err
# K1 (shipped default): launch list of the default bench + one full-set capture
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches_c2_r02g.csv \
    python bench.py --steps 6 --no-parity --no-e2e --no-cpu-baseline --no-api --sustain-seconds 0 > /dev/null 2> gpurun_out/ncu_c2_r02g.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ingest_single_bulk -s 2 -c 1 -o gpurun_out/prof_k1_r02g \
    python bench.py --steps 3 --no-parity --no-e2e --no-cpu-baseline --no-api --sustain-seconds 0 > gpurun_out/ncu_k1_r02g.log 2>&1
tail -2 gpurun_out/ncu_k1_r02g.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/sanitize_${tool}_r02g.txt 2>&1
  tail -3 gpurun_out/sanitize_${tool}_r02g.txt
done
