#!/bin/bash
# One gpurun call of the development loop: tests, bench line, sweeps, ncu captures.  Outputs in gpurun_out/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$R.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench c2"; timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_c2_$R.err | tee gpurun_out/bench_c2_$R.json | cut -c1-1500
echo "== bench c3"; timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --cpu-seconds 6 2>gpurun_out/bench_c3_$R.err | tee gpurun_out/bench_c3_$R.json | cut -c1-1500
echo "== sweep"; timeout 900 python tools/k1_sweep.py --n 1000000000 --variants 0,2,5,9 --streams U,L,C,Z,S --iters 3 --keyed 1024 --out gpurun_out/k1_sweep_$R.jsonl 2>&1 | cut -c1-400
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2_$R.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2_$R.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3_$R.csv python bench.py --workload c3 --n 200000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3_$R.log 2>&1
echo "== ncu full K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ingest_single -s 3 -c 1 -f -o gpurun_out/prof_k1_$R python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_k1_$R.log 2>&1
echo "== ncu full keyed"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_vec -s 3 -c 1 -f -o gpurun_out/prof_keyed_$R python bench.py --workload c3 --n 200000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_keyed_$R.log 2>&1
ls -la gpurun_out | tail -20
