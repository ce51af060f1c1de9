#!/bin/bash
# ncu evidence for the shipped default K1 (bulk2) and for the partitioned keyed kernel
set -u
R=${1:-r01b}
mkdir -p gpurun_out
echo "== ncu launches c2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2_$R.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2_$R.log 2>&1
echo "== ncu full K1 default"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_single -s 3 -c 1 -f -o gpurun_out/prof_k1_$R python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_k1_$R.log 2>&1
echo "== ncu full keyed_part"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_keyed_part -s 1 -c 1 -f -o gpurun_out/prof_kpart_$R python tools/keyed_part_probe.py 100000000 0 > gpurun_out/ncu_full_kpart_$R.log 2>&1
tail -3 gpurun_out/ncu_full_kpart_$R.log
ls -la gpurun_out | tail -8
