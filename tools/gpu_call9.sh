#!/bin/bash
# 1 GPU: keyed sweep over L2-prefetch distance / chunk / flush / tile shape, c5 with the fused Histogram+Timer launch, c3, default bench
mkdir -p gpurun_out
timeout 600 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02k.txt 2>&1
cut -c1-175 gpurun_out/keyed_sweep_r02k.txt
timeout 400 python bench.py --workload c5 --steps 10 --no-cpu-baseline > gpurun_out/bench_c5_r02k.json 2> gpurun_out/bench_c5_r02k.err
head -c 330 gpurun_out/bench_c5_r02k.json; echo; tail -3 gpurun_out/bench_c5_r02k.err
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_r02k.json 2> gpurun_out/bench_c3_r02k.err
head -c 330 gpurun_out/bench_c3_r02k.json; echo; tail -3 gpurun_out/bench_c3_r02k.err
timeout 600 python bench.py > gpurun_out/bench_c2_r02k.json 2> gpurun_out/bench_c2_r02k.err
head -c 330 gpurun_out/bench_c2_r02k.json; echo; tail -3 gpurun_out/bench_c2_r02k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2_r02k.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('api_e2e'))[:1500])
print(json.dumps(d.get('sustained'))[:400], d.get('value_sustained'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_c5_r02k.csv \
    python bench.py --workload c5 --steps 6 --no-parity --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_c5_r02k.err
