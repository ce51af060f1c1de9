#!/bin/bash
# 8-GPU box: scaling of the sharded stream with the peer-memory collective (and NCCL for comparison), keyed at N=8
mkdir -p gpurun_out
run() { # N collective tag extra...
  local N=$1 coll=$2 tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --warmup 3 --collective $coll --no-api "$@" > gpurun_out/bench_n${N}_${tag}_r02.json 2> gpurun_out/bench_n${N}_${tag}_r02.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${tag}_r02.json').read().strip().splitlines()[-1])
    print('N=${N} ${tag}', round(d['value']/1e9,1), 'G/s  ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'allreduce_ms', d.get('collective',{}).get('allreduce_ms'), 'parity', d.get('parity',{}).get('ok'), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print('N=${N} ${tag} FAILED', e); print(open('gpurun_out/bench_n${N}_${tag}_r02.err').read()[-1500:])
PY
}
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py -q > gpurun_out/pytest_n8_r02.txt 2>&1
tail -3 gpurun_out/pytest_n8_r02.txt
run 8 peer peer --steps 20 --no-e2e
run 8 nccl nccl --steps 20 --no-e2e
run 4 peer peer --steps 20 --no-e2e
run 2 peer peer --steps 20 --no-e2e
run 1 peer single --steps 20 --no-e2e --no-cpu-baseline
run 8 peer c3_peer --steps 10 --no-e2e --workload c3
run 8 peer peer_e2e --steps 5 --e2e-steps 3
