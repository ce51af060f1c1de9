#!/bin/bash
# 2-GPU box: two-shot all-reduce sized to the reserved SMs (tests, c3 / c5 / small-batch c3 at N=2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py -q > gpurun_out/pytest_n2_r02f.txt 2>&1
tail -4 gpurun_out/pytest_n2_r02f.txt
run() { # tag args...
  local tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus 2 --warmup 3 --collective peer --no-e2e --no-api "$@" > gpurun_out/bench_n2_${tag}_r02f.json 2> gpurun_out/bench_n2_${tag}_r02f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_${tag}_r02f.json').read().strip().splitlines()[-1])
print('${tag}', round(d['value']/1e9,1), 'G/s  ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'allreduce_ms', d['collective']['allreduce_ms'], 'parity', d['parity']['ok'])
PY
}
run c3 --workload c3 --steps 10
run c5 --workload c5 --steps 10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/peer_probe.py > gpurun_out/peer_probe_n2_r02f.txt 2>&1
grep "^world" gpurun_out/peer_probe_n2_r02f.txt
run default --steps 20
