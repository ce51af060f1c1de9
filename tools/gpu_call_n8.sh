#!/bin/bash
# 8-GPU box: scaling of the sharded stream with the peer-memory collective (and NCCL for comparison)
mkdir -p gpurun_out
run() { # N collective tag extra...
  local N=$1 coll=$2 tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 20 --warmup 3 --collective $coll "$@" > gpurun_out/bench_n${N}_${tag}_r02.json 2> gpurun_out/bench_n${N}_${tag}_r02.err
  head -c 330 gpurun_out/bench_n${N}_${tag}_r02.json; echo; tail -2 gpurun_out/bench_n${N}_${tag}_r02.err
}
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py -q > gpurun_out/pytest_n8_r02.txt 2>&1
tail -5 gpurun_out/pytest_n8_r02.txt
run 8 peer peer --no-e2e
run 8 nccl nccl --no-e2e
run 4 peer peer --no-e2e
run 2 peer peer --no-e2e
run 8 peer peer_e2e --e2e-steps 3
