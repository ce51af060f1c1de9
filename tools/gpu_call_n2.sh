#!/bin/bash
# 2-GPU box: peer-memory collective behind the C ABI (IPC path), NCCL path, C host; bench at N=2 with both collectives
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py -q > gpurun_out/pytest_n2_r02.txt 2>&1
tail -8 gpurun_out/pytest_n2_r02.txt
for coll in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 3 --collective $coll --no-e2e > gpurun_out/bench_n2_${coll}_r02.json 2> gpurun_out/bench_n2_${coll}_r02.err
  head -c 400 gpurun_out/bench_n2_${coll}_r02.json; echo; tail -3 gpurun_out/bench_n2_${coll}_r02.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus 2 --steps 10 --warmup 3 --collective peer --workload c3 --no-e2e > gpurun_out/bench_n2_c3_peer_r02.json 2> gpurun_out/bench_n2_c3_peer_r02.err
head -c 400 gpurun_out/bench_n2_c3_peer_r02.json; echo; tail -3 gpurun_out/bench_n2_c3_peer_r02.err
timeout 300 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02h.txt 2>&1
cat gpurun_out/keyed_sweep_r02h.txt | cut -c1-150
