#!/bin/bash
# 2-GPU box: two-shot peer all-reduce (tests + c3 / c5 at N=2), keyed sweep of the current write-combining kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_multi.py -q > gpurun_out/pytest_n2_r02b.txt 2>&1
tail -8 gpurun_out/pytest_n2_r02b.txt
for wl in c3 c5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus 2 --steps 10 --warmup 3 --collective peer --workload $wl --no-e2e --no-api > gpurun_out/bench_n2_${wl}_peer_r02b.json 2> gpurun_out/bench_n2_${wl}_peer_r02b.err
  head -c 300 gpurun_out/bench_n2_${wl}_peer_r02b.json; echo; tail -3 gpurun_out/bench_n2_${wl}_peer_r02b.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
      bench.py --gpus 2 --steps 10 --warmup 3 --collective nccl --workload c3 --no-e2e --no-api > gpurun_out/bench_n2_c3_nccl_r02b.json 2> gpurun_out/bench_n2_c3_nccl_r02b.err
head -c 300 gpurun_out/bench_n2_c3_nccl_r02b.json; echo
timeout 300 python tools/keyed_sweep.py 1000000000 1024 quick > gpurun_out/keyed_sweep_r02i.txt 2>&1
cut -c1-150 gpurun_out/keyed_sweep_r02i.txt
