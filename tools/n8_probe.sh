#!/bin/bash
run() { echo "== $*"; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['allreduce_ms'],3), d['count_ok'], d['pipeline'])"; }
run --pipeline-depth 0
run --pipeline-depth 0 --nccl-defaults
run --pipeline-depth 2 --nccl-defaults --reserve-sms 8
run --pipeline-depth 1 --reserve-sms 4
