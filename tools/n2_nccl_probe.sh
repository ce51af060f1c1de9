run() { echo "== $*"; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --no-e2e --reserve-sms ${RES:-2} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['allreduce_ms'],3))"; }
run NCCL_CGA_CLUSTER_SIZE=0
run NCCL_CGA_CLUSTER_SIZE=1 NCCL_MAX_NCHANNELS=1
RES=4 run NCCL_CGA_CLUSTER_SIZE=0 NCCL_MAX_NCHANNELS=4
RES=8 run NCCL_MAX_NCHANNELS=2
run NCCL_CGA_CLUSTER_SIZE=0 NCCL_PROTO=LL NCCL_ALGO=Ring
