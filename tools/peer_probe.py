"""Peer-memory all-reduce in isolation (torchrun, one rank per GPU): all H histograms touched, ranks aligned by a barrier
right before every snapshot, time of the all-reduce kernel alone over the number of SMs it may use."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import loghisto_b200 as lh
from loghisto_b200.distributed import ShardedEngine

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for H in (1, 1024):
    with lh.Engine(device=local, max_histograms=H, max_counters=4) as eng:
        sh = ShardedEngine(eng, local, collective="peer")
        n = 4_000_000
        d = eng.gen_stream(lh.STREAM_U, n, lh.DEFAULT_SEED, start=rank * n)
        ids = eng.gen_ids_u16(0, n, H, lh.DEFAULT_SEED)
        for reserve in ((1,) if H == 1 else (1, 4, 16, 64)):
            eng.tune("k1_reserve_sms", reserve)
            ms = []
            for it in range(12):
                if H == 1:
                    eng.ingest_f64(0, d, n)
                else:
                    eng.ingest_keyed_f64_u16(ids, d, n)
                eng.sync()
                dist.barrier()
                torch.cuda.synchronize()
                eng.snapshot_begin()
                seq = eng.snapshot_allreduce()
                red = eng.snapshot_reduce([0.5])
                eng.snapshot_end()
                ms.append(eng.comm_allreduce_ms(seq))
                assert int(red.counts.sum()) == world * n
            ms = sorted(ms[2:])
            if rank == 0:
                print("world %d H=%4d SMs for the all-reduce %2d: kernel %.3f ms median (min %.3f), %d bytes from peers"
                      % (world, H, reserve, ms[len(ms) // 2], ms[0], eng.comm_last_bytes()), flush=True)
dist.barrier()
dist.destroy_process_group()
