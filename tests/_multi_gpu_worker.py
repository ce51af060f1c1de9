"""torchrun worker for tests/test_gpu_multi.py and tests/test_gpu_comm.py: one rank per GPU; the bucket arrays are
summed by the library's peer-memory all-reduce (argv[1] == "peer") or by an NCCL all-reduce ("nccl", default);
argv[2] = number of histograms."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import loghisto_b200 as lh
from loghisto_b200.distributed import ShardedEngine, shard_range
from oracle import oracle as o


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total = 4_000_001
    ps = list(o.DEFAULT_PERCENTILES.values())
    a, b = shard_range(rank, world, n_total)
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 3      # >= 16 histograms: the all-reduce takes its two-shot (push) form
    with lh.Engine(device=local, max_histograms=H, max_counters=4) as eng:
        sh = ShardedEngine(eng, local, collective=sys.argv[1] if len(sys.argv) > 1 else "nccl")
        d = eng.gen_stream(lh.STREAM_S, b - a, lh.DEFAULT_SEED, start=a)
        eng.ingest_f64(1, d, b - a)
        ids = o.gen_ids(0, b - a, H, lh.DEFAULT_SEED, start=a).astype(np.uint16)
        d_ids = eng.upload(ids)
        eng.ingest_keyed_f64_u16(d_ids, d, b - a)
        eng.counter_add_u16_host(np.array([0, 3], np.uint16), np.array([rank + 1, 2 ** 63 + 1], np.uint64))
        for it in range(2):   # two intervals: the second must be empty on every rank
            red, sp = sh.snapshot(ps, export=True, counters=True)
            if it == 0:
                vals = o.gen_stream(o.STREAM_S, n_total, o.DEFAULT_SEED)
                all_ids = o.gen_ids(0, n_total, H, o.DEFAULT_SEED)
                want = o.ingest_keyed(all_ids, vals, H)
                want[1] += o.ingest(vals)
                for h in range(H):
                    got = np.zeros(65536, dtype=np.uint64)
                    for k, c in sp.histogram(h).items():
                        got[k & 0xFFFF] = c
                    assert (got == want[h]).all(), (rank, h)
                    ref = o.process_histogram(want[h], ps)
                    assert int(red.counts[h]) == ref["total"]
                    assert (red.pkeys[h] == ref["pkeys"]).all()
                assert int(sp.counter_deltas[0]) == sum(range(1, world + 1))
                assert int(sp.counter_deltas[3]) == (world * (2 ** 63 + 1)) % 2 ** 64
            else:
                assert int(red.counts.sum()) == 0 and int(sp.offsets[-1]) == 0
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK world=%d collective=%s" % (world, sh.collective))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
