"""N > 1 path on CPU: two gloo ranks shard the stream, all-reduce their bucket arrays, and must end with
exactly the single-process histogram and percentiles."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from loghisto_b200.distributed import allreduce_sum_u64, shard_range
    from oracle import oracle as o
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(rank, world, n_total)
    vals = o.gen_stream(o.STREAM_S, b - a, o.DEFAULT_SEED, start=a)     # index-addressable generator
    counts = o.ingest(vals)
    counters = np.array([rank + 1, 2 ** 63 + 5, 2 ** 63 + 7], dtype=np.uint64)   # wraps mod 2^64 across ranks
    t = torch.from_numpy(counts.view(np.int64))
    allreduce_sum_u64(t)
    c = torch.from_numpy(counters.view(np.int64))
    allreduce_sum_u64(c)
    q.put((rank, t.numpy().view(np.uint64).copy(), c.numpy().view(np.uint64).copy()))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from loghisto_b200.distributed import shard_range
    for n, w in ((10, 3), (1_000_000_007, 8), (5, 8), (0, 2), (10 ** 10, 8)):
        rs = [shard_range(r, w, n) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(3, 8, 10 ** 10) == (3 * 1_250_000_000, 4 * 1_250_000_000)


@pytest.mark.timeout(120)
def test_two_rank_allreduce_equals_single_process(oracle):
    import torch.multiprocessing as mp
    n_total, world = 400_001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = oracle.ingest(oracle.gen_stream(oracle.STREAM_S, n_total, oracle.DEFAULT_SEED))
    ps = list(oracle.DEFAULT_PERCENTILES.values())
    ref = oracle.process_histogram(want, ps)
    for rank, counts, counters in results:
        assert (counts == want).all(), rank
        got = oracle.process_histogram(counts, ps)
        assert (got["pkeys"] == ref["pkeys"]).all() and got["total"] == n_total
        # uint64 sums wrap mod 2^64 exactly like atomic.AddUint64: 2*(2^63+5) = 10, 2*(2^63+7) = 14
        assert counters.tolist() == [3, 10, 14]
