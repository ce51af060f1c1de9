"""Small end-to-end pass over every kernel, meant to run under compute-sanitizer (racecheck / memcheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

PS = [0.0, 0.5, 0.99, 1.0]
n = 300_001
total = 0
with lh.Engine(device=0, max_histograms=64, max_counters=64) as e:
    d = e.gen_stream(lh.STREAM_S, n, lh.DEFAULT_SEED)
    ids = e.gen_ids_u16(0, n, 64, lh.DEFAULT_SEED)
    amt = e.gen_stream(lh.STREAM_AMOUNTS, n, lh.DEFAULT_SEED)
    ns = e.gen_stream(lh.STREAM_TIMER_NS, n, lh.DEFAULT_SEED)
    for vi, name in enumerate(e.k1_variants()):        # every K1 shape
        if name.startswith("probe"):
            continue
        e.tune("k1", vi)
        e.ingest_f64(1, d.offset(1), n - 1)
        total += n - 1
    e.tune("keyed_mode", 1)                             # L2-atomic kernel
    e.ingest_keyed_f64_u16(ids, d, n); total += n
    e.ingest_keyed_i64ns_u16(ids, ns, n); total += n
    e.tune("keyed_mode", 2); e.tune("kp_chunk", 65536)  # write-combining owner kernel, every tile shape, several chunks
    for spt in (6, 4, 3, 8):
        e.tune("wc_spt", spt)
        e.ingest_keyed_f64_u16(ids, d, n); total += n
    e.ingest_keyed_pair_u16(ids, d, n, ids, ns, n); total += 2 * n   # float64 + int64 segments in one launch
    e.tune("keyed_mode", 0)
    e.counter_add_u16(ids, amt, n)                      # vector + scalar counter kernels
    e.counter_add_u16(ids.offset(1), amt.offset(1), n - 1)
    e.snapshot_begin()
    h = e.snapshot_reduce_async(PS)
    sp = e.snapshot_export()
    e.snapshot_end()
    red = e.snapshot_result(h)
    assert int(red.counts.sum()) == total, (int(red.counts.sum()), total)
with lh.Engine(device=0, max_histograms=4, max_counters=4) as e:    # few histograms: shared-memory privatised keyed kernel
    d = e.gen_stream(lh.STREAM_S, n, lh.DEFAULT_SEED)
    ids = e.gen_ids_u16(0, n, 4, lh.DEFAULT_SEED)
    e.ingest_keyed_f64_u16(ids, d, n)
    red, sp2 = e.snapshot(PS)
    assert int(red.counts.sum()) == n
# two contexts on one device: the peer all-reduce kernel
engs = [lh.Engine(device=0, max_histograms=3, max_counters=2) for _ in range(2)]
handles = b"".join(x.comm_export() for x in engs)
for r, x in enumerate(engs):
    x.comm_import(r, 2, handles)
for r, x in enumerate(engs):
    dd = x.gen_stream(lh.STREAM_S, n, lh.DEFAULT_SEED, start=r * n)
    x.ingest_f64(1, dd, n)
for x in engs:
    x.snapshot_begin(); x.snapshot_allreduce()
for x in engs:
    red = x.snapshot_reduce(PS); x.snapshot_end()
    assert int(red.counts[1]) == 2 * n
for x in engs:
    x.close()
print("sanitize pass ok:", total, "samples,", int(sp.offsets[-1]), "non-empty buckets")
