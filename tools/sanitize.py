"""Small end-to-end pass over every kernel, meant to run under compute-sanitizer (racecheck / memcheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

PS = [0.0, 0.5, 0.99, 1.0]
n = 300_001
with lh.Engine(device=0, max_histograms=64, max_counters=64) as e:
    d = e.gen_stream(lh.STREAM_S, n, lh.DEFAULT_SEED)
    ids = e.gen_ids_u16(0, n, 64, lh.DEFAULT_SEED)
    amt = e.gen_stream(lh.STREAM_AMOUNTS, n, lh.DEFAULT_SEED)
    ns = e.gen_stream(lh.STREAM_TIMER_NS, n, lh.DEFAULT_SEED)
    names = e.k1_variants()
    for vi in (0, 13, 21, 24, 25):
        e.tune("k1", vi)
        e.ingest_f64(1, d.offset(1), n - 1)
    e.ingest_keyed_f64_u16(ids, d, n)
    e.ingest_keyed_i64ns_u16(ids, ns, n)
    e.tune("keyed_mode", 2); e.tune("kp_chunk", 65536)
    for shape in (0, 1):
        e.tune("kp_shape", shape)
        e.ingest_keyed_f64_u16(ids, d, n)
    e.tune("keyed_mode", 0)
    e.counter_add_u16(ids, amt, n)
    e.snapshot_begin()
    h = e.snapshot_reduce_async(PS)
    sp = e.snapshot_export()
    e.snapshot_end()
    red = e.snapshot_result(h)
    assert int(red.counts.sum()) == 5 * (n - 1) + 4 * n, int(red.counts.sum())
    print("sanitize pass ok:", int(red.counts.sum()), "samples,", int(sp.offsets[-1]), "non-empty buckets")
