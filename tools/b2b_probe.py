"""Is K1 slower back-to-back?  Launch K1 R times without host syncs, with/without concurrent snapshots."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

n = 1_000_000_000
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
d = eng.gen_stream(0, n, lh.DEFAULT_SEED)
eng.sync()
PS = [0.5, 0.99]

def show(tag, seqs):
    eng.sync()
    print(tag, " ".join("%.3f" % eng.kernel_ms(s) for s in seqs), flush=True)

for gm in (1, 2):
    eng.tune("k1_grid_mult", gm)
    # (a) isolated: sync between launches
    seqs = []
    for _ in range(6):
        eng.ingest_f64(0, d, n); seqs.append(eng.ingest_seq()); eng.sync(); time.sleep(0.002)
    show("gm=%d isolated      " % gm, seqs)
    # (b) back-to-back, no snapshots
    seqs = []
    for _ in range(10):
        eng.ingest_f64(0, d, n); seqs.append(eng.ingest_seq())
    show("gm=%d back-to-back  " % gm, seqs)
    eng.snapshot(PS)
    # (c) back-to-back with async snapshot enqueued between launches
    seqs = []
    eng.ingest_f64(0, d, n); seqs.append(eng.ingest_seq())
    hs = []
    for i in range(9):
        eng.snapshot_begin(); h = eng.snapshot_reduce_async(PS); eng.snapshot_end()
        eng.ingest_f64(0, d, n); seqs.append(eng.ingest_seq())
        eng.snapshot_result(h)
    show("gm=%d pipelined snap" % gm, seqs)
    eng.snapshot(PS)
    # (d) back-to-back on half the data twice (same bytes per launch pair)
    seqs = []
    for _ in range(10):
        eng.ingest_f64(0, d, n // 2); seqs.append(eng.ingest_seq())
    show("gm=%d b2b half-size " % gm, seqs)
    eng.snapshot(PS)
for v in (5, 9):
    eng.tune("k1_grid_mult", 1); eng.tune("k1", v)
    seqs = []
    for _ in range(8):
        eng.ingest_f64(0, d, n); seqs.append(eng.ingest_seq())
    show("variant %d back-to-back" % v, seqs)
    eng.snapshot(PS)
