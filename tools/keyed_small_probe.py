"""Times the keyed path across histogram counts and value distributions (default kernel selection)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
hs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 12, 64, 256, 1024]
for H in hs:
    eng = lh.Engine(device=0, max_histograms=H, max_counters=1)
    d = eng.alloc(n, "float64")
    ids = eng.gen_ids_u16(0, n, H, lh.DEFAULT_SEED)
    for sname, kind in (("U", 0), ("L", 1), ("C", 3)):
        eng.gen_stream(kind, n, lh.DEFAULT_SEED, out=d)
        t = []
        for _ in range(4):
            eng.ingest_keyed_f64_u16(ids, d, n)
            t.append(eng.last_kernel_ms())
        red, _ = eng.snapshot([0.5], export=False)
        ms = sorted(t)[1]
        print("H=%-4d stream %s  %.3f ms  %6.1f G samples/s  %.2f TB/s  count_ok=%s" % (H, sname, ms, n / ms / 1e6, n * 10 / ms / 1e9, int(red.counts.sum()) == 4 * n), flush=True)
    d.free(); ids.free(); eng.close()
