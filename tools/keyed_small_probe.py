"""Times the keyed path for small histogram counts (k_ingest_keyed_small) against the L2-atomic kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = 1_000_000_000
for H in (1, 8, 11, 12, 64):
    eng = lh.Engine(device=0, max_histograms=H, max_counters=1)
    d = eng.gen_stream(1, n, lh.DEFAULT_SEED)
    ids = eng.gen_ids_u16(0, n, H, lh.DEFAULT_SEED)
    for mode in (0, 1):
        eng.tune("keyed_mode", mode)
        t = []
        for _ in range(5):
            eng.ingest_keyed_f64_u16(ids, d, n)
            t.append(eng.last_kernel_ms())
        red, _ = eng.snapshot([0.5], export=False)
        ms = sorted(t)[2]
        print("H=%-3d mode=%d  %.3f ms  %.1f G samples/s  %.2f TB/s  count_ok=%s" % (H, mode, ms, n / ms / 1e6, n * 10 / ms / 1e9, int(red.counts.sum()) == 5 * n), flush=True)
    d.free(); ids.free(); eng.close()
