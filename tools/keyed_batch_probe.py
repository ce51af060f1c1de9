"""Write-combining keyed kernel vs batch size: time = a + b * n?  Also the fused Histogram+Timer launch vs two launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

H = 1024
nmax = 400_000_000
eng = lh.Engine(device=0, max_histograms=H, max_counters=1)
d = eng.gen_stream(lh.STREAM_U, nmax, lh.DEFAULT_SEED)
ns = eng.gen_stream(lh.STREAM_TIMER_NS, nmax // 4, lh.DEFAULT_SEED)
ids = eng.gen_ids_u16(0, nmax, H, lh.DEFAULT_SEED)
eng.tune("keyed_mode", 2)
for shape in (6, 4):
    eng.tune("wc_spt", shape)
    for n in (8_000_000, 25_000_000, 50_000_000, 75_000_000, 100_000_000, 200_000_000, 400_000_000):
        t = []
        for _ in range(5):
            eng.ingest_keyed_f64_u16(ids, d, n)
            t.append(eng.last_kernel_ms())
        eng.snapshot([0.5])
        ms = sorted(t)[2]
        print("shape %d  n=%10d  %8.1f us  %6.1f G samples/s" % (shape, n, ms * 1e3, n / ms / 1e6), flush=True)
    for nf, nn in ((50_000_000, 25_000_000), (200_000_000, 100_000_000)):
        t2, t1 = [], []
        for _ in range(5):
            eng.ingest_keyed_f64_u16(ids, d, nf); a = eng.last_kernel_ms()
            eng.ingest_keyed_i64ns_u16(ids, ns, nn); b = eng.last_kernel_ms()
            t2.append(a + b)
            eng.ingest_keyed_pair_u16(ids, d, nf, ids, ns, nn); t1.append(eng.last_kernel_ms())
        eng.snapshot([0.5])
        print("shape %d  %d f64 + %d ns: two launches %8.1f us, one fused launch %8.1f us" % (shape, nf, nn, sorted(t2)[2] * 1e3, sorted(t1)[2] * 1e3), flush=True)
eng.close()
