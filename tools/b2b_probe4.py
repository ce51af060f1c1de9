"""Sustained behaviour of the pure read probe vs K1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = 1_000_000_000
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
d = eng.gen_stream(0, n, lh.DEFAULT_SEED)
eng.sync()
names = eng.k1_variants()
def run(k, v):
    eng.tune("k1", v)
    times = []
    eng.ingest_f64(0, d, n)
    for i in range(k):
        seq = eng.ingest_seq()
        if i + 1 < k:
            eng.ingest_f64(0, d, n)
        times.append(eng.kernel_ms(seq))
    eng.sync()
    pick = [0, 1, 5, 10, 20, 40, 80, 120, 160, 199, 299]
    print("%-28s mean %.3f :" % (names[v], sum(times) / len(times)), " ".join("%d:%.3f" % (i, times[i]) for i in pick if i < k), flush=True)
vs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [11, 0, 11, 0]
for v in vs:
    time.sleep(1.5)
    run(300, v)
