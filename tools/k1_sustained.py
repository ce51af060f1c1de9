"""K1 variants in the sustained regime: back-to-back launches for `seconds` each, mean device time of the second half
(CUDA events around every launch, lh_kernel_ms), per value stream.  The burst figure is the mean of the first 8.

  python tools/k1_sustained.py [n] [seconds] [variants] [streams]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
names = eng.k1_variants()
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [i for i, nm in enumerate(names)]
streams = sys.argv[4].split(",") if len(sys.argv) > 4 else ["U", "N"]
kinds = {"U": 0, "L": 1, "S": 2, "C": 3, "N": 8}
d = eng.alloc(n, "float64")
for sname in streams:
    eng.gen_stream(kinds[sname], n, lh.DEFAULT_SEED, out=d)
    eng.sync()
    for vi in variants:
        eng.tune("k1", vi)
        est = 1.2e-3 * n / 1e9
        iters = max(16, int(seconds / est))
        ms = []
        for i in range(iters):
            eng.ingest_f64(0, d, n)
            if i >= 8:
                ms.append(eng.kernel_ms(eng.ingest_seq() - 8))   # read 8 launches behind: the queue never drains
        for j in range(8):
            ms.append(eng.kernel_ms(eng.ingest_seq() - 7 + j))
        red, _ = eng.snapshot([0.5], export=False)
        burst = sum(ms[:8]) / 8
        tail = ms[len(ms) // 2:]
        sus = sum(tail) / len(tail)
        ok = names[vi].startswith("probe") or int(red.counts[0]) == iters * n
        print("stream %s  variant %d %-34s burst %.4f ms (%.0f GB/s)  sustained %.4f ms (%.0f GB/s, %d launches)  count_ok=%s"
              % (sname, vi, names[vi], burst, n * 8 / burst / 1e6, sus, n * 8 / sus / 1e6, iters, ok), flush=True)
eng.close()
