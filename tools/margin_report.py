"""Measured error of the two FP32 bucket estimators (bucket units) on the device, per stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = 100_000_000
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
print("# lh_fastpath_margin over %d samples per stream; LH_FAST_EPS = 2^-12 = %.4e bucket units" % (n, 2.0 ** -12))
for name, kind in (("U", 0), ("L", 1), ("S", 2)):
    d = eng.gen_stream(kind, n, lh.DEFAULT_SEED ^ 0x77)
    err, slow = eng.fastpath_margin(d, n)
    e1, e2 = eng.fastpath_margin_detail()
    print("stream %s: max|estimate - 100 ln(1+|v|)|  fast_candidate %.4e   packed-FP32 (K1) %.4e   exact-path samples %d (%.4f %%)"
          % (name, e1, e2, slow, 100.0 * slow / n), flush=True)
    d.free()
