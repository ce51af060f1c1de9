"""Times the keyed path at H histograms over kernel choice, tile shape, chunk size, value stream and id skew, and checks
every configuration's buckets against the first one (the L2-atomic kernel) -- the full oracle comparison is
tests/test_gpu_parity.py::test_full_size_keyed_1024 and bench.py --workload c3.

  python tools/keyed_sweep.py [n] [H] [quick]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
quick = len(sys.argv) > 3
eng = lh.Engine(device=0, max_histograms=H, max_counters=1)
d = eng.alloc(n, "float64")
ids = eng.alloc(n, "uint16")
# (label, keyed_mode, tile shape code, chunk samples, samples between flushes)
configs = [("vec", 1, 4, 8 << 20, 24576)]
if quick:
    for spt in (6, 4, 3, 8):
        configs.append(("wc", 2, spt, 32 << 20, 24576))
else:
    for spt in (6, 4, 3):
        for chunk in (16 << 20, 32 << 20, 64 << 20):
            for flush in (16384, 24576, 28672):
                configs.append(("wc", 2, spt, chunk, flush))
for sname, kind, idkind in (("U", 0, 0), ("L", 1, 0), ("C", 3, 0), ("U/zipf-ids", 0, 1)):
    eng.gen_stream(kind, n, lh.DEFAULT_SEED, out=d)
    eng.gen_ids_u16(idkind, n, H, lh.DEFAULT_SEED, out=ids)
    ref = None
    for name, mode, spt, chunk, flush in configs:
        eng.tune("keyed_mode", mode); eng.tune("wc_spt", spt); eng.tune("kp_chunk", chunk); eng.tune("wc_flush", flush)
        t = []
        for _ in range(4):
            eng.ingest_keyed_f64_u16(ids, d, n)
            t.append(eng.last_kernel_ms())
        red, sp = eng.snapshot([0.5], export=True)
        sig = (sp.offsets.tobytes(), sp.keys.tobytes(), sp.counts.tobytes())
        if ref is None:
            ref = sig
        ms = sorted(t)[1]
        print("H=%-4d stream %-10s %-4s shape=%-2d chunk=%-9d flush=%-5d %8.3f ms %7.1f G samples/s %5.2f TB/s  kernel=%s  count_ok=%s same_buckets=%s"
              % (H, sname, name, spt, chunk, flush, ms, n / ms / 1e6, n * 10 / ms / 1e9, eng.keyed_kernel_name(),
                 int(red.counts.sum()) == 4 * n, sig == ref), flush=True)
eng.close()
