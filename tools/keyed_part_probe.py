"""Runs the owner-partitioned keyed kernel a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
shape = int(sys.argv[2]) if len(sys.argv) > 2 else 0
eng = lh.Engine(device=0, max_histograms=1024, max_counters=1)
d = eng.gen_stream(0, n, lh.DEFAULT_SEED)
ids = eng.gen_ids_u16(0, n, 1024, lh.DEFAULT_SEED)
eng.tune("keyed_mode", 2); eng.tune("kp_shape", shape)
for _ in range(3):
    eng.ingest_keyed_f64_u16(ids, d, n)
    print("ms", eng.last_kernel_ms(), flush=True)
red, _ = eng.snapshot([0.5], export=False)
print("count_ok", int(red.counts.sum()) == 3 * n)
