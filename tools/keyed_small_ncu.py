import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = 200_000_000
eng = lh.Engine(device=0, max_histograms=8, max_counters=1)
d = eng.gen_stream(1, n, lh.DEFAULT_SEED)
ids = eng.gen_ids_u16(0, n, 8, lh.DEFAULT_SEED)
for _ in range(3):
    eng.ingest_keyed_f64_u16(ids, d, n)
    print("ms", eng.last_kernel_ms(), flush=True)
red, _ = eng.snapshot([0.5], export=False)
print("count_ok", int(red.counts.sum()) == 3 * n)
