"""Write-combining keyed kernel over its tunables on one stream (A/B runs of single ideas): python tools/keyed_pf_probe.py key=v1,v2 ..."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

n, H = 1_000_000_000, 1024
axes = [(a.split("=")[0], [int(x) for x in a.split("=")[1].split(",")]) for a in sys.argv[1:]] or [("wc_pf", [0, 1, 2, 4])]
eng = lh.Engine(device=0, max_histograms=H, max_counters=1)
d = eng.gen_stream(lh.STREAM_U, n, lh.DEFAULT_SEED)
ids = eng.gen_ids_u16(0, n, H, lh.DEFAULT_SEED)
eng.tune("keyed_mode", 2)
ref = None
for combo in itertools.product(*[v for _, v in axes]):
    for (k, _), v in zip(axes, combo):
        eng.tune(k, v)
    t = []
    for _ in range(5):
        eng.ingest_keyed_f64_u16(ids, d, n)
        t.append(eng.last_kernel_ms())
    red, sp = eng.snapshot([0.5], export=True)
    sig = (sp.offsets.tobytes(), sp.keys.tobytes(), sp.counts.tobytes())
    ref = ref or sig
    ms = sorted(t)[2]
    print(" ".join("%s=%d" % (k, v) for (k, _), v in zip(axes, combo)), " %8.3f ms  %6.1f G samples/s  count_ok=%s same_buckets=%s"
          % (ms, n / ms / 1e6, int(red.counts.sum()) == 5 * n, sig == ref), flush=True)
eng.close()
