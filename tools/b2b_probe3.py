"""Does K1 slow down over a long sustained run?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh
n = 1_000_000_000
eng = lh.Engine(device=0, max_histograms=1, max_counters=1)
d = eng.gen_stream(0, n, lh.DEFAULT_SEED)
eng.sync()
PS = [0.0, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 0.9999, 1.0]
def run(k, tag, snap=True):
    times = []
    eng.ingest_f64(0, d, n)
    t0 = time.perf_counter()
    for i in range(k):
        seq = eng.ingest_seq()
        if snap:
            eng.snapshot_begin(); h = eng.snapshot_reduce_async(PS); eng.snapshot_end()
        if i + 1 < k:
            eng.ingest_f64(0, d, n)
        if snap:
            eng.snapshot_result(h)
        else:
            eng.kernel_ms(seq)
        times.append(eng.kernel_ms(seq))
    wall = (time.perf_counter() - t0) / k * 1e3
    eng.sync()
    pick = [0, 1, 5, 10, 20, 40, 80, 120, 160, 199, 299, 399]
    print("%-20s wall/step %.3f mean %.3f :" % (tag, wall, sum(times) / len(times)),
          " ".join("%d:%.3f" % (i, times[i]) for i in pick if i < k), flush=True)
run(200, "pipelined 200")
run(400, "pipelined 400")
run(200, "no snapshot 200", snap=False)
time.sleep(1.0)
run(50, "after 1s idle, 50")
