"""Per-call API path (MetricSystem.Histogram once per sample) over a ladder of thread counts and name counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loghisto_b200.metric_system import MetricSystem

ncpu = os.cpu_count() or 1
for names_n in (1, 1024):
    names = ["histogram%d" % i for i in range(names_n)]
    for threads in sorted({1, 32, ncpu}):
        if threads > ncpu:
            continue
        ms = MetricSystem(3600.0, False, device=0, max_histograms=max(16, names_n), max_counters=16)
        n = min(8_000_000 * threads, 500_000_000)
        ms.histogram_stream(names, 0, 0x10C415C0, 0, n // 4, threads)        # warm the staging ring
        dt = ms.histogram_stream(names, 0, 0x10C415C0, n, n, threads)
        raw, _ = ms.collect_and_process()
        got = sum(sum(b.values()) for b in raw["Histograms"].values())
        print("names %4d threads %3d: %6.1f M calls/s  (%5.1f ns per call per thread)  count_ok=%s dropped=%d"
              % (names_n, threads, n / dt / 1e6, dt * 1e9 * threads / n, got == n + n // 4, ms.dropped()), flush=True)
        ms.close()
