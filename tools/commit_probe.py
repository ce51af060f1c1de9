"""Host-side cost of one staging commit (lh_staging_acquire + fill + lh_staging_commit_keyed_f64_u16) from ONE thread:
the serial section every shard of the per-call API path goes through once per 419 K samples."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loghisto_b200 as lh

for H in (16, 1024):
    eng = lh.Engine(device=0, max_histograms=H, max_counters=16, staging_bytes=4 << 20, staging_slots=64)
    cap = ((4 << 20) // 10) & ~15
    vals = np.random.default_rng(1).random(cap) * 1e6
    ids = (np.arange(cap) % min(H, 1024)).astype(np.uint16)
    for rnd in range(3):
        t_acq = t_fill = t_commit = 0.0
        n = 100
        t0 = time.perf_counter()
        for i in range(n):
            a = time.perf_counter()
            s = eng.staging_acquire()
            b = time.perf_counter()
            eng.staging_view(s, np.float64, cap)[:] = vals
            eng.staging_view(s, np.uint16, cap, cap * 8)[:] = ids
            c = time.perf_counter()
            eng.staging_commit_keyed_f64_u16(s, cap, cap * 8)
            d = time.perf_counter()
            t_acq += b - a; t_fill += c - b; t_commit += d - c
        eng.sync()
        dt = time.perf_counter() - t0
        print("H=%d round %d: %d commits of %d pairs: acquire %.1f us, fill %.1f us, commit %.1f us each; total %.3f s -> %.1f M pairs/s through one thread"
              % (H, rnd, n, cap, t_acq / n * 1e6, t_fill / n * 1e6, t_commit / n * 1e6, dt, n * cap / dt / 1e6), flush=True)
    eng.close()
