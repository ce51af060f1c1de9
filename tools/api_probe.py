"""Per-call API path (MetricSystem.Histogram once per sample) over thread counts, name counts and staging-slot sizes.
Prints the wall rate (synthetic generator included), the rate of the call loops alone, and the generator-only rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loghisto_b200.metric_system import MetricSystem

ncpu = os.cpu_count() or 1
ladder = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else sorted({1, 16, 32, 64, ncpu})
slots = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4 << 20]
names_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 1024]
print("cpus", ncpu, "cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "?", flush=True)
lock_modes = ["0", "1"] if os.environ.get("PROBE_LOCK_AB") else ["0"]
for names_n in names_list:
    names = ["histogram%d" % i for i in range(names_n)]
    for sb, unl in [(a, b) for a in slots for b in lock_modes]:
        os.environ["LOGHISTO_B200_STAGING_BYTES"] = str(sb)
        os.environ["LOGHISTO_B200_SHARD_LOCK"] = unl
        print("  shards: %s" % ("spin-locked (LOGHISTO_B200_SHARD_LOCK=1)" if unl == "1" else "exclusive per thread, membarrier handshake"), flush=True)
        for threads in ladder:
            if threads > ncpu:
                continue
            ms = MetricSystem(3600.0, False, device=0, max_histograms=max(16, names_n), max_counters=16)
            n = min(8_000_000 * threads, 400_000_000)
            dry, _ = ms.histogram_stream_timed(names, 0, 0x10C415C0, 0, n, threads, dry=True)
            ms.histogram_stream(names, 0, 0x10C415C0, 0, n // 4, threads)        # warm the staging ring
            wall, calls = ms.histogram_stream_timed(names, 0, 0x10C415C0, n, n, threads)
            raw, _ = ms.collect_and_process()
            got = sum(sum(b.values()) for b in raw["Histograms"].values())
            print("names %4d slot %4d KiB threads %3d: wall %7.1f M calls/s | call loops only %7.1f M calls/s (%5.1f ns per call per thread) | "
                  "generator alone %8.1f M/s  count_ok=%s dropped=%d"
                  % (names_n, sb >> 10, threads, n / wall / 1e6, n / calls / 1e6, calls * 1e9 * threads / n, n / dry / 1e6,
                     got == n + n // 4, ms.dropped()), flush=True)
            ms.close()
