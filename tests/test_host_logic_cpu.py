"""Host logic on CPU: the C++ MetricSystem mirror (loghisto_b200/host/metric_system.cc) compiled against a TEST-ONLY
stub of the C ABI that is backed by the oracle (tests/stub_abi/lh_stub.c), then driven through the same replays of the
reference's metrics_test.go that the GPU suite runs.  What this covers without a GPU: name interning, staging batches
and flushes, RawMetricSet / ProcessedMetricSet reconstruction, counter store and rates, percentile labels, the reaper,
channel back-pressure.  The product never loads the stub (it has no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def stub_host_lib():
    os.makedirs(BUILD, exist_ok=True)
    stub = os.path.join(BUILD, "liblh_stub.so")
    host = os.path.join(BUILD, "libloghisto_host_stub.so")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I", inc,
                    os.path.join(ROOT, "tests", "stub_abi", "lh_stub.c"), os.path.join(ROOT, "oracle", "loghisto_oracle.c"),
                    "-o", stub, "-lm", "-lpthread"], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", inc,
                    os.path.join(ROOT, "loghisto_b200", "host", "metric_system.cc"),
                    os.path.join(ROOT, "loghisto_b200", "host", "print_benchmark.cc"), "-o", host,
                    "-L", BUILD, "-llh_stub", "-Wl,-rpath," + BUILD, "-lpthread"], check=True)
    return host


@pytest.fixture()
def MS(stub_host_lib, monkeypatch):
    import loghisto_b200.metric_system as m
    monkeypatch.setattr(m, "_lib", m._bind(ctypes.CDLL(stub_host_lib)))
    made = []

    def make(interval_s=1e-6, **kw):
        ms = m.MetricSystem(interval_s, False, max_histograms=kw.get("max_histograms", 64),
                            max_counters=kw.get("max_counters", 64))
        made.append(ms)
        return ms
    yield make
    for ms in made:
        ms.close()


def _cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("host_ms_cases", os.path.join(ROOT, "tests", "test_host_metric_system.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_example_metric_system_keys(MS):
    _cases().test_example_metric_system_keys(MS)


def test_timer(MS):
    _cases().test_timer(MS)


def test_rate(MS):
    _cases().test_rate(MS)


def test_counter(MS):
    _cases().test_counter(MS)


def test_processed_broadcast(MS):
    _cases().test_processed_broadcast(MS)


def test_raw_broadcast(MS):
    _cases().test_raw_broadcast(MS)


def test_slow_subscriber_is_closed_not_blocked_on(MS):
    _cases().test_slow_subscriber_is_closed_not_blocked_on(MS)


def test_stop_is_idempotent_and_leaves_no_thread(MS):
    _cases().test_stop_is_idempotent_and_leaves_no_thread(MS)


def test_mixed_ops_match_oracle_port(MS, oracle):
    _cases().test_mixed_ops_match_oracle_port(MS, oracle)


def test_names_beyond_capacity_are_dropped_and_counted(MS):
    _cases().test_names_beyond_capacity_are_dropped_and_counted(MS)


def test_print_benchmark_driver(stub_host_lib, monkeypatch):
    """print_benchmark.go:49-70: concurrent StartTimer/Stop loops feed one histogram; every interval reports a count."""
    import loghisto_b200.metric_system as m
    monkeypatch.setattr(m, "_lib", m._bind(ctypes.CDLL(stub_host_lib)))
    count = m.PrintBenchmark("benchmark1234", 4, seconds=0.6, interval_s=0.1)
    assert count > 100


def test_counter_zero_amount_appears_in_rates(MS):
    _cases().test_counter_zero_amount_appears_in_rates(MS)


def test_per_call_api_small(MS):
    """The per-call fast path (thread-local name cache, spin-locked shards) against the oracle-backed stub."""
    import numpy as np
    from oracle import oracle as o
    H, n = 16, 300_000
    names = ["h%d" % i for i in range(H)]
    ms = MS(interval_s=3600.0, max_histograms=H)
    ms.histogram_stream(names, o.STREAM_U, o.DEFAULT_SEED, 7, n, 4)
    raw, _ = ms.collect_and_process()
    want = o.stream_ingest_keyed(o.STREAM_U, n, H, o.DEFAULT_SEED, val_start=7, ids_start=7)
    for h in range(H):
        got = np.zeros(65536, dtype=np.uint64)
        for k, c in raw["Histograms"].get(names[h], {}).items():
            got[k & 0xFFFF] = c
        assert (got == want[h]).all(), h


@pytest.mark.parametrize("shard_lock", ["0", "1"])
def test_collects_race_with_per_call_ingest(MS, monkeypatch, shard_lock):
    """12 threads call Histogram() while the main thread collects in a loop: with 4 exclusive shards (membarrier
    handshake between owner and collector; shard_lock=1: the spin-locked fallback) plus the shared overflow shards,
    no sample may be lost or counted twice, and the union of all intervals must be the oracle's histogram."""
    import threading
    import numpy as np
    from oracle import oracle as o
    monkeypatch.setenv("LOGHISTO_B200_SHARDS", "4")
    monkeypatch.setenv("LOGHISTO_B200_STAGING_BYTES", "65536")
    monkeypatch.setenv("LOGHISTO_B200_SHARD_LOCK", shard_lock)
    H, n = 8, 1_200_000
    names = ["name%d" % i for i in range(H)]
    ms = MS(interval_s=3600.0, max_histograms=H)
    total = np.zeros((H, 65536), dtype=np.uint64)

    def add(raw):
        for h in range(H):
            for k, c in raw["Histograms"].get(names[h], {}).items():
                total[h, k & 0xFFFF] += c

    for rnd in range(2):          # the second round's threads take over the shards the first round's threads handed back
        t = threading.Thread(target=ms.histogram_stream, args=(names, o.STREAM_U, o.DEFAULT_SEED, rnd * n, n, 12))
        t.start()
        collects = 0
        while t.is_alive():
            raw, _ = ms.collect_and_process()
            add(raw)
            collects += 1
        t.join()
        raw, _ = ms.collect_and_process()
        add(raw)
        assert collects >= 1
    want = o.stream_ingest_keyed(o.STREAM_U, 2 * n, H, o.DEFAULT_SEED)
    assert ms.dropped() == 0
    assert int(total.sum()) == 2 * n
    assert (total == want).all()


def test_name_lengths_around_the_cache_word_boundaries(MS):
    """The per-thread name cache compares names of up to 16 bytes as two (overlapping) 8-byte words and hashes tails with
    overlapping loads: names of every length 0..40, some differing only in one byte at either end, must stay distinct."""
    names = []
    for n in range(0, 41):
        base = ("abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGH")[:n]
        names.append(base)
        if n:
            names.append(base[:-1] + "#")            # differs in the last byte
            names.append("#" + base[1:])             # differs in the first byte
        if n >= 9:
            names.append(base[:4] + "#" + base[5:])  # differs in the middle of the first word
    names = list(dict.fromkeys(names))
    ms = MS(interval_s=3600.0, max_histograms=len(names) + 4, max_counters=len(names) + 4)
    for rnd in range(3):                                 # the second and third rounds hit the caches
        for i, nm in enumerate(names):
            for _ in range(i % 5 + 1):
                ms.Histogram(nm, float(i))
            ms.Counter(nm, i + 1)
    raw, _ = ms.collect_and_process()
    assert len(raw["Histograms"]) == len(names) and len(raw["Rates"]) == len(names)
    for i, nm in enumerate(names):
        assert sum(raw["Histograms"][nm].values()) == 3 * (i % 5 + 1), (i, nm)
        assert raw["Rates"][nm] == 3 * (i + 1), (i, nm)
