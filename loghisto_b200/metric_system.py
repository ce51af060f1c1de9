"""Python face of the C++ MetricSystem mirror (loghisto_b200/host/metric_system.{h,cc}).

Method names follow the Go type (reference metrics.go) so the tests read like metrics_test.go.  All work
happens in the C++ layer and, below it, in the CUDA library; this module only marshals.
"""
from __future__ import annotations

import ctypes as C

from . import build as _build

_EMIT = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_uint64, C.c_double)
_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build() or _build.needs_build_host():
        _build.build_host()
    _lib = _bind(C.CDLL(_build.HOST_LIB))
    return _lib


def _bind(L):
    """Declare the lhms_* signatures on a loaded host library."""
    vp = C.c_void_p
    L.lhms_new.restype = vp
    L.lhms_new.argtypes = [C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_int]
    L.lhms_free.argtypes = [vp]
    L.lhms_histogram.argtypes = [vp, C.c_char_p, C.c_double]
    L.lhms_counter.argtypes = [vp, C.c_char_p, C.c_uint64]
    L.lhms_histogram_many.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    L.lhms_start_timer.restype = vp
    L.lhms_start_timer.argtypes = [vp, C.c_char_p]
    L.lhms_timer_stop.restype = C.c_int64
    L.lhms_timer_stop.argtypes = [vp]
    L.lhms_specify_percentiles.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]
    L.lhms_register_constant_gauge.argtypes = [vp, C.c_char_p, C.c_double]
    L.lhms_collect_and_process.restype = C.c_int
    L.lhms_collect_and_process.argtypes = [vp, _EMIT, vp, C.c_char_p, C.c_int]
    L.lhms_start.argtypes = [vp]
    L.lhms_stop.argtypes = [vp]
    L.lhms_dropped.restype = C.c_uint64
    L.lhms_dropped.argtypes = [vp]
    L.lhms_timer_free.argtypes = [vp]
    L.lhms_histogram_stream.restype = C.c_double
    L.lhms_histogram_stream.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint]
    L.lhms_histogram_stream2.restype = C.c_double
    L.lhms_histogram_stream2.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint,
                                         C.c_int, C.POINTER(C.c_double)]
    L.lhms_timer_loop.restype = C.c_double
    L.lhms_timer_loop.argtypes = [C.c_char_p, C.c_uint, C.c_double, C.c_int64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.lhms_print_benchmark.restype = C.c_double
    L.lhms_print_benchmark.argtypes = [C.c_char_p, C.c_uint, C.c_double, C.c_int64, C.c_int, C.c_int]
    for kind in ("processed", "raw"):
        getattr(L, "lhms_subscribe_" + kind).restype = vp
        getattr(L, "lhms_subscribe_" + kind).argtypes = [vp, C.c_int]
        getattr(L, "lhms_unsubscribe_" + kind).argtypes = [vp, vp]
        getattr(L, "lhms_recv_" + kind).restype = C.c_int
        getattr(L, "lhms_recv_" + kind).argtypes = [vp, C.c_int64, _EMIT, vp]
        getattr(L, "lhms_free_%s_channel" % kind).argtypes = [vp]
    return L


class _Collector:
    def __init__(self):
        self.raw = {"Counters": {}, "Rates": {}, "Histograms": {}, "Gauges": {}}
        self.metrics = {}

        def emit(_ctx, kind, name, key, u, f):
            name = name.decode()
            if kind == 0:
                self.raw["Counters"][name] = int(u)
            elif kind == 1:
                self.raw["Rates"][name] = int(u)
            elif kind == 2:
                self.raw["Histograms"].setdefault(name, {})[int(key)] = int(u)
            elif kind == 4:
                self.raw["Gauges"][name] = float(f)
            else:
                self.metrics[name] = float(f)

        self.cb = _EMIT(emit)


class TimerToken:
    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def Stop(self) -> int:
        """Submits the duration as a histogram sample and returns it in nanoseconds (metrics.go:242-246).  The Go token
        may be stopped repeatedly (one sample each time); this handle is consumed by the first Stop, later calls
        return 0 without submitting anything."""
        if self._h is None:
            return 0
        ns = self._lib.lhms_timer_stop(self._h)
        self._h = None
        return int(ns)

    def __del__(self):
        try:
            if self._h is not None:      # never stopped: release the C++ token
                self._lib.lhms_timer_free(self._h)
                self._h = None
        except Exception:
            pass


class Subscription:
    def __init__(self, ms, kind, capacity):
        self._ms, self._kind = ms, kind
        self._ch = getattr(ms._lib, "lhms_subscribe_" + kind)(ms._h, capacity)

    def receive(self, timeout_s: float):
        """dict of metrics (processed) / raw dict, None on timeout; raises EOFError if the reaper closed the channel."""
        col = _Collector()
        rc = getattr(self._ms._lib, "lhms_recv_" + self._kind)(self._ch, int(timeout_s * 1e9), col.cb, None)
        if rc == 1:
            return col.metrics if self._kind == "processed" else col.raw
        if rc == -1:
            raise EOFError("channel closed by the reaper")
        return None

    def unsubscribe(self):
        if self._ch is not None and self._ms._h:
            getattr(self._ms._lib, "lhms_unsubscribe_" + self._kind)(self._ms._h, self._ch)

    def close(self):
        """Unsubscribe and free the channel (with whatever metric sets are still queued in it)."""
        if self._ch is not None:
            self.unsubscribe()
            getattr(self._ms._lib, "lhms_free_%s_channel" % self._kind)(self._ch)
            self._ch = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MetricSystem:
    """NewMetricSystem(interval, sysStats) -- sysStats (Go runtime gauges) is accepted and ignored."""

    def __init__(self, interval_s: float, sysStats: bool = False, device: int = 0, max_histograms: int = 1024,
                 max_counters: int = 1024):
        """A Channel(capacity=0) is treated as capacity 1 by the C++ mirror (Go's unbuffered rendezvous has no
        equivalent for a non-blocking sender; the reaper never blocks either way, metrics.go:570-573)."""
        self._lib = _load()
        err = C.create_string_buffer(512)
        self._h = self._lib.lhms_new(max(int(interval_s * 1e9), 1), device, max_histograms, max_counters, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())

    def close(self):
        if self._h:
            self._lib.lhms_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SpecifyPercentiles(self, percentiles: dict):
        labels = (C.c_char_p * len(percentiles))(*[k.encode() for k in percentiles])
        ps = (C.c_double * len(percentiles))(*list(percentiles.values()))
        self._lib.lhms_specify_percentiles(self._h, len(percentiles), labels, ps)

    def Histogram(self, name: str, value: float):
        self._lib.lhms_histogram(self._h, name.encode(), float(value))

    def HistogramMany(self, name: str, values):
        import numpy as np
        values = np.ascontiguousarray(values, dtype=np.float64)
        self._lib.lhms_histogram_many(self._h, name.encode(), values.ctypes.data, values.size)

    def Counter(self, name: str, amount: int):
        self._lib.lhms_counter(self._h, name.encode(), int(amount))

    def StartTimer(self, name: str) -> TimerToken:
        return TimerToken(self._lib, self._lib.lhms_start_timer(self._h, name.encode()))

    def RegisterConstantGauge(self, name: str, value: float):
        self._lib.lhms_register_constant_gauge(self._h, name.encode(), float(value))

    def SubscribeToProcessedMetrics(self, capacity: int = 128) -> Subscription:
        return Subscription(self, "processed", capacity)

    def SubscribeToRawMetrics(self, capacity: int = 128) -> Subscription:
        return Subscription(self, "raw", capacity)

    def Start(self):
        self._lib.lhms_start(self._h)

    def Stop(self):
        self._lib.lhms_stop(self._h)

    def collect_and_process(self):
        """processMetrics(collectRawMetrics()) -> (raw dict, metrics dict), as metrics_test.go calls it."""
        col = _Collector()
        err = C.create_string_buffer(512)
        if self._lib.lhms_collect_and_process(self._h, col.cb, None, err, 512) != 0:
            raise RuntimeError(err.value.decode())
        return col.raw, col.metrics

    def dropped(self) -> int:
        return int(self._lib.lhms_dropped(self._h))

    def histogram_stream(self, names, kind: int, seed: int, start: int, n: int, threads: int) -> float:
        """Per-call load generator: `threads` OS threads call Histogram(names[id_i], value_i) once per sample for the
        samples [start, start + n) of synthetic stream `kind`; returns the seconds the calls took."""
        arr = (C.c_char_p * len(names))(*[x.encode() for x in names])
        return float(self._lib.lhms_histogram_stream(self._h, arr, len(names), kind, seed, start, n, threads))

    def histogram_stream_timed(self, names, kind: int, seed: int, start: int, n: int, threads: int, dry: bool = False):
        """As histogram_stream; returns (wall seconds incl. the synthetic generator, largest per-thread seconds spent
        inside the Histogram() call loops alone).  dry=True runs the generator only."""
        arr = (C.c_char_p * len(names))(*[x.encode() for x in names])
        calls = C.c_double()
        wall = self._lib.lhms_histogram_stream2(self._h, arr, len(names), kind, seed, start, n, threads, 1 if dry else 0, C.byref(calls))
        return float(wall), float(calls.value)


def timer_loop(name: str, threads: int, seconds: float, interval_s: float = 0.1, device: int = 0):
    """print_benchmark.go:59-67 as a measurement: returns (calls per second, total calls, sum of the <name>_count
    values every interval reported)."""
    total = C.c_uint64()
    rep = C.c_double()
    rate = _load().lhms_timer_loop(name.encode(), threads, seconds, int(interval_s * 1e9), device, C.byref(total), C.byref(rep))
    return float(rate), int(total.value), float(rep.value)


def PrintBenchmark(name: str, concurrency: int, seconds: float = 3.0, interval_s: float = 1.0, device: int = 0,
                   print_metrics: bool = False) -> float:
    """print_benchmark.go:49 with an empty op: `concurrency` threads loop StartTimer/Stop for `seconds`; returns the
    last interval's <name>_count (timer samples ingested per interval -- the figure readme.md:34 quotes)."""
    return float(_load().lhms_print_benchmark(name.encode(), concurrency, seconds, int(interval_s * 1e9), device,
                                              1 if print_metrics else 0))
