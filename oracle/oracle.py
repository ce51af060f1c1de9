"""ctypes binding of the CPU oracle (oracle/loghisto_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(loghisto_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liblh_oracle.so")

STREAM_U, STREAM_L, STREAM_S, STREAM_C, STREAM_Z = 0, 1, 2, 3, 4
STREAM_RAW, STREAM_TIMER_NS, STREAM_AMOUNTS = 5, 6, 7   # raw u64 bits / int64 ns / counter amounts 1..16
STREAM_N = 8                                            # stream U with a random sign
DEFAULT_SEED = 0x10C415C0

DEFAULT_PERCENTILES = {
    "%s_min": 0.0, "%s_50": 0.5, "%s_75": 0.75, "%s_90": 0.9, "%s_95": 0.95,
    "%s_99": 0.99, "%s_99.9": 0.999, "%s_99.99": 0.9999, "%s_max": 1.0,
}


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "loghisto_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    dp = C.POINTER(C.c_double)
    u64p = C.POINTER(C.c_uint64)
    u32p = C.POINTER(C.c_uint32)
    L.lho_go_log.restype = C.c_double
    L.lho_go_log.argtypes = [C.c_double]
    L.lho_go_exp.restype = C.c_double
    L.lho_go_exp.argtypes = [C.c_double]
    L.lho_go_exp_purego.restype = C.c_double
    L.lho_go_exp_purego.argtypes = [C.c_double]
    L.lho_compress.restype = C.c_int16
    L.lho_compress.argtypes = [C.c_double]
    L.lho_decompress.restype = C.c_double
    L.lho_decompress.argtypes = [C.c_int16]
    L.lho_decompress_purego.restype = C.c_double
    L.lho_decompress_purego.argtypes = [C.c_int16]
    L.lho_go_f64_to_u64.restype = C.c_uint64
    L.lho_go_f64_to_u64.argtypes = [C.c_double]
    L.lho_compress_p.restype = C.c_int16
    L.lho_compress_p.argtypes = [C.c_double, C.c_double]
    L.lho_decompress_p.restype = C.c_double
    L.lho_decompress_p.argtypes = [C.c_int16, C.c_double]
    L.lho_compress_many_p.argtypes = [dp, C.c_size_t, C.POINTER(C.c_int16), C.c_double]
    L.lho_ingest_p.argtypes = [dp, C.c_size_t, u64p, C.c_double]
    L.lho_process_histogram_p.restype = C.c_uint64
    L.lho_process_histogram_p.argtypes = [u64p, dp, C.c_int, dp, dp, C.POINTER(C.c_int32), C.c_double]
    L.lho_ingest.argtypes = [dp, C.c_size_t, u64p]
    L.lho_ingest_mt.argtypes = [dp, C.c_size_t, u64p, C.c_int]
    L.lho_compress_many.argtypes = [dp, C.c_size_t, C.POINTER(C.c_int16)]
    L.lho_ingest_keyed.argtypes = [u32p, dp, C.c_size_t, u64p]
    L.lho_ingest_keyed_i64.argtypes = [u32p, C.POINTER(C.c_int64), C.c_size_t, u64p]
    L.lho_counter_add.argtypes = [u32p, u64p, C.c_size_t, u64p]
    L.lho_process_histogram.restype = C.c_uint64
    L.lho_process_histogram.argtypes = [u64p, dp, C.c_int, dp, dp, C.POINTER(C.c_int32)]
    L.lho_percentile.restype = C.c_int
    L.lho_percentile.argtypes = [C.c_uint64, dp, u64p, C.c_int, C.c_double, dp]
    L.lho_stream_bits.restype = C.c_uint64
    L.lho_stream_bits.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    L.lho_gen_stream.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, dp]
    L.lho_gen_ids.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, C.c_uint32, u32p]
    L.lho_stream_ingest_mt.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, u64p, C.c_int]
    L.lho_stream_ingest_keyed_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t,
                                             C.c_uint32, u64p, C.c_int]
    L.lho_stream_counter_mt.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, C.c_uint32,
                                        u64p, C.c_int]
    L.lho_ms_new.restype = C.c_void_p
    L.lho_ms_free.argtypes = [C.c_void_p]
    L.lho_ms_specify_percentiles.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), dp]
    L.lho_ms_histogram.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.lho_ms_counter.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    L.lho_ms_collect_and_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lho_ms_bench_ingest.restype = C.c_double
    L.lho_ms_bench_ingest.argtypes = [C.c_void_p, dp, u32p, C.c_size_t, C.POINTER(C.c_char_p), C.c_int]
    L.lho_ms_peek_bucket.restype = C.c_uint64
    L.lho_ms_peek_bucket.argtypes = [C.c_void_p, C.c_char_p, C.c_int16]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def compress(v: float, precision: float = 100.0) -> int:
    return int(lib().lho_compress_p(float(v), float(precision)))


def decompress(k: int, precision: float = 100.0) -> float:
    return float(lib().lho_decompress_p(int(k), float(precision)))


def compress_many(vals: np.ndarray, precision: float = 100.0) -> np.ndarray:
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.empty(vals.size, dtype=np.int16)
    lib().lho_compress_many_p(_dp(vals), vals.size, out.ctypes.data_as(C.POINTER(C.c_int16)), float(precision))
    return out


def decompress_table(precision: float = 100.0) -> np.ndarray:
    """out[(uint16)key] = decompress(key) for every int16 key."""
    out = np.empty(65536, dtype=np.float64)
    for k in range(-32768, 32768):
        out[k & 0xFFFF] = lib().lho_decompress_p(k, float(precision))
    return out


def ingest(vals: np.ndarray, counts: np.ndarray | None = None, threads: int = 1, precision: float = 100.0) -> np.ndarray:
    """Dense histogram: counts[(uint16)compress(v)] += 1 (uint64[65536])."""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    if counts is None:
        counts = np.zeros(65536, dtype=np.uint64)
    if precision != 100.0:
        lib().lho_ingest_p(_dp(vals), vals.size, _u64p(counts), float(precision))
        return counts
    if threads > 1:
        lib().lho_ingest_mt(_dp(vals), vals.size, _u64p(counts), threads)
    else:
        lib().lho_ingest(_dp(vals), vals.size, _u64p(counts))
    return counts


def stream_ingest(kind: int, n: int, seed: int = DEFAULT_SEED, start: int = 0, counts: np.ndarray | None = None,
                  threads: int | None = None) -> np.ndarray:
    """Dense histogram of stream `kind` over indices [start, start+n), regenerated and bucketed on the fly on
    `threads` host threads (default: all) -- the full-size checker for 1e9 / 1e10-sample device runs."""
    if counts is None:
        counts = np.zeros(65536, dtype=np.uint64)
    lib().lho_stream_ingest_mt(kind, seed, start, n, _u64p(counts), threads or os.cpu_count() or 1)
    return counts


def stream_ingest_keyed(val_kind: int, n: int, n_histograms: int, seed: int = DEFAULT_SEED, val_start: int = 0,
                        ids_start: int = 0, id_kind: int = 0, as_i64: bool = False,
                        counts: np.ndarray | None = None, threads: int | None = None) -> np.ndarray:
    """counts[id_i][(uint16)compress(v_i)] += 1 for the (ids, values) streams, regenerated on the fly."""
    if counts is None:
        counts = np.zeros((n_histograms, 65536), dtype=np.uint64)
    lib().lho_stream_ingest_keyed_mt(val_kind, id_kind, 1 if as_i64 else 0, seed, val_start, ids_start, n,
                                     n_histograms, _u64p(counts), threads or os.cpu_count() or 1)
    return counts


def stream_counter(n: int, n_counters: int, seed: int = DEFAULT_SEED, val_start: int = 0, ids_start: int = 0,
                   id_kind: int = 0, counters: np.ndarray | None = None, threads: int | None = None) -> np.ndarray:
    """counters[id_i] += amount_i for the (ids, amounts = stream A) streams, regenerated on the fly."""
    if counters is None:
        counters = np.zeros(n_counters, dtype=np.uint64)
    lib().lho_stream_counter_mt(STREAM_AMOUNTS, id_kind, seed, val_start, ids_start, n, n_counters, _u64p(counters),
                                threads or os.cpu_count() or 1)
    return counters


def ingest_keyed(ids: np.ndarray, vals: np.ndarray, n_histograms: int,
                 counts: np.ndarray | None = None) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    assert ids.size == vals.size
    if counts is None:
        counts = np.zeros((n_histograms, 65536), dtype=np.uint64)
    lib().lho_ingest_keyed(_u32p(ids), _dp(vals), vals.size, _u64p(counts))
    return counts


def ingest_keyed_i64(ids: np.ndarray, ns: np.ndarray, n_histograms: int,
                     counts: np.ndarray | None = None) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    ns = np.ascontiguousarray(ns, dtype=np.int64)
    if counts is None:
        counts = np.zeros((n_histograms, 65536), dtype=np.uint64)
    lib().lho_ingest_keyed_i64(_u32p(ids), ns.ctypes.data_as(C.POINTER(C.c_int64)), ns.size, _u64p(counts))
    return counts


def counter_add(ids: np.ndarray, amounts: np.ndarray, n_counters: int,
                counters: np.ndarray | None = None) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    amounts = np.ascontiguousarray(amounts, dtype=np.uint64)
    if counters is None:
        counters = np.zeros(n_counters, dtype=np.uint64)
    lib().lho_counter_add(_u32p(ids), _u64p(amounts), ids.size, _u64p(counters))
    return counters


def process_histogram(counts: np.ndarray, ps, precision: float = 100.0) -> dict:
    """processHistograms on one dense uint64[65536] histogram."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    stats = np.zeros(3, dtype=np.float64)
    pv = np.zeros(ps.size, dtype=np.float64)
    pk = np.zeros(ps.size, dtype=np.int32)
    total = lib().lho_process_histogram_p(_u64p(counts), _dp(ps), ps.size, _dp(stats), _dp(pv),
                                          pk.ctypes.data_as(C.POINTER(C.c_int32)), float(precision))
    return {"total": int(total), "count": stats[0], "sum": stats[1], "avg": stats[2],
            "pvals": pv, "pkeys": pk}


def percentile(total: int, values, counts, p: float):
    values = np.ascontiguousarray(values, dtype=np.float64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    out = C.c_double(0)
    rc = lib().lho_percentile(int(total), _dp(values), _u64p(counts), values.size, float(p), C.byref(out))
    if rc != 0:
        raise ValueError("Invalid percentile.  Should be between 0 and 1.")
    return out.value


def gen_stream(kind: int, n: int, seed: int = DEFAULT_SEED, start: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    lib().lho_gen_stream(kind, seed, start, n, _dp(out))
    return out


def gen_ids(kind: int, n: int, n_histograms: int, seed: int = DEFAULT_SEED, start: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().lho_gen_ids(kind, seed, start, n, n_histograms, _u32p(out))
    return out


_EMIT = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_uint64, C.c_double)


class OracleMetricSystem:
    """Structure-faithful CPU port of MetricSystem's ingest + snapshot + reduce path."""

    def __init__(self):
        self._h = lib().lho_ms_new()

    def close(self):
        if self._h:
            lib().lho_ms_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SpecifyPercentiles(self, percentiles: dict):
        labels = (C.c_char_p * len(percentiles))(*[k.encode() for k in percentiles])
        ps = np.array(list(percentiles.values()), dtype=np.float64)
        lib().lho_ms_specify_percentiles(self._h, len(percentiles), labels, _dp(ps))

    def Histogram(self, name: str, value: float):
        lib().lho_ms_histogram(self._h, name.encode(), float(value))

    def Counter(self, name: str, amount: int):
        lib().lho_ms_counter(self._h, name.encode(), int(amount))

    def collect_and_process(self):
        """Returns (raw, processed): raw = {Counters, Rates, Histograms}; processed = {name: f64}."""
        raw = {"Counters": {}, "Rates": {}, "Histograms": {}}
        processed = {}

        def emit(_ctx, kind, name, key, u, f):
            name = name.decode()
            if kind == 0:
                raw["Counters"][name] = int(u)
            elif kind == 1:
                raw["Rates"][name] = int(u)
            elif kind == 2:
                raw["Histograms"].setdefault(name, {})[int(key)] = int(u)
            else:
                processed[name] = float(f)

        cb = _EMIT(emit)
        lib().lho_ms_collect_and_process(self._h, C.cast(cb, C.c_void_p), None)
        return raw, processed

    def bench_ingest(self, vals: np.ndarray, ids, names, threads: int) -> float:
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            idp = _u32p(ids)
        return float(lib().lho_ms_bench_ingest(self._h, _dp(vals), idp, vals.size, arr, threads))
