"""Size-independent properties of the restated reference functions (hypothesis)."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

finite = st.floats(allow_nan=False, allow_infinity=False, width=64)


@settings(max_examples=400, deadline=None)
@given(finite)
def test_sign_symmetry(oracle, v):
    # compress(-v) == -compress(v) with int16 wrap (metrics.go:318-321)
    a, b = oracle.compress(v), oracle.compress(-v)
    assert b == ((-a + 32768) % 65536) - 32768 or (v == 0 and a == b == 0)


@settings(max_examples=400, deadline=None)
@given(st.floats(min_value=0.0, max_value=9.0e18), st.floats(min_value=0.0, max_value=9.0e18))
def test_monotone_below_2_63(oracle, a, b):
    lo, hi = min(a, b), max(a, b)
    assert oracle.compress(lo) <= oracle.compress(hi)


@settings(max_examples=400, deadline=None)
@given(st.floats(min_value=0.51, max_value=1e140))
def test_roundtrip_within_one_percent(oracle, v):
    # the reference's own claim (metrics.go:312-315, readme.md:5): within 1% above 0.51 and below 1e142
    for x in (v, -v):
        r = oracle.decompress(oracle.compress(x))
        assert abs(x / r - 1) <= 0.01


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(min_value=-1e9, max_value=1e9, allow_nan=False), min_size=1, max_size=200),
       st.lists(st.floats(min_value=0.0, max_value=1.0), min_size=1, max_size=6))
def test_percentile_rule(oracle, values, ps):
    counts = oracle.ingest(np.array(values))
    ref = oracle.process_histogram(counts, sorted(ps))
    keys = np.sort(oracle.compress_many(np.array(values)).astype(np.int64))
    n = len(values)
    assert ref["total"] == n
    prev = -10 ** 9
    for p, k in zip(sorted(ps), ref["pkeys"]):
        # first bucket (ascending) whose cumulative share reaches p  (metrics.go:411-416)
        cum = np.searchsorted(keys, k, side="right")
        assert cum / n >= p
        below = np.searchsorted(keys, k, side="left")
        assert below == 0 or below / n < p
        assert k >= prev                       # percentiles are monotone in p
        prev = k
    assert math.isclose(ref["avg"], ref["sum"] / n, rel_tol=1e-15)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5), st.integers(0, 2 ** 64 - 1)), min_size=1, max_size=100))
def test_counters_wrap_like_uint64(oracle, ops):
    ids = np.array([i for i, _ in ops], dtype=np.uint32)
    amts = np.array([a for _, a in ops], dtype=np.uint64)
    got = oracle.counter_add(ids, amts, 6)
    want = [0] * 6
    for i, a in ops:
        want[i] = (want[i] + a) % 2 ** 64
    assert got.tolist() == want
