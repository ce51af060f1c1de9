"""Configurable `precision` (reference metrics.go:40-43: `precision = 100`; compress/decompress :316-332 use it):
every kernel family against the oracle at three precisions.  Bar: bit-exact keys, counts, percentile buckets and
decompressed values."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x10C415C0
PS = [0.0, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 0.9999, 1.0]
PRECISIONS = [50, 100, 200]


@pytest.fixture(scope="module")
def lh():
    import loghisto_b200
    return loghisto_b200


def dense_from_sparse(sp, hid):
    out = np.zeros(65536, dtype=np.uint64)
    for k, c in sp.histogram(hid).items():
        out[k & 0xFFFF] = c
    return out


def thresholds(oracle, precision, kmax):
    """T[k] = smallest positive double (as bits) whose un-wrapped bucket is >= k, for k = 1..kmax (bisection on the oracle)."""
    ks = np.arange(1, kmax + 1, dtype=np.int64)
    lo = np.zeros(ks.size, dtype=np.uint64)
    hi = np.full(ks.size, 0x7FEFFFFFFFFFFFFF, dtype=np.uint64)

    def pre_wrap(bits):
        v = bits.view(np.float64)
        k16 = oracle.compress_many(v, precision).astype(np.int64) & 0xFFFF
        approx = np.floor(precision * np.log1p(v) + 0.5)
        wraps = np.round((approx - k16) / 65536.0)
        return k16 + wraps.astype(np.int64) * 65536
    for _ in range(64):
        mid = lo + (hi - lo) // np.uint64(2)
        ge = pre_wrap(mid) >= ks
        hi = np.where(ge, mid, hi)
        lo = np.where(ge, lo, mid)
    return hi


@pytest.mark.parametrize("precision", PRECISIONS)
def test_compress_at_every_threshold(lh, oracle, precision):
    """Every bucket boundary of the whole finite range, +-3 ulps, both signs, both evaluators, at this precision."""
    kmax = int(np.floor(precision * np.log1p(1.7976931348623157e308) + 0.5))
    T = thresholds(oracle, precision, kmax)
    offs = np.arange(-3, 4, dtype=np.int64)
    bits = (T[:, None].astype(np.int64) + offs[None, :]).reshape(-1).astype(np.uint64)
    bits = np.concatenate([bits, bits | np.uint64(0x8000000000000000)])
    vals = bits.view(np.float64)
    want = oracle.compress_many(vals, precision)
    with lh.Engine(device=0, precision=precision) as eng:
        for mode in (0, 1):
            got = eng.compress(vals, mode)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (precision, mode, bad.size, vals[bad[:5]], got[bad[:5]], want[bad[:5]])
        tab = eng.decompress_table()
        assert (tab.view(np.uint64) == oracle.decompress_table(precision).view(np.uint64)).all()
        # the FP32 estimate stays inside its epsilon at this precision
        d = eng.gen_stream(lh.STREAM_U, 20_000_000, SEED ^ 0x77)
        err, slow = eng.fastpath_margin(d, 20_000_000)
        eps = 2.0 ** -12 * max(1.0, precision / 100.0)
        assert err < eps / 2, (precision, err)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_every_ingest_kernel(lh, oracle, precision):
    n = 2_000_003
    for stream in (lh.STREAM_S, lh.STREAM_U, lh.STREAM_N):
        vals = oracle.gen_stream(stream, n, SEED ^ precision)
        want = oracle.ingest(vals, precision=precision)
        ref = oracle.process_histogram(want, PS, precision)
        # K1, every variant
        with lh.Engine(device=0, max_histograms=2, precision=precision) as eng:
            d = eng.upload(vals)
            for vi, name in enumerate(eng.k1_variants()):
                if name.startswith("probe"):
                    continue
                eng.tune("k1", vi)
                eng.ingest_f64(1, d, n)
                red, sp = eng.snapshot(PS)
                assert (dense_from_sparse(sp, 1) == want).all(), (precision, stream, name)
                assert int(red.counts[1]) == n and (red.pkeys[1] == ref["pkeys"]).all()
                assert (red.pvals[1].view(np.uint64) == ref["pvals"].view(np.uint64)).all()
                assert abs(red.sums[1] - ref["sum"]) <= 1e-12 * abs(ref["sum"])
        # keyed kernels: few ids (shared-memory windows), many ids (L2 atomics), many ids (owner-partitioned)
        for H, mode in ((3, 0), (300, 1), (300, 2)):
            ids = oracle.gen_ids(0, n, H, SEED ^ precision)
            wantk = np.zeros((H, 65536), dtype=np.uint64)
            keys = oracle.compress_many(vals, precision).view(np.uint16)
            np.add.at(wantk, (ids, keys), 1)
            with lh.Engine(device=0, max_histograms=H, precision=precision) as eng:
                eng.tune("keyed_mode", mode)
                d, di = eng.upload(vals), eng.upload(ids.astype(np.uint16))
                eng.ingest_keyed_f64_u16(di, d, n)
                red, sp = eng.snapshot(PS)
                assert (red.counts == wantk.sum(axis=1)).all(), (precision, stream, H, mode, eng.keyed_kernel_name())
                for h in (0, 1, H - 1):
                    assert (dense_from_sparse(sp, h) == wantk[h]).all(), (precision, stream, H, mode, h)


def test_precision_range_is_checked(lh):
    with pytest.raises(lh.LhError):
        lh.Engine(device=0, precision=251)
    with lh.Engine(device=0, precision=250) as e:     # the largest supported: K1 falls back to the register-pipelined kernel
        vals = np.array([1.0, -1.0, 1e18, 0.0, 3.5], dtype=np.float64)
        d = e.upload(vals)
        e.ingest_f64(0, d, vals.size)
        red, _ = e.snapshot(PS)
        assert int(red.counts[0]) == vals.size
