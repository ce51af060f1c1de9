"""Cross-check of the oracle's restated Go arithmetic against arbitrary-precision math (mpmath, 60 digits).

This does not pin Go's last bit (only a Go toolchain can), but it rules out transcription errors in the restated
constants and operation order: a faithful log is within 1 ulp of the true value everywhere, and the bucket
floor(100*ln(1+|v|)+0.5) computed in exact arithmetic agrees with the oracle on every random input (a disagreement
needs the true value within ~1e-13 of a bucket boundary)."""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")


def ulp(x):
    return math.ulp(x)


def test_go_log_within_one_ulp_of_truth(oracle):
    mp.mp.dps = 60
    L = oracle.lib()
    rng = np.random.default_rng(11)
    xs = np.concatenate([1.0 + np.exp(rng.uniform(-40, 0, 4000)), np.exp(rng.uniform(0, 709, 8000))])
    worst = 0.0
    for x in xs:
        x = float(x)
        got = L.lho_go_log(x)
        true = mp.log(mp.mpf(x))
        err = abs(mp.mpf(got) - true) / mp.mpf(ulp(got) if got != 0 else 5e-324)
        worst = max(worst, float(err))
    assert worst < 1.0, worst          # fdlibm's e_log claims < 1 ulp


def test_go_exp_within_a_few_ulp_of_truth(oracle):
    mp.mp.dps = 60
    L = oracle.lib()
    worst_asm = worst_pure = 0.0
    for k in range(0, 32768, 37):
        x = k / 100.0
        true = mp.exp(mp.mpf(x))
        for fn, name in ((L.lho_go_exp, "asm"), (L.lho_go_exp_purego, "pure")):
            got = fn(x)
            if math.isinf(got):
                continue
            err = float(abs(mp.mpf(got) - true) / mp.mpf(ulp(got)))
            if name == "asm":
                worst_asm = max(worst_asm, err)
            else:
                worst_pure = max(worst_pure, err)
    assert worst_pure < 1.0 and worst_asm < 3.0, (worst_pure, worst_asm)


def test_buckets_agree_with_exact_arithmetic_on_random_inputs(oracle):
    mp.mp.dps = 60
    for kind in (oracle.STREAM_U, oracle.STREAM_L):
        vals = oracle.gen_stream(kind, 6000, oracle.DEFAULT_SEED ^ 0xABC)
        keys = oracle.compress_many(vals)
        for v, k in zip(vals, keys):
            exact = int(mp.floor(100 * mp.log(1 + mp.mpf(float(v))) + mp.mpf("0.5")))
            assert exact == int(k), (v, exact, int(k))
    # small and negative magnitudes too (where 1+|v| rounds)
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.uniform(-3, 3, 3000), np.exp(rng.uniform(-12, 2, 3000)) * rng.choice([-1, 1], 3000)])
    keys = oracle.compress_many(vals)
    bad = 0
    for v, k in zip(vals, keys):
        # Go adds 1.0 + |v| in float64 first: restate that rounding, then exact math
        x = mp.mpf(float(1.0 + abs(float(v))))
        exact = int(mp.floor(100 * mp.log(x) + mp.mpf("0.5")))
        exact = -exact if v < 0 else exact
        bad += exact != int(k)
    assert bad == 0
