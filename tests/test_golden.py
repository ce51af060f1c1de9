"""Committed golden vectors: the oracle must still reproduce them (CPU), and the CUDA path must too (GPU)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
KINDS = {"U": 0, "L": 1, "S": 2}
SEED = 0x10C415C0


def load_counts(name):
    out = np.zeros(65536, dtype=np.uint64)
    for line in open(os.path.join(GOLD, "stream_%s_1e6.counts" % name)):
        if line.startswith("#"):
            continue
        k, c = line.split()
        out[int(k) & 0xFFFF] = int(c)
    return out


def summaries():
    return json.load(open(os.path.join(GOLD, "stream_summaries.json")))


@pytest.mark.parametrize("name", ["U", "L", "S"])
def test_oracle_reproduces_golden(oracle, name):
    want = load_counts(name)
    assert int(want.sum()) == 1_000_000
    vals = oracle.gen_stream(KINDS[name], 1_000_000, SEED)
    s = summaries()
    assert [float(v).hex() for v in vals[:4]] == s["streams"][name]["first_values_hex"]
    assert (oracle.ingest(vals) == want).all()
    ref = oracle.process_histogram(want, s["percentiles"])
    g = s["streams"][name]
    assert [int(k) for k in ref["pkeys"]] == g["pkeys"]
    assert [float(v).hex() for v in ref["pvals"]] == g["pvals_hex"]
    assert float(ref["sum"]).hex() == g["sum_hex"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["U", "L", "S"])
def test_cuda_path_reproduces_golden(name):
    import loghisto_b200 as lh
    s = summaries()
    g = s["streams"][name]
    want = load_counts(name)
    with lh.Engine(device=0, max_histograms=1, max_counters=1) as eng:
        d = eng.gen_stream(KINDS[name], 1_000_000, SEED)
        eng.ingest_f64(0, d, 1_000_000)
        red, sp = eng.snapshot(s["percentiles"])
        got = np.zeros(65536, dtype=np.uint64)
        for k, c in sp.histogram(0).items():
            got[k & 0xFFFF] = c
        assert (got == want).all()
        assert [int(k) for k in red.pkeys[0]] == g["pkeys"]
        assert [float(v).hex() for v in red.pvals[0]] == g["pvals_hex"]
        assert abs(red.sums[0] - float.fromhex(g["sum_hex"])) <= 1e-12 * abs(float.fromhex(g["sum_hex"]))
        assert int(red.counts[0]) == g["count"]
