"""The oracle (oracle/aqe_oracle.c) against the golden vectors minted from the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Bit-exact: exact sums, closed-interval WHERE sums, index sets,
and the CLI estimators on those index sets."""
import hashlib
import os

import numpy as np
import pytest

from conftest import fhex, golden_files, load_golden
from oracle import make_params

FILES = golden_files()
SMALL = [f for f in FILES if "n1000000" not in f]


def test_golden_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=os.path.basename)
def test_generator_and_exact(oracle, path):
    g = load_golden(path)
    rows = oracle.synth(g["n"], seed=g["seed"], dist=g["dist"])
    assert hashlib.sha256(rows.tobytes()).hexdigest() == g["rows_sha256"]
    assert oracle.sum_amount(rows) == fhex(g["sum_amount"])          # strict left-to-right, bit-exact
    assert oracle.avg_amount(rows) == fhex(g["avg_amount"])
    for key, v in g["sum_amount_where"].items():
        lo, hi = (float(x) for x in key.split(","))
        s, c = oracle.sum_amount_where(rows, lo, hi)
        assert s == fhex(v["sum"]) and c == v["count"]
    assert oracle.tree_height(g["n"]) == g["tree_height"]
    assert oracle.node_count(g["n"]) == g["node_count"]
    for col, want in g["int_sums"].items():
        assert oracle.scan(rows, agg_col=col).isum == want
    # exactly rounded sum agrees with the reference's serial sum to 1e-12 (the fp64 gate)
    exact = oracle.sum_amount_exactly_rounded(rows)
    assert abs(exact - fhex(g["sum_amount"])) <= 1e-12 * abs(exact)


@pytest.mark.parametrize("path", SMALL, ids=os.path.basename)
def test_sampler_index_sets(oracle, path):
    g = load_golden(path)
    rows = oracle.synth(g["n"], seed=g["seed"])
    order = None
    for v in g["samplers"]:
        prm = make_params(v["method"], v["percent"], **v["kw"])
        idx = oracle.indices(rows, v["method"], prm)
        if v["method"] == "stratified_block":
            if order is None:
                order = np.argsort(rows["amount"], kind="stable")
            idx = order[idx]
        tag = (v["method"], v["percent"], v["kw"])
        assert len(idx) == v["count"], tag
        assert hashlib.sha256(np.ascontiguousarray(idx, dtype="<i8").tobytes()).hexdigest() == v["idx_sha256"], tag
        if v["est"] is None:
            continue
        # CLI estimators on the same index list: bit-exact (same left-to-right Python-order arithmetic)
        s = oracle.stats(rows, idx)
        assert s.sum == fhex(v["est"]["sample_sum"]), tag
        e_sum, lo, hi = oracle.estimate(s, g["n"], "sum", 1.96, legacy_ci=True)
        assert e_sum == fhex(v["est"]["sum"]), tag
        e_avg, alo, ahi = oracle.estimate(s, g["n"], "avg", 1.96)
        assert e_avg == fhex(v["est"]["avg"]), tag
        if "m2" in v["est"]:
            assert s.m2 == fhex(v["est"]["m2"]), tag
            assert [lo, hi] == [fhex(x) for x in v["est"]["sum_ci_legacy"]], tag
            assert [alo, ahi] == [fhex(x) for x in v["est"]["avg_ci"]], tag


def test_sampler_index_sets_1m_subset(oracle):
    """One pass at 1M rows over the CLI-facing methods (random / clt / parallel / block)."""
    big = [f for f in FILES if "n1000000" in f]
    if not big:
        pytest.skip("1M golden not minted")
    g = load_golden(big[0])
    rows = oracle.synth(g["n"], seed=g["seed"])
    for v in g["samplers"]:
        if v["method"] not in ("memory_stride", "parallel_pointer", "block", "parallel_block", "optimized_clt",
                               "random_pointer", "index_based", "balanced_tree"):
            continue
        if v["kw"]:
            continue
        idx = oracle.indices(rows, v["method"], make_params(v["method"], v["percent"]))
        assert len(idx) == v["count"]
        assert hashlib.sha256(np.ascontiguousarray(idx, dtype="<i8").tobytes()).hexdigest() == v["idx_sha256"]
