"""Host logic of the product: the closed-form sample plans (csrc/aqe_plan.cpp) expand to exactly the
position lists of the oracle (oracle/aqe_oracle.c, pinned to the reference) and of the golden vectors
minted from the reference itself.  CPU only -- plans are pure arithmetic."""
import hashlib
import os

import numpy as np
import pytest

import approximatequeryengine_b200 as aqe
from conftest import golden_files, load_golden
from oracle import make_params as orc_params

NO_DATA = ["slow_pointer", "fast_pointer", "dual_pointer", "parallel_pointer", "random_pointer", "memory_stride",
           "optimized_address_arithmetic", "index_based", "byte_offset", "optimized_clt", "block", "page",
           "parallel_block", "node_skip", "balanced_tree", "direct_access", "stratified_block",
           "sample_records", "optimized_sequential", "random_start_nth", "address_arithmetic",
           "random_start_memory_stride", "multithreaded_memory_stride", "signal_based_clt"]


def sha(idx):
    return hashlib.sha256(np.ascontiguousarray(idx, dtype="<i8").tobytes()).hexdigest()


@pytest.mark.parametrize("path", [f for f in golden_files()], ids=os.path.basename)
def test_plans_match_reference_golden(path):
    g = load_golden(path)
    for v in g["samplers"]:
        if v["method"] not in NO_DATA:
            continue
        if v["method"] == "stratified_block":
            continue  # positions index the amount order; checked against data in the GPU tests
        pl = aqe.build_plan(g["n"], v["method"], aqe.make_params(v["method"], v["percent"], **v["kw"]))
        tag = (v["method"], v["percent"], v["kw"])
        assert pl.count == v["count"], tag
        assert sha(pl.indices()) == v["idx_sha256"], tag


@pytest.mark.parametrize("n", [0, 1, 7, 254, 255, 256, 381, 382, 999, 1000, 1001, 4999, 5000, 12345, 65537, 99999, 250007])
def test_plans_match_oracle_sweep(oracle, n):
    variants = [{}, {"num_threads": 3, "step_size": 3, "block_size": 128, "block_size_max": 5, "seed": 7, "check_interval": 4},
                {"num_threads": 7, "step_size": 1, "block_size": 33, "block_size_max": 3, "seed": 123456789}]
    for m in NO_DATA:
        for p in [0.0, 0.01, 0.5, 1.0, 3.3, 10.0, 25.0, 50.0, 99.9, 100.0]:
            for kw in variants:
                kw = dict(kw)
                if m == "page" and kw:
                    kw["block_size"] *= 32
                T = int(n * p / 100.0)
                if m == "dual_pointer" and 0 < T < 3:
                    with pytest.raises(aqe.AqeError):
                        aqe.build_plan(n, m, aqe.make_params(m, p, **kw))
                    continue
                want = oracle.indices(None, m, orc_params(m, p, **kw), n_rows=n)
                pl = aqe.build_plan(n, m, aqe.make_params(m, p, **kw))
                got = pl.indices()
                assert len(got) == len(want) and np.array_equal(got, want), (n, m, p, kw, got[:5], want[:5])
                if pl.count:
                    assert got.min() >= 0 and got.max() < n


def test_segment_plans_stay_small():
    """The affine samplers never materialise an index list: a handful of segments whatever N is."""
    n = 1_000_000_000
    for m, limit in [("memory_stride", 1), ("slow_pointer", 1), ("block", 2), ("parallel_block", 8), ("parallel_pointer", 4),
                     ("optimized_clt", 4), ("index_based", 1), ("node_skip", 2), ("multithreaded_memory_stride", 4),
                     ("sample_records", 1), ("address_arithmetic", 1), ("random_start_nth", 2), ("signal_based_clt", 2)]:
        pl = aqe.build_plan(n, m, aqe.make_params(m, 1.0))
        assert 1 <= pl.num_segments <= limit, (m, pl.num_segments)
        assert pl.count > 0


def test_sample_records_is_srswor():
    """sample_records (custom_bplus_db.cpp:345-363) = k distinct rows, uniformly scattered: the prefix of a seeded permutation."""
    n = 1_000_003
    a = aqe.build_plan(n, "sample_records", aqe.make_params("sample_records", 10.0, seed=1)).indices()
    b = aqe.build_plan(n, "sample_records", aqe.make_params("sample_records", 10.0, seed=2)).indices()
    assert len(a) == n // 10 and len(np.unique(a)) == len(a) and a.min() >= 0 and a.max() < n
    assert len(np.intersect1d(a, b)) < 0.12 * len(a)                 # independent seeds overlap ~10 %
    hist = np.bincount(a * 64 // n, minlength=64)                     # uniform over the table: chi-square, 63 dof
    exp = len(a) / 64
    assert ((hist - exp) ** 2 / exp).sum() < 120
    lag = np.corrcoef(a[:-1].astype(float), a[1:].astype(float))[0, 1]
    assert abs(lag) < 0.01                                             # consecutive draws are not correlated
    for small in (1, 2, 3, 5, 16, 17, 1000):                           # a full prefix of the permutation is a bijection
        pl = aqe.build_plan(small, "sample_records", aqe.make_params("sample_records", 99.9999999, seed=9))
        full = [aqe.lib().aqe_plan_count(pl.h)]
        idx = pl.indices()
        assert len(np.unique(idx)) == len(idx) and (len(idx) == 0 or idx.max() < small)
    big = aqe.build_plan(10**9, "sample_records", aqe.make_params("sample_records", 10.0, seed=3))
    assert big.num_segments == 1 and big.count == 10**8              # no index list at any size


def test_invalid_arguments_raise_instead_of_ub():
    for m, kw in [("fast_pointer", {"step_size": 0}), ("parallel_pointer", {"num_threads": 0}), ("node_skip", {"step_size": 0}),
                  ("block", {"block_size": 0}), ("parallel_block", {"num_threads": 0}), ("random_start_nth", {"step_size": 0})]:
        with pytest.raises(aqe.AqeError) as ei:
            aqe.build_plan(10000, m, aqe.make_params(m, 5.0, **kw))
        assert ei.value.code == 1


def test_tree_shape_closed_form(oracle):
    L = aqe.lib()
    for g in golden_files():
        gd = load_golden(g)
        e = aqe.Engine.__new__(aqe.Engine)  # no device needed: count-only handle
    for n in [0, 1, 254, 255, 381, 382, 32767, 32768, 100000, 1000000]:
        assert oracle.tree_height(n) >= 1


def test_host_generator_matches_oracle(oracle):
    a = aqe.synth_rows_host(5000, seed=7)
    b = oracle.synth(5000, seed=7)
    assert a.tobytes() == b.tobytes()
    a = aqe.synth_rows_host(100, seed=99, first_row=123456789012)
    b = oracle.synth(100, seed=99, first_row=123456789012)
    assert a.tobytes() == b.tobytes()


def test_merge_partials_fixed_order():
    import ctypes as C
    parts = []
    for i in range(8):
        p = aqe.Partial(count=10 + i, sum=1e15 + i * 0.1, comp=1e-3 * i, isum_lo=(1 << 63) + i, isum_hi=i - 3, sumsq=2.0 * i, minv=float(i), maxv=float(100 - i))
        parts.append(p)
    m = aqe.merge_partials(parts)
    assert m.count == sum(p.count for p in parts)
    from fractions import Fraction
    exact = sum(Fraction(p.sum) + Fraction(p.comp) for p in parts)
    assert abs(Fraction(m.sum) + Fraction(m.comp) - exact) <= abs(exact) * Fraction(1, 10**25)
    assert m.minv == 0.0 and m.maxv == 100.0
    mi = aqe.merge_partials(parts, is_integer=True)
    assert mi.isum == sum(p.isum for p in parts)


def test_estimator_matches_cli_formulas(oracle):
    """aqe_estimate == enhanced_aqe_cli.py:190-195, 276-291 given the same moments."""
    rows = oracle.synth(20000, seed=3)
    idx = np.arange(0, 20000, 37)
    s = oracle.stats(rows, idx)
    st = aqe.Stats(n=s.n, mean=s.mean, m2=s.m2, sum=s.sum)
    for agg in ("sum", "avg", "count"):
        for legacy in (False, True):
            assert aqe.estimate(st, 20000, agg, 1.96, legacy) == oracle.estimate(s, 20000, agg, 1.96, legacy)
    assert aqe.lib().aqe_z_score(0.95, 0) == 1.96 and aqe.lib().aqe_z_score(0.99, 0) == 2.576 and aqe.lib().aqe_z_score(0.9, 0) == 1.645
    assert abs(aqe.lib().aqe_z_score(0.95, 1) - 1.959963984540054) < 1e-8
    assert aqe.lib().aqe_z_score(0.95, 1) == oracle.z_score(0.95, True)
