"""GPU parity tests (run with `-m gpu` on a B200).  Every call goes through the C-ABI of libaqe_b200.so
(ctypes) or through the drop-in `aqe_backend` module on top of it; the oracle (oracle/aqe_oracle.c, pinned to
the compiled reference) and the golden vectors minted from the reference are the checkers.

Stated tolerances (north_star): integer / index / byte results bit-exact; fp64 sums relative <= 1e-12 against
the reference's serial sum (and <= 4 ulp against the exactly rounded sum); sampled estimates on the same
index list relative <= 1e-12, second moments relative <= 1e-9.
"""
import hashlib
import math
import os

import numpy as np
import pytest

import approximatequeryengine_b200 as aqe
from conftest import fhex, golden_files, load_golden
from oracle import ApproxSpec, make_params as orc_params

pytestmark = pytest.mark.gpu

REL = 1e-12
FILES = golden_files()


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def sha(idx):
    return hashlib.sha256(np.ascontiguousarray(idx, dtype="<i8").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tables(oracle):
    """(golden, rows, engine) per golden file; engines stay resident for the module."""
    out = []
    for path in FILES:
        g = load_golden(path)
        rows = oracle.synth(g["n"], seed=g["seed"])
        out.append((g, rows, aqe.Engine(0).from_rows(rows)))
    return out


def test_device_present():
    n = aqe.C.c_int()
    aqe.check(aqe.lib().aqe_device_count(aqe.C.byref(n)))
    assert n.value >= 1


# ---------------------------------------------------------------------------------------------------------
# exact aggregates vs the reference's own full scan
# ---------------------------------------------------------------------------------------------------------
def test_exact_sum_against_reference_golden(tables, oracle):
    for g, rows, e in tables:
        assert e.count == g["total_records"]
        s = e.sum_amount()
        assert rel(s, fhex(g["sum_amount"])) <= REL
        exact = math.fsum(rows["amount"])
        assert abs(s - exact) <= 4 * math.ulp(exact)      # compensated accumulation: ~exactly rounded
        assert e.sum_amount() == s                          # run-to-run bit stable
        for key, v in g["sum_amount_where"].items():
            lo, hi = (float(x) for x in key.split(","))
            w, c = e.sum_amount_where(lo, hi)
            assert c == v["count"]
            assert rel(w, fhex(v["sum"])) <= REL if v["count"] else w == 0.0
        for col, want in g["int_sums"].items():
            assert e.sum_int(col) == want                   # bit exact (int128)


def test_scan_generic_predicates(tables, oracle):
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    t0 = 1700000000
    cases = [("amount", None, 0, 0), ("amount", "amount", 250.0, 750.0), ("amount", "timestamp", t0 + 5000, t0 + 70000),
             ("amount", "region", 2, 5), ("amount", "id", 10, 10), ("id", None, 0, 0), ("id", "amount", 1.0, 300.0),
             ("timestamp", "region", 0, 0), ("timestamp", "timestamp", t0 + 1, t0 + 3), ("region", None, 0, 0),
             ("product_id", "product_id", 100, 899), ("product_id", "amount", 999.0, 1000.0), ("region", "id", 50000, 10**12),
             ("amount", "amount", 5000.0, 6000.0)]
    for agg, pred, lo, hi in cases:
        p = e.scan(agg, pred, lo, hi)
        q = oracle.scan(rows, agg, pred, lo, hi)
        assert p.count == q.count, (agg, pred)
        if agg == "amount":
            assert (rel(p.sum, q.sum) <= REL) if q.count else p.sum == 0.0
            if q.count:
                assert rel(p.sumsq, q.sumsq) <= 1e-11 and p.minv == q.minv and p.maxv == q.maxv
        else:
            assert p.isum == q.isum, (agg, pred)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 31, 32, 33, 255, 1023, 1025, 4099])
def test_ragged_sizes(oracle, n):
    rows = oracle.synth(n, seed=5)
    e = aqe.Engine(0).from_rows(rows)
    assert e.count == n
    if n == 0:
        assert e.sum_amount() == 0.0 and e.sum_amount_where(0, 1e9) == (0.0, 0)
        return
    assert rel(e.sum_amount(), oracle.sum_amount(rows)) <= REL
    w, c = e.sum_amount_where(100.0, 500.0)
    wo, co = oracle.sum_amount_where(rows, 100.0, 500.0)
    assert c == co and (rel(w, wo) <= REL if co else w == 0.0)
    assert e.sum_int("id") == n * (n + 1) // 2
    assert e.read_rows().tobytes() == rows.tobytes()


def test_unaligned_attached_columns(oracle):
    """Torch hand-off: columns borrowed from another allocation at an odd element offset take the scalar path."""
    import torch
    rows = oracle.synth(50001, seed=9)
    amt = torch.from_numpy(rows["amount"].copy()).cuda()
    ids = torch.from_numpy(rows["id"].copy()).cuda()
    reg = torch.from_numpy(rows["region"].copy()).cuda()
    for off in (0, 1, 3):
        n = len(rows) - off
        e = aqe.Engine(0).attach(n, id=ids[off:].data_ptr(), amount=amt[off:].data_ptr(), region=reg[off:].data_ptr())
        sub = rows[off:]
        assert rel(e.sum_amount(), oracle.sum_amount(sub)) <= REL
        p = e.scan("amount", "region", 1, 3); q = oracle.scan(sub, "amount", "region", 1, 3)
        assert p.count == q.count and rel(p.sum, q.sum) <= REL
        assert e.sum_int("id") == int(sub["id"].astype(object).sum())
        e.close()


def test_dropin_from_torch(oracle):
    import torch
    b = aqe.backend()
    rows = oracle.synth(77777, seed=12)
    amt = torch.from_numpy(rows["amount"].copy()).cuda(); ts = torch.from_numpy(rows["timestamp"].copy()).cuda()
    db = b.CustomBPlusDB()
    db.from_torch(amount=amt, timestamp=ts)
    assert db.get_total_records() == len(rows) and rel(db.sum_amount(), oracle.sum_amount(rows)) <= REL
    assert db.sum_column("timestamp") == int(rows["timestamp"].astype(object).sum())
    d = db.scan("amount", "timestamp", 1700000000 + 10, 1700000000 + 50000)
    q = oracle.scan(rows, "amount", "timestamp", 1700000000 + 10, 1700000000 + 50000)
    assert d["count"] == q.count and rel(d["sum"], q.sum) <= REL
    with pytest.raises(ValueError):
        db.from_torch(amount=amt.float())
    with pytest.raises(ValueError):
        db.from_torch(amount=amt.cpu())
    with pytest.raises(RuntimeError):
        db.sum_column("id")                                  # column was not handed over


def test_special_values(oracle):
    rows = oracle.synth(10000, seed=1)
    rows["amount"][::7] = -rows["amount"][::7]
    rows["amount"][5] = float("nan"); rows["amount"][6] = float("inf"); rows["amount"][8] = 0.0
    rows["id"][::3] *= -1
    rows["timestamp"][10] = -(2 ** 62); rows["timestamp"][11] = 2 ** 62
    e = aqe.Engine(0).from_rows(rows)     # ids now unsorted -> engine orders by id (load_from_file semantics)
    srt = rows[np.argsort(rows["id"], kind="stable")]
    assert e.read_rows().tobytes() == srt.tobytes()
    w, c = e.sum_amount_where(-1e6, 1e6)   # NaN and inf fail the closed interval
    wo, co = oracle.sum_amount_where(srt, -1e6, 1e6)
    assert c == co and rel(w, wo) <= 1e-11
    assert math.isnan(e.sum_amount())
    for col in ("id", "timestamp"):
        assert e.sum_int(col) == int(srt[col].astype(object).sum())


def test_duplicate_ids_and_all_equal_values(oracle):
    """Collisions: duplicate ids are allowed by the format (lower_bound insert, custom_bplus_db.cpp:32-37); aggregates do not
    care, rows stay ordered by id, and a constant column has zero variance."""
    n = 30000
    rows = oracle.synth(n, seed=6)
    rows["id"][:] = np.repeat(np.arange(1, n // 3 + 1), 3)            # every id three times
    rows["amount"][:] = 123.25
    e = aqe.Engine(0).from_rows(rows[np.random.default_rng(3).permutation(n)])
    back = e.read_rows()
    assert np.array_equal(back["id"], rows["id"]) and e.count == n
    assert e.sum_amount() == 123.25 * n and e.sum_amount_where(123.25, 123.25) == (123.25 * n, n)
    assert e.sum_int("id") == int(rows["id"].astype(object).sum())
    st = e.stats(e.plan("memory_stride", aqe.make_params("memory_stride", 10.0)))
    assert st.mean == 123.25 and st.m2 == 0.0 and st.sum == 123.25 * st.n
    a = e.approx("avg", error_percent=1.0, seed=2)                     # zero variance: converges at the first look
    assert a.estimate == 123.25 and a.ci_lower == a.ci_upper == 123.25 and a.status == 0 and a.rounds == 1


def test_int128_overflow_range(oracle):
    n = 70000
    rows = oracle.synth(n, seed=2)
    rows["timestamp"][:] = 2 ** 62 + np.arange(n)        # sum needs > 64 bits
    e = aqe.Engine(0).from_rows(rows)
    assert e.sum_int("timestamp") == int(rows["timestamp"].astype(object).sum()) > 2 ** 64
    rows["timestamp"][:] = -(2 ** 62) - np.arange(n)
    e = aqe.Engine(0).from_rows(rows)
    assert e.sum_int("timestamp") == int(rows["timestamp"].astype(object).sum()) < -(2 ** 64)


def test_integer_predicate_bounds(oracle):
    """Integer predicate columns are compared as integers against [ilo, ihi]; that must select exactly the rows the
    reference-shaped `lo <= (double)v <= hi` selects, including fractional, huge (> 2^53), negative and empty bounds."""
    n = 40000
    rows = oracle.synth(n, seed=4)
    big = 2 ** 53
    rows["timestamp"][:] = np.concatenate([np.arange(-5, 5), big + np.arange(-20, 20), -big + np.arange(-20, 20),
                                           [2 ** 63 - 1, -(2 ** 63), 2 ** 62, -(2 ** 62)],
                                           np.random.default_rng(1).integers(-(2 ** 63), 2 ** 63 - 1, size=n - 94)])
    rows["region"][:] = np.random.default_rng(2).integers(-(2 ** 31), 2 ** 31 - 1, size=n)
    rows["region"][:6] = [-(2 ** 31), 2 ** 31 - 1, 0, -1, 1, 7]
    e = aqe.Engine(0).from_rows(rows)
    inf = float("inf")
    bounds = [(-0.5, 0.5), (0.0, 0.0), (-3.999, 3.001), (float(big) - 2, float(big) + 2), (float(big), float(big) + 6), (-float(big) - 4, -float(big) + 4),
              (9.2e18, inf), (-inf, -9.2e18), (2.0 ** 63, inf), (-inf, inf), (5.0, 4.0), (float("nan"), 1.0), (0.0, float("nan")),
              (-2.0 ** 31, 2.0 ** 31), (2.0 ** 31 - 1, 2.0 ** 31 - 1), (-2147483648.5, -2147483647.5), (1e-300, 6.9999999), (9.223372036854775e18, 9.223372036854776e18)]
    for lo, hi in bounds:
        for agg, pred in (("amount", "timestamp"), ("id", "timestamp"), ("timestamp", "timestamp"), ("amount", "region"), ("region", "region"),
                          ("product_id", "region"), ("region", "timestamp")):
            p = e.scan(agg, pred, lo, hi)
            q = oracle.scan(rows, agg, pred, lo, hi)
            assert p.count == q.count, (agg, pred, lo, hi, p.count, q.count)
            if agg == "amount":
                assert rel(p.sum, q.sum) <= REL if q.count else p.sum == 0.0
            else:
                assert p.isum == q.isum, (agg, pred, lo, hi)


# ---------------------------------------------------------------------------------------------------------
# file format, generator
# ---------------------------------------------------------------------------------------------------------
def test_file_round_trip(oracle, tmp_path):
    rows = oracle.synth(123457, seed=21)
    p1, p2 = str(tmp_path / "a.aqe"), str(tmp_path / "b.aqe")
    oracle.save_file(p1, rows)
    e = aqe.Engine(0).load_file(p1)
    assert e.count == len(rows) and e.read_rows().tobytes() == rows.tobytes()
    e.save_file(p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()          # byte-identical to save_to_file's layout
    shard = aqe.Engine(0).load_file(p1, first_row=1000, n_rows=5000)
    assert shard.read_rows().tobytes() == rows[1000:6000].tobytes()
    # unsorted file: rows come back ordered by id (insert_batch, custom_bplus_db.cpp:198-200)
    perm = np.random.default_rng(0).permutation(len(rows))
    oracle.save_file(p1, rows[perm])
    assert aqe.Engine(0).load_file(p1).read_rows().tobytes() == rows.tobytes()
    with pytest.raises(aqe.AqeError):
        aqe.Engine(0).load_file(str(tmp_path / "missing.aqe"))


def test_device_generator_bit_identical(oracle):
    for n, seed, first in [(100003, 7, 0), (4097, 123, 10**12)]:
        e = aqe.Engine(0).generate(n, seed=seed, first_row=first)
        assert e.read_rows().tobytes() == oracle.synth(n, seed=seed, first_row=first).tobytes()
    ln = aqe.Engine(0).generate(200000, seed=7, dist=1).read_rows()
    host = oracle.synth(200000, seed=7, dist=1)
    assert np.allclose(ln["amount"], host["amount"], rtol=1e-12)    # libdevice vs libm: not bit-identical
    assert np.array_equal(ln["region"], host["region"])


# ---------------------------------------------------------------------------------------------------------
# samplers: identical rows to the reference, estimator parity on the same index list
# ---------------------------------------------------------------------------------------------------------
def test_samplers_against_reference_golden(tables, oracle):
    for g, rows, e in tables:
        if g["n"] > 100000:
            continue
        order = None
        for v in g["samplers"]:
            prm = aqe.make_params(v["method"], v["percent"], **v["kw"])
            pl = e.plan(v["method"], prm)
            got = e.gather(pl)
            tag = (g["n"], v["method"], v["percent"], v["kw"])
            assert len(got) == v["count"], tag
            assert sha(got["id"] - 1) == v["idx_sha256"], tag
            if v["est"] is None or "m2" not in v["est"]:
                continue
            idx = got["id"] - 1
            assert np.array_equal(got["amount"], rows["amount"][idx])
            s = e.stats(pl)
            assert s.n == v["count"]
            assert rel(s.sum, fhex(v["est"]["sample_sum"])) <= REL, tag
            est, lo, hi = aqe.estimate(s, g["n"], "sum", 1.96, legacy_ci=True)
            assert rel(est, fhex(v["est"]["sum"])) <= REL, tag
            avg, alo, ahi = aqe.estimate(s, g["n"], "avg", 1.96)
            assert rel(avg, fhex(v["est"]["avg"])) <= REL, tag
            if fhex(v["est"]["m2"]) > 0:
                assert rel(s.m2, fhex(v["est"]["m2"])) <= 1e-9, tag
                assert rel(ahi - alo, fhex(v["est"]["avg_ci"][1]) - fhex(v["est"]["avg_ci"][0])) <= 1e-9, tag


def test_samplers_1m_all_reference_vectors(tables):
    """All 144 sampler vectors minted from the reference at 1 M rows: identical rows, estimator parity on the stats path."""
    big = [t for t in tables if t[0]["n"] == 1000000]
    if not big:
        pytest.skip("1M golden not present")
    g, rows, e = big[0]
    for v in g["samplers"]:
        pl = e.plan(v["method"], aqe.make_params(v["method"], v["percent"], **v["kw"]))
        got = e.gather(pl)
        tag = (v["method"], v["percent"], v["kw"])
        assert len(got) == v["count"] and sha(got["id"] - 1) == v["idx_sha256"], tag
        if v["est"] and "m2" in v["est"]:
            s = e.stats(pl)
            assert rel(s.sum, fhex(v["est"]["sample_sum"])) <= REL, tag
            assert rel(aqe.estimate(s, g["n"], "sum", 1.96, True)[0], fhex(v["est"]["sum"])) <= REL, tag
            if fhex(v["est"]["m2"]) > 0:
                assert rel(s.m2, fhex(v["est"]["m2"])) <= 1e-9, tag


def test_same_index_list_estimates(tables, oracle):
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    rng = np.random.default_rng(4)
    for n in (1, 2, 50, 4096, 60000):
        idx = rng.integers(0, g["n"], size=n)
        s = e.stats_from_indices(idx)
        o = oracle.stats(rows, idx)
        assert s.n == o.n and rel(s.sum, o.sum) <= REL and rel(s.mean, o.mean) <= REL
        if n > 1:
            assert rel(s.m2, o.m2) <= 1e-9
        assert np.array_equal(e.gather_indices(idx)["id"], rows["id"][idx])
        for col in ("timestamp", "region", "id"):
            sc = e.stats_from_indices(idx, col)
            oc = oracle.stats(rows, idx, col)
            assert rel(sc.sum, oc.sum) <= REL and (n == 1 or abs(sc.m2 - oc.m2) <= 1e-9 * max(oc.m2, 1.0))
    with pytest.raises(aqe.AqeError):
        e.stats_from_indices([0, g["n"]])


def test_seeded_and_lockstep_samplers_match_oracle(tables, oracle):
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    for m in ("sample_records", "optimized_sequential", "random_start_nth", "address_arithmetic", "random_start_memory_stride",
              "multithreaded_memory_stride", "signal_based_clt"):
        for p in (1.0, 12.5):
            for seed in (1, 99):
                got = e.gather(e.plan(m, aqe.make_params(m, p, seed=seed)))
                want = oracle.indices(rows, m, orc_params(m, p, seed=seed))
                assert np.array_equal(got["id"] - 1, want), (m, p, seed)
    for p, ci, th, err, conf in [(20.0, 10, 4, 1.0, 0.95), (10.0, 10, 4, 5.0, 0.95), (20.0, 10, 4, 0.5, 0.95), (15.0, 20, 6, 2.0, 0.99),
                                 (10.0, 4, 2, 3.0, 0.90), (5.0, 10, 5, 0.05, 0.95)]:
        kw = dict(check_interval=ci, num_threads=th, max_error_percent=err, confidence_level=conf)
        got = e.gather(e.plan("clt_validated_dual_pointer", aqe.make_params("clt_validated_dual_pointer", p, **kw)))
        want = oracle.indices(rows, "clt_validated_dual_pointer", orc_params("clt_validated_dual_pointer", p, **kw))
        assert np.array_equal(got["id"] - 1, want), (p, kw, len(got), len(want))
    s, n = e.fast_aggregated(aqe.make_params("multithreaded_memory_stride", 2.0, seed=5))
    so, no = oracle.fast_aggregated(rows, orc_params("multithreaded_memory_stride", 2.0, seed=5))
    assert n == no and rel(s, so) <= REL


# ---------------------------------------------------------------------------------------------------------
# persistent CLT kernel
# ---------------------------------------------------------------------------------------------------------
def ospec(agg="sum", design="srs", where=None, where_col="amount", eps=1.0, conf=0.95, seed=0, min_samples=0, max_samples=0, block=0):
    return ApproxSpec(agg=aqe.AGG[agg], design=aqe.DESIGN[design], agg_col=1, pred_col=aqe.COLS[where_col] if where else -1,
                      lo=where[0] if where else 0.0, hi=where[1] if where else 0.0, error_percent=eps, confidence_level=conf,
                      seed=seed, min_samples=min_samples, max_samples=max_samples, block_size=block)


def test_approx_matches_restated_kernel(tables, oracle):
    g, rows, e = [t for t in tables if t[0]["n"] == 1000000][0] if any(t[0]["n"] == 1000000 for t in tables) else tables[-1]
    cases = [dict(agg="sum"), dict(agg="avg"), dict(agg="sum", eps=0.25), dict(agg="avg", eps=5.0, min_samples=64),
             dict(agg="sum", where=(100.0, 500.0)), dict(agg="count", where=(100.0, 500.0)), dict(agg="avg", where=(100.0, 500.0)),
             dict(agg="sum", design="block", min_samples=64), dict(agg="avg", design="block", min_samples=32, block=512, eps=2.0),
             dict(agg="sum", eps=0.01, max_samples=200000), dict(agg="count"), dict(agg="sum", where=(3.0, 2.0), max_samples=50000)]
    for kw in cases:
        for seed in (0, 7):
            a = e.approx(seed=seed, error_percent=kw.get("eps", 1.0), confidence_level=kw.get("conf", 0.95), agg=kw["agg"],
                         design=kw.get("design", "srs"), where=kw.get("where"), min_samples=kw.get("min_samples", 0),
                         max_samples=kw.get("max_samples", 0), block_size=kw.get("block", 0))
            o = oracle.approx(rows, ospec(seed=seed, **kw))
            assert (a.n_units, a.n_samples, a.rounds, a.status) == (o.n_units, o.n_samples, o.rounds, o.status), (kw, seed)
            if o.estimate != 0:
                assert rel(a.estimate, o.estimate) <= 1e-10, (kw, seed)
                assert rel(a.ci_upper - a.ci_lower, o.ci_upper - o.ci_lower) <= 1e-8, (kw, seed)
            else:
                assert a.estimate == 0


def test_approx_small_tables_are_scanned_exactly(oracle):
    for n in (1, 100, 16384, 16385):
        rows = oracle.synth(n, seed=3)
        e = aqe.Engine(0).from_rows(rows)
        for agg, where in (("sum", None), ("avg", None), ("sum", (100.0, 500.0)), ("avg", (100.0, 500.0)), ("count", (100.0, 500.0)), ("count", None)):
            a = e.approx(agg, error_percent=1.0, seed=1, where=where)
            o = oracle.approx(rows, ospec(agg=agg, where=where, seed=1))
            assert (a.n_units, a.n_samples, a.status, a.rounds) == (o.n_units, o.n_samples, o.status, o.rounds), (n, agg, where)
            assert rel(a.estimate, o.estimate) <= 1e-10 if o.estimate else a.estimate == 0.0
            if n <= 16384:
                assert a.ci_lower == a.estimate == a.ci_upper


def test_approx_ci_coverage(tables, oracle):
    """CI coverage on repeated seeds >= nominal (minus 2.5 binomial sigma of the finite seed count)."""
    g, rows, e = [t for t in tables if t[0]["n"] == 1000000][0] if any(t[0]["n"] == 1000000 for t in tables) else tables[-1]
    truth_sum = math.fsum(rows["amount"])
    seeds = 400
    for agg, truth in (("sum", truth_sum), ("avg", truth_sum / g["n"])):
        for eps in (0.5, 1.0, 5.0):
            hit = 0
            for seed in range(seeds):
                a = e.approx(agg, error_percent=eps, seed=1000 + seed)
                hit += a.ci_lower <= truth <= a.ci_upper
                assert a.status == 0 and a.error_margin * 100 <= eps + 1e-9
            cov = hit / seeds
            assert cov >= 0.95 - 2.5 * math.sqrt(0.95 * 0.05 / seeds), (agg, eps, cov)


# ---------------------------------------------------------------------------------------------------------
# drop-in module: the calls enhanced_aqe_cli.py makes, checked with the CLI's own estimator code
# ---------------------------------------------------------------------------------------------------------
def test_dropin_cli_flows(oracle, tmp_path):
    g = load_golden([f for f in FILES if "n100000_s7" in f][0])
    rows = oracle.synth(g["n"], seed=g["seed"])
    path = str(tmp_path / "sales.aqe")
    oracle.save_file(path, rows)
    b = aqe.backend()
    db = b.CustomBPlusDB()
    assert db.open_database(path) is True                           # reference: deadlocks
    assert db.open_database(str(tmp_path / "nope")) is False
    assert db.load_from_file(path) is True
    N = db.get_total_records()
    assert N == g["n"] and db.get_node_count() == g["node_count"]
    assert rel(db.sum_amount(), fhex(g["sum_amount"])) <= REL       # execute_exact_query, cli:336
    by = {(v["method"], v["percent"]): v for v in g["samplers"] if not v["kw"]}
    # execute_random_sampling (cli:158-225): N > 50000 -> memory_stride_sample(p, 0)
    samples = db.memory_stride_sample(1.0, 0)
    v = by[("memory_stride", 1.0)]
    assert sha(np.array([r.id - 1 for r in samples])) == v["idx_sha256"]
    est = sum(r.amount for r in samples) * (N / len(samples))
    assert est == fhex(v["est"]["sum"])                              # same rows, same Python arithmetic: bit-equal
    assert sum(r.amount for r in samples) / len(samples) == fhex(v["est"]["avg"])
    # method 'block' / 'parallel'
    for name, call in (("block", lambda: db.block_sample(5.0)), ("parallel_block", lambda: db.parallel_block_sample(5.0)),
                       ("parallel_pointer", lambda: db.parallel_pointer_sample(5.0)), ("random_pointer", lambda: db.random_pointer_sample(5.0)),
                       ("direct_access", lambda: db.direct_access_sample(5.0)), ("stratified_block", lambda: db.stratified_block_sample(5.0)),
                       ("adaptive_block", lambda: db.adaptive_block_sample(5.0)), ("page", lambda: db.page_sample(5.0)),
                       ("node_skip", lambda: db.node_skip_sample(5.0)), ("dual_pointer", lambda: db.dual_pointer_sample(5.0))):
        got = call()
        assert sha(np.array([r.id - 1 for r in got])) == by[(name, 5.0)]["idx_sha256"], name
    # execute_clt_approximation (cli:230-315): clt_validated_dual_pointer_sample(20, 0.95, 10, 4, eps)
    clt = db.clt_validated_dual_pointer_sample(20, 0.95, 10, 4, 1.0)
    want = oracle.indices(rows, "clt_validated_dual_pointer", orc_params("clt_validated_dual_pointer", 20.0, max_error_percent=1.0))
    assert np.array_equal(np.array([r.id - 1 for r in clt]), want)
    vals = [r.amount for r in clt]
    mean = sum(vals) / len(vals)
    moe = 1.96 * (sum((x - mean) ** 2 for x in vals) / (len(vals) - 1)) ** 0.5 / len(vals) ** 0.5
    truth = fhex(g["avg_amount"])
    assert abs(mean - truth) / truth < 0.02 and moe / mean * 100 < 2.5
    # fused estimator: same result fields (+ a correct interval)
    r = db.approx_sum(error_percent=1.0, seed=3)
    assert r.status == b.CustomApproximationStatus.STABLE and r.error_margin <= 0.01 and r.samples_used > 0
    assert r.ci_lower < r.value < r.ci_upper and r.confidence_level == 0.95
    assert hasattr(r.computation_time, "total_seconds")
    assert db.sum_column("id") == N * (N + 1) // 2
    # numpy / device-stats forms of the same samplers (no per-row Python objects)
    arr = db.sample_array("memory_stride", 1.0)
    assert arr.dtype == b.CustomBPlusDB.record_dtype and sha(arr["id"] - 1) == by[("memory_stride", 1.0)]["idx_sha256"]
    assert np.array_equal(arr["amount"], rows["amount"][arr["id"] - 1])
    st = db.sample_array("memory_stride", 1.0, stats=True)
    assert st["n"] == len(arr) and rel(st["sum"], fhex(by[("memory_stride", 1.0)]["est"]["sample_sum"])) <= REL
    assert len(db.sample_array("block", 5.0, block_size=1000)) == by[("block", 5.0)]["count"]
    assert len(db.sample_array("slow_pointer", 0.0)) == 0
    with pytest.raises(ValueError):
        db.sample_array("no_such_sampler", 1.0)
    db.close_database()


def test_dropin_create_insert_flow(oracle, tmp_path):
    """SURVEY Appendix D.1 flow: create_database + insert_record per row (unsorted), exact + save + reload."""
    b = aqe.backend()
    rows = oracle.synth(3000, seed=13)
    perm = np.random.default_rng(1).permutation(len(rows))
    path = str(tmp_path / "created.aqe")
    db = b.CustomBPlusDB()
    assert db.create_database(path)
    for i in perm:
        r = b.Record()
        r.id, r.amount, r.region, r.product_id, r.timestamp = (int(rows["id"][i]), float(rows["amount"][i]), int(rows["region"][i]),
                                                                int(rows["product_id"][i]), int(rows["timestamp"][i]))
        assert db.insert_record(r)
    assert db.get_total_records() == 3000
    assert rel(db.sum_amount(), oracle.sum_amount(rows)) <= REL
    assert [r.id for r in db.slow_pointer_sample(10.0)] == list(oracle.indices(rows, "slow_pointer", orc_params("slow_pointer", 10.0)) + 1)
    del db                                                           # destructor -> close_database -> save (cbd:131-133,157-162)
    import gc; gc.collect()
    assert open(path, "rb").read()[24:] == rows.tobytes()
    s = b.CustomApproximateScheduler(0.05)
    assert s.open_database(path) and s.get_total_records() == 3000
    ex = s.execute_exact_sum()
    assert rel(ex.value, oracle.sum_amount(rows)) <= REL and ex.confidence_level == 1.0 and ex.error_margin == 0.0 and ex.samples_used == 3000
    assert s.execute_exact_count().value == 3000.0
    q = s.execute_sum_query("SELECT SUM(amount) FROM sales", 10.0, 4)
    assert q.status == b.CustomApproximationStatus.STABLE and q.samples_used == 300 and q.error_margin == 0.1 and q.confidence_level == 0.85
    assert abs(q.value - ex.value) / ex.value < 0.2
    w = s.execute_sum_query("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 50.0, 4)
    assert abs(w.value - oracle.sum_amount_where(rows, 100, 500)[0]) / w.value < 0.2
    assert abs(s.get_database_size_mb() - 3000 * 32 / 1048576) < 1e-12
    bm = s.benchmark_query("SUM", 10.0, 4)
    assert bm.exact_value == ex.value and bm.threads_used == 4


def test_scheduler_fields_match_reference_golden(oracle, tmp_path):
    g = load_golden([f for f in FILES if "n100000_s42" in f][0])
    rows = oracle.synth(g["n"], seed=g["seed"])
    path = str(tmp_path / "s.aqe"); oracle.save_file(path, rows)
    b = aqe.backend()
    s = b.CustomApproximateScheduler()
    assert s.open_database(path)
    for name, call in (("exact_sum", s.execute_exact_sum), ("exact_avg", s.execute_exact_avg), ("exact_count", s.execute_exact_count)):
        r, w = call(), g["scheduler"][name]
        assert rel(r.value, fhex(w["value"])) <= REL and int(r.status) == w["status"]
        assert (r.confidence_level, r.error_margin, r.samples_used) == (w["confidence_level"], w["error_margin"], w["samples_used"])
    r, w = s.execute_sum_query("SELECT SUM(amount) FROM sales", 10.0, 4), g["scheduler"]["sum_query_fields"]
    assert (int(r.status), r.confidence_level, r.error_margin, r.samples_used) == (w["status"], w["confidence_level"], w["error_margin"], w["samples_used"])
    assert s.get_database_size_mb() == fhex(g["scheduler"]["size_mb"]) and s.get_tree_height() == g["tree_height"]


# ---------------------------------------------------------------------------------------------------------
# end-to-end host column path, sharding
# ---------------------------------------------------------------------------------------------------------
def test_host_column_scan(oracle):
    rows = oracle.synth(3_000_001, seed=17)
    col = np.ascontiguousarray(rows["amount"])
    os.environ.setdefault("AQE_E2E_CHUNK_MB", "8")                   # several chunks
    p = aqe.host_scan_column(col)
    assert p.count == len(col) and rel(p.sum, oracle.sum_amount(rows)) <= REL
    p = aqe.host_scan_column(col, 100.0, 500.0, use_pred=True)
    wo, co = oracle.sum_amount_where(rows, 100.0, 500.0)
    assert p.count == co and rel(p.sum, wo) <= REL
    ts = np.ascontiguousarray(rows["timestamp"])
    assert aqe.host_scan_column(ts).isum == int(ts.astype(object).sum())


def test_host_column_scan_over_several_devices(oracle):
    """aqe_scan_host_column_multi: the chunks of one host column handed out to the GPUs of this process from one counter.  The
    merged partial does not depend on which device took which chunk: bit-equal to the one-device call, every time."""
    import ctypes as C
    c = C.c_int(0)
    aqe.lib().aqe_device_count(C.byref(c))
    devices = list(range(min(c.value, 8)))
    rows = oracle.synth(3_000_001, seed=23)
    col = np.ascontiguousarray(rows["amount"])
    os.environ.setdefault("AQE_E2E_CHUNK_MB", "8")
    one = aqe.host_scan_column(col, 100.0, 500.0, use_pred=True)
    wo, co = oracle.sum_amount_where(rows, 100.0, 500.0)
    assert one.count == co and rel(one.sum, wo) <= REL
    for _ in range(5):
        many = aqe.host_scan_column(col, 100.0, 500.0, use_pred=True, device=devices)
        assert many.count == one.count and many.sum == one.sum
    ts = np.ascontiguousarray(rows["timestamp"])
    assert aqe.host_scan_column(ts, device=devices).isum == int(ts.astype(object).sum())
    assert aqe.host_scan_column(col[:0], device=devices).count == 0
    with pytest.raises(aqe.AqeError):
        aqe.host_scan_column(col, device=[0, 0])


def test_shard_merge_equals_whole(oracle):
    """Contiguous shards merged in rank order give the single-GPU answer."""
    n = 1_000_003
    whole = aqe.Engine(0).generate(n, seed=7)
    ref = whole.scan("amount", "amount", 100.0, 500.0)
    for G in (2, 3, 8):
        parts, iparts = [], []
        for r in range(G):
            a, b = n * r // G, n * (r + 1) // G
            sh = aqe.Engine(0).generate(b - a, seed=7, first_row=a)
            parts.append(sh.scan("amount", "amount", 100.0, 500.0))
            iparts.append(sh.scan("id"))
        m = aqe.merge_partials(parts)
        assert m.count == ref.count and abs(m.sum - ref.sum) <= math.ulp(ref.sum)
        assert aqe.merge_partials(iparts, is_integer=True).isum == n * (n + 1) // 2


# ---------------------------------------------------------------------------------------------------------
# BASELINE full sizes: size-independent properties (known answers, linearity, idempotence)
# ---------------------------------------------------------------------------------------------------------
def test_full_size_properties():
    n = int(os.environ.get("AQE_TEST_FULL_N", 1_000_000_000))
    e = aqe.Engine(0).generate(n, seed=7, columns=("id", "amount", "timestamp"))
    assert e.sum_int("id") == n * (n + 1) // 2                                     # closed forms, bit exact
    assert e.sum_int("timestamp") == 1700000000 * n + n * (n - 1) // 2
    tot = e.scan("amount")
    assert tot.count == n and e.scan("amount").sum == tot.sum                       # idempotent / bit stable
    assert abs(tot.sum / n - 500.5) < 6 * 288.4 / math.sqrt(n)                      # U(1,1000): mean 500.5
    a = e.scan("amount", "amount", 1.0, 400.0); b = e.scan("amount", "amount", math.nextafter(400.0, 1e9), 1000.0)
    assert a.count + b.count == n and abs((a.sum + b.sum) - tot.sum) <= 2 * math.ulp(tot.sum)   # linearity over a partition
    t0 = 1700000000
    h1 = e.scan("amount", "timestamp", t0, t0 + n // 2 - 1); h2 = e.scan("amount", "timestamp", t0 + n // 2, t0 + n)
    assert h1.count == n // 2 and h1.count + h2.count == n and abs((h1.sum + h2.sum) - tot.sum) <= 2 * math.ulp(tot.sum)
    ids = e.scan("id", "amount", 100.0, 500.0)
    assert ids.count == e.scan("amount", "amount", 100.0, 500.0).count
    r = e.approx("sum", error_percent=0.5, seed=1)
    assert r.status == 0 and abs(r.estimate - tot.sum) / tot.sum < 0.01


def test_duplicate_ids_sit_where_the_reference_tree_puts_them(oracle, tmp_path, monkeypatch):
    """Tables with duplicate ids: the engine's row order (reads, samplers, files) is the reference's leaf-chain order after the same
    history of inserts -- golden orders minted from the unmodified reference (tests/golden/make_order_golden.py; csrc/aqe_order.cpp),
    through from_rows, a record file, appended batches / single rows, a sharded handle and the drop-in module."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "order_cases.json")))
    monkeypatch.setenv("AQE_MIN_SHARD_ROWS", "1")
    for c in g["cases"]:
        ids = np.asarray(c["ids"], dtype=np.int64)
        rows = oracle.synth(len(ids), seed=3)
        rows["id"] = ids
        rows["timestamp"] = np.arange(len(ids))            # arrival number: tells rows with equal ids apart
        want = np.asarray(c["order"], dtype=np.int64)
        ops = c["ops"]
        for devs in (None, [0, 0, 0]):
            e = aqe.Engine(0) if devs is None else aqe.Engine(devices=devs)
            at = 0
            for cnt, _ in ops:                               # the history, call by call (one row = insert_record)
                e.append(rows[at:at + cnt]); at += cnt
            back = e.read_rows()
            assert np.array_equal(back["timestamp"], want), (c["name"], devs)
            assert np.array_equal(back["id"], ids[want])
            if len(ops) == 1:
                e2 = (aqe.Engine(0) if devs is None else aqe.Engine(devices=devs)).from_rows(rows)
                assert np.array_equal(e2.read_rows()["timestamp"], want), (c["name"], "from_rows", devs)
                e2.close()
            # a query in between must not disturb the history: the rest of the rows appended after an upload
            if len(ops) > 1:
                e3 = aqe.Engine(0) if devs is None else aqe.Engine(devices=devs)
                e3.append(rows[:ops[0][0]])
                assert e3.count == ops[0][0] and e3.sum_int("timestamp") >= 0      # uploads
                at = ops[0][0]
                for cnt, _ in ops[1:]:
                    e3.append(rows[at:at + cnt]); at += cnt
                assert np.array_equal(e3.read_rows()["timestamp"], want), (c["name"], "query in between", devs)
                e3.close()
            # samplers read the table in that order
            if len(ids) >= 1000:
                prm = aqe.make_params("memory_stride", 10.0)
                got = e.gather(e.plan("memory_stride", prm))
                idx = oracle.indices(rows[want], "memory_stride", orc_params("memory_stride", 10.0))
                assert np.array_equal(got["timestamp"], want[idx]), (c["name"], "memory_stride", devs)
            # the file written from it holds the rows in table order; loading a file is one insert_batch of its rows in file order
            path = str(tmp_path / "dups.aqe")
            e.save_file(path)
            assert np.array_equal(np.fromfile(path, dtype=rows.dtype, offset=24)["timestamp"], want)
            e.close()
        if len(ops) == 1:
            path = str(tmp_path / "arrival.aqe")
            with open(path, "wb") as f:                      # total | tree_height | count | rows in ARRIVAL order (custom_bplus_db.cpp:665-683)
                np.asarray([len(rows), 1, len(rows)], dtype=np.uint64).tofile(f)
                rows.tofile(f)
            e = aqe.Engine(0).load_file(path)
            assert np.array_equal(e.read_rows()["timestamp"], want), (c["name"], "load_file")
            e.close()
