"""SQL-string path on the GPU (run with `-m gpu` on a B200): aqe_sql_* through the C-ABI and run_query* through the
drop-in module, against golden vectors minted from the unmodified reference executor (tests/golden/make_sql_golden.py)
and against the oracle's restatement (oracle/aqe_oracle_sql.c).

Stated tolerances: group keys, row counts and integer-column sums bit-exact; fp64 values relative <= 1e-12 of the
reference's result (which itself is rounded to 15 significant digits by SQLite's TEXT interface); interval ends to
1e-12 of the value's magnitude or 1e-6 of the half width (the reference's own sum_sq - sum^2/n cancels 15-digit
inputs).  The engine's fp64 sums are additionally checked to be the EXACTLY rounded sums (math.fsum) for U(1,1000)
data, and bit-identical across runs, shard counts and row visiting order.
"""
import math
import os

import numpy as np
import pytest

import approximatequeryengine_b200 as aqe
from oracle import SqlError
from sql_helpers import MODE_OF, REL, WIDE_RANGE_QUERIES, engine_rows, golden_rows, load, rows_close, sql_golden_files, wide_range_rows

pytestmark = pytest.mark.gpu

FILES = sql_golden_files()


@pytest.fixture(scope="module")
def tables(oracle):
    out = []
    for path in FILES:
        g = load(path)
        rows = oracle.synth(g["n"], seed=g["seed"])
        out.append((g, rows, aqe.Engine(0).from_rows(rows)))
    return out


def run_engine(e, sql, p, mode):
    """What the binding does (aqe_pybind.cpp CustomBPlusDB::sql): the non-grouped calls ignore GROUP BY, the grouped
    ones require it."""
    q = aqe.sql_parse(sql, p)
    grouped = "groupby" in mode
    if grouped and q.group_col < 0:
        raise RuntimeError("No GROUP BY column found")
    if not grouped:
        q.group_col = -1
    rows = (aqe.SqlRow * aqe.SQL_MAX_GROUPS)()
    n = aqe.C.c_uint32()
    aqe.check(e.L.aqe_sql_execute(e.h, aqe.C.byref(q), aqe.SQL_MODE[MODE_OF[mode]], rows, aqe.SQL_MAX_GROUPS, aqe.C.byref(n)))
    return engine_rows(rows[: n.value])


def check_case(e, c):
    tag = (c["sql"], c["p"], c["mode"])
    want_err = c.get("error")
    try:
        got = run_engine(e, c["sql"], c["p"], c["mode"])
    except ValueError:
        return None if want_err in ("stod", "terminate") else (tag, "engine raised stod", want_err)
    except RuntimeError as ex:  # AqeError is a RuntimeError
        if "OR outside parentheses" in str(ex):
            return None   # refused on purpose (include/aqe_b200.h, aqe_sql_query.top_level_or)
        return None if want_err == "runtime_error" else (tag, f"engine raised {ex}", want_err)
    if want_err == "terminate" or (want_err == "runtime_error" and "integer overflow" in c.get("msg", "")):
        return None  # the reference aborts / SQLite's int64 SUM(col*col) overflows; the 128-bit accumulators do not
    if want_err:
        return (tag, "engine returned rows", want_err)
    why = rows_close(got, golden_rows(c), REL)
    return (tag, why) if why else None


def test_sql_against_reference_golden(tables):
    for g, rows, e in tables:
        bad = [b for b in (check_case(e, c) for c in g["cases"]) if b]
        assert not bad, (g["n"], bad[:5])


def test_sql_against_oracle_extra_queries(tables, oracle):
    g, rows, e = [t for t in tables if t[0]["n"] == 20000][0]
    t0 = 1700000000
    queries = [
        f"SELECT SUM(amount) FROM sales WHERE timestamp BETWEEN {t0 + 100} AND {t0 + 15000} AND region >= 2 AND product_id < 500 AND amount > 3.5 AND id != 777",
        "SELECT AVG(region) FROM sales WHERE amount <= 10",
        "SELECT COUNT(region) FROM sales WHERE product_id = 999 GROUP BY region",
        "SELECT SUM(id) FROM sales WHERE amount BETWEEN 400 AND 600 GROUP BY product_id",
        "SELECT AVG(amount) FROM sales WHERE id > 19990 GROUP BY region",
        "SELECT SUM(timestamp) FROM sales WHERE region = 7 GROUP BY region",
        "SELECT COUNT(*) FROM sales GROUP BY product_id",
        "SELECT SUM(amount) FROM sales WHERE (region = 1 OR region = 6) AND (amount < 50 OR amount > 950)",
        "SELECT AVG(amount) FROM sales WHERE (product_id < 10 OR product_id >= 990 OR region = 2) GROUP BY region",
        "SELECT COUNT(amount) FROM sales WHERE (id <= 100 OR timestamp > 1700019000) GROUP BY product_id",
        "SELECT SUM(amount) FROM sales WHERE amount < 0",
        "SELECT COUNT(amount) FROM sales WHERE amount < 0 GROUP BY region",
        "SELECT SUM(amount) FROM sales WHERE region != 1 AND region != 3 AND region != 6 AND product_id != 10 AND product_id != 20 GROUP BY region",
        "SELECT AVG(amount) FROM sales WHERE region NOT IN (0, 7) AND region != 4 AND amount != 500 AND amount != 250.5",
        # branches that differ in one small-domain integer column fold into one membership bitmap (aqe_engine.cu, sql_scan_impl)
        "SELECT SUM(amount) FROM sales WHERE region IN (1, 3, 5, 7)",
        "SELECT AVG(amount) FROM sales WHERE region IN (0, 2, 6) AND amount BETWEEN 100 AND 900 AND product_id != 17 GROUP BY region",
        "SELECT COUNT(*) FROM sales WHERE region NOT IN (1, 2, 3) GROUP BY product_id",
        "SELECT SUM(timestamp) FROM sales WHERE (region < 2 OR region > 5 OR region = 4)",
        "SELECT SUM(amount) FROM sales WHERE region IN (8, 9)",
        "SELECT SUM(amount) FROM sales WHERE product_id IN (1, 3, 5, 7) AND region = 2",      # 1000 keys: the bitmap lives in shared memory
        "SELECT AVG(amount) FROM sales WHERE product_id NOT IN (0, 31, 32, 63, 64, 999) AND product_id != 500 GROUP BY region",
        "SELECT COUNT(*) FROM sales WHERE (product_id < 10 OR product_id > 990 OR product_id BETWEEN 400 AND 420) GROUP BY product_id",
        # ... in a floating-point or wide column: one predicate pass per branch
        "SELECT SUM(amount) FROM sales WHERE amount NOT BETWEEN 100 AND 500",
        "SELECT AVG(amount) FROM sales WHERE (amount < 50 OR amount > 950 OR amount BETWEEN 400 AND 410) AND region >= 2 GROUP BY region",
        "SELECT SUM(amount) FROM sales WHERE amount != 500 AND amount != 250.5 AND amount != 7 GROUP BY product_id",
        f"SELECT SUM(region) FROM sales WHERE timestamp NOT BETWEEN {t0 + 5000} AND {t0 + 15000} AND amount < 700",
        "SELECT COUNT(id) FROM sales WHERE (id <= 100 OR id > 19900 OR id = 5000) GROUP BY region",
    ]
    for sql in queries:
        for p in (0, 3, 10, 25, 50, 99):
            for mode in ("run_query", "run_query_with_ci", "run_query_groupby", "run_query_groupby_with_ci"):
                if ("GROUP BY" in sql) != ("groupby" in mode):
                    continue
                tag = (sql, p, mode)
                try:
                    want = oracle.sql(rows, sql, p, mode)
                except SqlError as ex:
                    if ex.kind == "stod" or "integer overflow" in ex.msg:
                        if ex.kind == "stod":
                            with pytest.raises(ValueError):
                                run_engine(e, sql, p, mode)
                        continue
                    with pytest.raises(RuntimeError):
                        run_engine(e, sql, p, mode)
                    continue
                got = run_engine(e, sql, p, mode)
                assert rows_close(got, want, REL) is None, (tag, rows_close(got, want, REL))


def test_sql_sums_are_exactly_rounded_and_integer_sums_exact(tables):
    for g, rows, e in tables:
        r = e.sql("SELECT SUM(amount) FROM sales")[0]
        assert r.value == math.fsum(rows["amount"]) and r.count == len(rows)
        for k, row in zip(range(8), e.sql("SELECT SUM(amount) FROM sales GROUP BY region")):
            m = rows["region"] == row.key
            assert row.count == int(m.sum()) and row.value == math.fsum(rows["amount"][m])
        for col in ("id", "timestamp", "region", "product_id"):
            assert e.sql(f"SELECT SUM({col}) FROM sales WHERE amount >= 500")[0].isum == int(rows[col][rows["amount"] >= 500].astype(object).sum())
        by_pid = e.sql("SELECT SUM(timestamp) FROM sales GROUP BY product_id")
        assert sum(r.isum for r in by_pid) == int(rows["timestamp"].astype(object).sum())
        a, b = (e.sql("SELECT SUM(amount) FROM sales GROUP BY product_id", 10) for _ in range(2))    # shared-atomic bins: still bit stable
        assert [(r.key, r.count, r.sum) for r in a] == [(r.key, r.count, r.sum) for r in b]


def test_sql_non_dense_ids_take_the_modulus_predicate(oracle):
    rng = np.random.default_rng(3)
    rows = oracle.synth(30011, seed=21)
    rows["id"] = np.sort(rng.choice(10**7, size=len(rows), replace=False)) - 5000   # gaps, some negative ids
    e = aqe.Engine(0).from_rows(rows)
    for sql, mode in (("SELECT SUM(amount) FROM sales WHERE region < 6", "run_query"), ("SELECT AVG(amount) FROM sales", "run_query_with_ci"),
                      ("SELECT COUNT(amount) FROM sales GROUP BY region", "run_query_groupby"),
                      ("SELECT SUM(amount) FROM sales WHERE amount > 20 GROUP BY region", "run_query_groupby_with_ci")):
        for p in (0, 7, 20, 50):
            want = oracle.sql(rows, sql, p, mode)
            got = run_engine(e, sql, p, mode)
            assert rows_close(got, want, REL) is None, (sql, p, mode, rows_close(got, want, REL))
    e.close()


def test_sql_dense_ids_with_offset_and_ragged_sizes(oracle):
    for n, first_id in ((0, 1), (1, 1), (3, 5), (4, 2), (257, -100), (4099, 1000003), (65537, 1)):
        rows = oracle.synth(n, seed=33)
        rows["id"] = np.arange(first_id, first_id + n)
        e = aqe.Engine(0).from_rows(rows)
        for p in (0, 10, 33, 50):
            for sql, mode in (("SELECT COUNT(*) FROM sales", "run_query"), ("SELECT COUNT(amount) FROM sales WHERE amount > 500", "run_query"),
                              ("SELECT COUNT(amount) FROM sales GROUP BY region", "run_query_groupby")):
                want = oracle.sql(rows, sql, p, mode)
                assert rows_close(run_engine(e, sql, p, mode), want, REL) is None, (n, first_id, p, sql)
            if n:
                try:
                    want = oracle.sql(rows, "SELECT SUM(amount) FROM sales", p, "run_query_with_ci")
                except SqlError:
                    with pytest.raises(ValueError):
                        run_engine(e, "SELECT SUM(amount) FROM sales", p, "run_query_with_ci")
                    continue
                assert rows_close(run_engine(e, "SELECT SUM(amount) FROM sales", p, "run_query_with_ci"), want, REL) is None, (n, first_id, p)
        e.close()


def test_sql_unaligned_attached_columns(oracle):
    import torch
    rows = oracle.synth(50001, seed=9)
    cols = {c: torch.from_numpy(rows[c].copy()).cuda() for c in ("id", "amount", "region", "product_id", "timestamp")}
    for off in (0, 1, 3):
        sub = rows[off:]
        e = aqe.Engine(0).attach(len(sub), **{c: t[off:].data_ptr() for c, t in cols.items()})
        for sql, p, mode in (("SELECT SUM(amount) FROM sales WHERE timestamp > 1700000100 GROUP BY region", 0, "run_query_groupby"),
                             ("SELECT AVG(amount) FROM sales WHERE product_id < 300", 10, "run_query_with_ci"),
                             ("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, "run_query_groupby")):
            assert rows_close(run_engine(e, sql, p, mode), oracle.sql(sub, sql, p, mode), REL) is None, (off, sql)
        e.close()


def test_sql_shards_merge_bit_exactly(tables, oracle):
    """Three shards' accumulators merged on the host == the single-table accumulators, word for word."""
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    cuts = [0, 33333, 71003, len(rows)]
    shards = [aqe.Engine(0).from_rows(rows[a:b]) for a, b in zip(cuts, cuts[1:])]
    for sql, p, flags in (("SELECT SUM(amount) FROM sales WHERE amount < 900 GROUP BY region", 10, aqe.SQL_MOMENTS),
                          ("SELECT AVG(amount) FROM sales GROUP BY product_id", 0, aqe.SQL_MOMENTS), ("SELECT SUM(timestamp) FROM sales", 7, aqe.SQL_MOMENTS)):
        q = aqe.sql_parse(sql, p)
        layout = aqe.sql_layout(q, [s.sql_facts(q) for s in shards])
        whole_layout = aqe.sql_layout(q, [e.sql_facts(q)])
        assert bytes(layout) == bytes(whole_layout)
        acc = np.zeros(layout.n_groups * 5, dtype=np.uint64)
        for s in shards:
            aqe.sql_merge(acc, s.sql_scan(q, layout, flags))
        whole = e.sql_scan(q, layout, flags)
        assert (acc == whole).all(), sql
        a = aqe.sql_finish(q, layout, acc, "ci_reference")
        b = e.sql(sql, p, "ci_reference")
        assert [(r.key, r.count, r.value, r.ci_lower, r.ci_upper) for r in a] == [(r.key, r.count, r.value, r.ci_lower, r.ci_upper) for r in b], sql
    for s in shards:
        s.close()


def test_sql_large_table_properties():
    """BASELINE.json's full size (1 B rows generated on the device; AQE_TEST_FULL_N overrides): linearity of the integer
    accumulators, agreement with the exact scan kernel, exact counts through the strided sample."""
    n = int(os.environ.get("AQE_TEST_FULL_N", 1_000_000_000))
    e = aqe.Engine(0).generate(n, seed=7, columns=("id", "amount", "region", "product_id"))
    q_all = aqe.sql_parse("SELECT SUM(amount) FROM sales GROUP BY region", 0)
    layout = aqe.sql_layout(q_all, [e.sql_facts(q_all)])
    assert layout.n_groups == 8 and layout.sum_shift == 52
    whole = e.sql_scan(q_all, layout)
    lo = e.sql_scan(aqe.sql_parse("SELECT SUM(amount) FROM sales WHERE amount < 300 GROUP BY region", 0), layout)
    hi = e.sql_scan(aqe.sql_parse("SELECT SUM(amount) FROM sales WHERE amount >= 300 GROUP BY region", 0), layout)
    assert (aqe.sql_merge(lo.copy(), hi) == whole).all()
    assert (e.sql_scan(q_all, layout) == whole).all()                      # run-to-run identical
    rows = aqe.sql_finish(q_all, layout, whole)
    assert sum(r.count for r in rows) == n
    # with squares the private bins pack (rows, sum, squares) into three words and drain every ~4000 rows per thread: the count and
    # sum words must be those of the plain scan, the squares linear over a split, and the strided visit (other kernel) must agree
    with_sq = e.sql_scan(q_all, layout, aqe.SQL_MOMENTS)
    assert (with_sq.reshape(-1, 5)[:, :3] == whole.reshape(-1, 5)[:, :3]).all()
    lo_sq = e.sql_scan(aqe.sql_parse("SELECT SUM(amount) FROM sales WHERE amount < 300 GROUP BY region", 0), layout, aqe.SQL_MOMENTS)
    hi_sq = e.sql_scan(aqe.sql_parse("SELECT SUM(amount) FROM sales WHERE amount >= 300 GROUP BY region", 0), layout, aqe.SQL_MOMENTS)
    assert (aqe.sql_merge(lo_sq.copy(), hi_sq) == with_sq).all()
    halves = [e.sql_scan(aqe.sql_parse(f"SELECT SUM(amount) FROM sales WHERE id {op} {n // 2} GROUP BY region", 50), layout, aqe.SQL_MOMENTS) for op in ("<=", ">")]
    assert (aqe.sql_merge(halves[0].copy(), halves[1]) == e.sql_scan(aqe.sql_parse("SELECT SUM(amount) FROM sales GROUP BY region", 50), layout, aqe.SQL_MOMENTS)).all()
    total = e.sql("SELECT SUM(amount) FROM sales")[0]
    p = e.scan("amount")
    assert abs(total.value - p.sum) <= 4 * math.ulp(p.sum)                 # compensated scan vs exactly rounded fixed point
    assert abs(math.fsum(r.value for r in rows) - total.value) <= 8 * math.ulp(total.value)
    by_pid = e.sql("SELECT COUNT(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY product_id")
    w = e.scan("amount", "amount", 100.0, 500.0)
    assert len(by_pid) == 1000 and sum(r.count for r in by_pid) == w.count
    # 1000 groups (packed shared bins, three atomics per row): the groups' integer accumulators add up to the ungrouped scan's, word for
    # word, and equal the general form's (AQE_SQL_PACKED=0)
    q_pid = aqe.sql_parse("SELECT SUM(amount) FROM sales GROUP BY product_id", 0)
    l_pid = aqe.sql_layout(q_pid, [e.sql_facts(q_pid)])
    assert l_pid.n_groups == 1000 and l_pid.sum_shift == layout.sum_shift
    pid = e.sql_scan(q_pid, l_pid, aqe.SQL_MOMENTS).reshape(-1, 5)
    q_one = aqe.sql_parse("SELECT SUM(amount) FROM sales", 0)
    one = e.sql_scan(q_one, aqe.sql_layout(q_one, [e.sql_facts(q_one)]), aqe.SQL_MOMENTS)
    as_int = lambda lo, hi: int(lo) + (int(hi) << 64)
    assert sum(int(c) for c in pid[:, 0]) == int(one[0]) == n
    assert sum(as_int(r[1], r[2]) for r in pid) % (1 << 128) == as_int(one[1], one[2])
    assert sum(as_int(r[3], r[4]) for r in pid) % (1 << 128) == as_int(one[3], one[4])
    os.environ["AQE_SQL_PACKED"] = "0"
    try:
        assert (e.sql_scan(q_pid, l_pid, aqe.SQL_MOMENTS).reshape(-1, 5) == pid).all()
    finally:
        del os.environ["AQE_SQL_PACKED"]
    # 1-in-10 systematic sample through the strided visit: count is exact, estimate is close
    s = e.sql("SELECT SUM(amount) FROM sales", 10, "ci_correct")[0]
    half = (s.ci_upper - s.ci_lower) / 2
    assert s.count == n // 10 and 0 < half < 2e-3 * total.value and abs(s.value - total.value) < 4 * half / 1.96, (s.value, total.value, half)
    e.close()


def test_dropin_run_query_functions(oracle, tmp_path):
    b = aqe.backend()
    rows = oracle.synth(20000, seed=7)
    path = str(tmp_path / "sales.aqe")
    oracle.save_file(path, rows)
    v = b.run_query("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", path, 10)
    assert abs(v - oracle.sql(rows, "SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10)[0][1]) <= REL * abs(v)
    # the non-grouped entry points ignore GROUP BY, as execute_query does
    assert b.run_query("SELECT COUNT(amount) FROM sales GROUP BY region", path) == 20000.0
    d = b.run_query_groupby("SELECT AVG(amount) FROM sales GROUP BY region", path, 20, 2)
    want = oracle.sql(rows, "SELECT AVG(amount) FROM sales GROUP BY region", 20, "run_query_groupby")
    assert list(d) == sorted(str(k) for k, *_ in want) and all(abs(d[str(k)] - val) <= REL * abs(val) for k, val, *_ in want)
    r = b.run_query_with_ci("SELECT AVG(amount) FROM sales", path, 5)
    k, val, lo, hi = oracle.sql(rows, "SELECT AVG(amount) FROM sales", 5, "run_query_with_ci")[0]
    assert abs(r.value - val) <= REL * val and abs(r.ci_lower - lo) <= REL * val and abs(r.ci_upper - hi) <= REL * val
    dc = b.run_query_groupby_with_ci("SELECT SUM(amount) FROM sales GROUP BY region", path, 10)
    want = {str(k): (val, lo, hi) for k, val, lo, hi in oracle.sql(rows, "SELECT SUM(amount) FROM sales GROUP BY region", 10, "run_query_groupby_with_ci")}
    assert set(dc) == set(want) and all(abs(dc[k].value - want[k][0]) <= REL * abs(want[k][0]) for k in want)
    keys = list(b.run_query_groupby("SELECT COUNT(amount) FROM sales GROUP BY product_id", path))
    assert keys == sorted(keys) and keys[:3] == ["0", "1", "10"]        # std::map<std::string,...>: keys sort as text
    # error behaviour (bindings.cpp: std::runtime_error -> RuntimeError, std::invalid_argument -> ValueError)
    with pytest.raises(RuntimeError, match="Unsupported aggregation function"):
        b.run_query("SELECT MAX(amount) FROM sales", path)
    with pytest.raises(RuntimeError, match="No GROUP BY"):
        b.run_query_groupby("SELECT SUM(amount) FROM sales", path)
    with pytest.raises(ValueError, match="stod"):
        b.run_query("SELECT SUM(amount) FROM sales WHERE amount > 5000", path)
    with pytest.raises(RuntimeError, match="Cannot open database"):
        b.run_query("SELECT SUM(amount) FROM sales", str(tmp_path / "missing.aqe"))
    # the resident copy follows the file
    rows2 = oracle.synth(5000, seed=8)
    oracle.save_file(path, rows2)
    os.utime(path, ns=(1, 1))
    assert b.run_query("SELECT COUNT(*) FROM sales", path) == 5000.0
    # the same calls on an open table
    db = b.CustomBPlusDB(); db.open_database(path)
    assert db.query("SELECT COUNT(*) FROM sales WHERE region = 1") == float((rows2["region"] == 1).sum())
    c = db.query_with_ci("SELECT SUM(amount) FROM sales", 10, correct_ci=True)
    assert c.ci_lower < math.fsum(rows2["amount"]) < c.ci_upper
    b.close_cached_tables()


def _random_query(rng, t0, n):
    agg = rng.choice(["SUM", "AVG", "COUNT"])
    col = rng.choice(["amount", "amount", "region", "product_id", "timestamp", "id"])
    terms = []
    for _ in range(rng.integers(0, 4)):
        c = rng.choice(["amount", "region", "product_id", "timestamp", "id"])
        if c == "amount":
            a, b = sorted(rng.uniform(-50, 1200, size=2))
            terms.append(rng.choice([f"amount BETWEEN {a:.3f} AND {b:.3f}", f"amount > {a:.2f}", f"amount <= {b:.1f}", f"{a:.2f} < amount", f"amount != {a:.0f}"]))
        elif c == "region":
            k = int(rng.integers(-1, 9))
            terms.append(rng.choice([f"region = {k}", f"region != {k}", f"region >= {k}", f"region < {k}.5", f"region BETWEEN 2 AND {k}", f"region <> '{k}'", f"region IN ({k}, 2, 7)"]))
        elif c == "product_id":
            a, b = sorted(int(x) for x in rng.integers(-5, 1100, size=2))
            terms.append(rng.choice([f"product_id BETWEEN {a} AND {b}", f"product_id < {b}", f"(product_id >= {a})", f"product_id = {a}"]))
        elif c == "timestamp":
            a = t0 + int(rng.integers(-10, n + 10))
            terms.append(rng.choice([f"timestamp > {a}", f"timestamp <= {a}", f"timestamp BETWEEN {t0} AND {a}"]))
        else:
            a = int(rng.integers(-100, 3 * n))
            terms.append(rng.choice([f"id > {a}", f"rowid <= {a}", f"id != {a}"]))
    group = rng.choice([None, None, "region", "product_id", "ts_bucket"])
    sql = f"SELECT {agg}({col}) FROM sales"
    if terms and rng.random() < 0.2:                 # a negated first term
        terms[0] = f"NOT ({terms[0]})"
    if len(terms) >= 2 and rng.random() < 0.4:      # a parenthesised OR of the first two terms, AND-ed with the rest
        terms = [f"({terms[0]} OR {terms[1]})"] + terms[2:]
    if len(terms) >= 2 and rng.random() < 0.15:     # ... and sometimes the whole clause as one parenthesised OR
        terms = ["(" + " OR ".join(terms) + ")"]
    if terms:
        sql += " WHERE " + " AND ".join(terms)
    return sql, group


def test_sql_random_queries_against_oracle(oracle):
    """Seeded random tables (signed heavy-tailed amounts, ids with gaps, a narrow int64 GROUP BY column) x random queries
    from the supported grammar x random sample percentages, all four entry points, against the oracle."""
    rng = np.random.default_rng(20261018)
    t0 = 1700000000
    for trial in range(6):
        n = int(rng.choice([1, 37, 1000, 4097, 30000]))
        rows = oracle.synth(n, seed=100 + trial)
        if trial % 2:
            rows["amount"] = np.exp(rng.normal(2.0, 2.5, size=n)) * rng.choice([-1.0, 1.0], size=n, p=[0.2, 0.8])   # signed, 6 decades
            rows["id"] = np.sort(rng.choice(4 * n + 10, size=n, replace=False)) - 7
        if trial % 3 == 0:
            rows["timestamp"] = t0 + (np.arange(n) % 40)       # int64 GROUP BY column with a narrow range
        e = aqe.Engine(0).from_rows(rows)
        scale = float(np.abs(rows["amount"]).sum()) or 1.0
        for _ in range(60):
            sql, group = _random_query(rng, t0, n)
            if group == "ts_bucket":
                if trial % 3:
                    continue
                group = "timestamp"
            if group:
                sql += f" GROUP BY {group}"
            p = int(rng.choice([0, 0, 1, 5, 10, 13, 25, 50, 99, 100]))
            for mode in (("run_query_groupby", "run_query_groupby_with_ci") if group else ("run_query", "run_query_with_ci")):
                tag = (trial, n, sql, p, mode)
                try:
                    want = oracle.sql(rows, sql, p, mode)
                except SqlError as ex:
                    if ex.kind == "stod":
                        with pytest.raises(ValueError):
                            run_engine(e, sql, p, mode)
                    elif "integer overflow" not in ex.msg:
                        with pytest.raises(RuntimeError):
                            run_engine(e, sql, p, mode)
                    continue
                try:
                    got = run_engine(e, sql, p, mode)
                except aqe.AqeError as ex:
                    assert ex.code == 6 and ("more than one !=" in str(ex) or "OR branches" in str(ex)), tag    # the documented gaps this generator can hit
                    continue
                assert len(got) == len(want), tag
                for g, w in zip(got, want):
                    assert g[0] == w[0], tag
                    # fixed-point sums are exact to 2^-63 of the column's largest magnitude per row; with signed data the
                    # reference's own 15-digit result can cancel, so values are compared on the scale of sum |x|
                    tol = 1e-12 * max(abs(w[1]), scale if "amount" in sql.split("FROM")[0] else abs(w[1]))
                    assert abs(g[1] - w[1]) <= tol or (math.isnan(g[1]) and math.isnan(w[1])) or g[1] == w[1], (tag, g, w)
                    # interval widths: the reference forms sum_sq - sum^2/n from 15-digit TEXT; for timestamp / id (values ~1e9
                    # with a tiny spread) that difference is rounding noise in the reference itself, so it is not compared there
                    well_conditioned = not any(c in sql.split("FROM")[0] for c in ("timestamp", "(id)"))
                    if well_conditioned and math.isfinite(w[2]) and math.isfinite(w[3]) and math.isfinite(g[2]):
                        half_w, half_g = (w[3] - w[2]) / 2, (g[3] - g[2]) / 2
                        assert abs(half_g - half_w) <= 1e-6 * abs(half_w) + tol, (tag, g, w)
        e.close()


def test_sql_register_kernel_and_ring_kernel_agree_word_for_word(tables, monkeypatch):
    """k_sql_agg (register-staged: strided samples, unaligned columns) forced onto full scans (AQE_SQL_VARIANT=1) leaves the
    same integer accumulators as k_sql_ring, and so does the ring's row-number filter forced onto every small step (=2): the fixed-point sums do not depend on which kernel, grid or order visited the rows."""
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    for sql, p, flags in (("SELECT SUM(amount) FROM sales WHERE amount > 250.5 AND region != 4", 0, aqe.SQL_MOMENTS),
                          ("SELECT AVG(amount) FROM sales WHERE timestamp <= 1700090000 GROUP BY region", 50, aqe.SQL_MOMENTS),
                          ("SELECT SUM(timestamp) FROM sales GROUP BY product_id", 0, 0),
                          ("SELECT COUNT(amount) FROM sales WHERE id > 77 GROUP BY product_id", 25, 0),
                          ("SELECT SUM(amount) FROM sales GROUP BY region", 34, aqe.SQL_MOMENTS),
                          ("SELECT AVG(amount) FROM sales", 20, aqe.SQL_MOMENTS),
                          ("SELECT SUM(amount) FROM sales WHERE region IN (1, 4, 6) AND amount > 10 GROUP BY region", 34, aqe.SQL_MOMENTS),
                          ("SELECT SUM(timestamp) FROM sales WHERE region NOT IN (0, 3)", 0, 0),
                          ("SELECT SUM(amount) FROM sales WHERE product_id IN (5, 64, 65, 999) AND amount < 900 GROUP BY region", 0, aqe.SQL_MOMENTS),
                          ("SELECT COUNT(*) FROM sales WHERE product_id NOT IN (1, 2, 3) GROUP BY product_id", 20, 0),
                          ("SELECT SUM(amount) FROM sales WHERE amount NOT BETWEEN 200 AND 800 GROUP BY region", 0, aqe.SQL_MOMENTS),
                          ("SELECT SUM(amount) FROM sales WHERE (amount < 10 OR amount > 990)", 50, aqe.SQL_MOMENTS),
                          ("SELECT AVG(timestamp) FROM sales WHERE (id < 1000 OR id > 90000) AND amount > 5", 25, 0),
                          ("SELECT SUM(amount) FROM sales GROUP BY region", 5, aqe.SQL_MOMENTS)):
        q = aqe.sql_parse(sql, p)
        layout = aqe.sql_layout(q, [e.sql_facts(q)])
        monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
        auto = e.sql_scan(q, layout, flags)              # steps 2-3 with a WHERE clause: ring + row-number filter; other samples: strided visit
        monkeypatch.setenv("AQE_SQL_VARIANT", "2")
        ring = e.sql_scan(q, layout, flags)              # ring + row-number filter for every step below 8
        monkeypatch.setenv("AQE_SQL_VARIANT", "1")
        regs = e.sql_scan(q, layout, flags)
        monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
        assert (ring == regs).all() and (auto == regs).all(), sql
        # the packed row counters of the private bins are drained before they can overflow: force a drain at every tile / batch
        monkeypatch.setenv("AQE_SQL_DRAIN_ROWS", "16")
        monkeypatch.setenv("AQE_SQL_VARIANT", "2")
        drained_ring = e.sql_scan(q, layout, flags)
        monkeypatch.setenv("AQE_SQL_VARIANT", "1")
        drained_regs = e.sql_scan(q, layout, flags)
        monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
        monkeypatch.delenv("AQE_SQL_DRAIN_ROWS", raising=False)
        assert (ring == drained_ring).all() and (ring == drained_regs).all(), sql


def test_sql_beyond_two_to_the_32_rows():
    """4.5 B rows of one column (36 GB) on one GPU: row counts above 2^32 and the 128-bit sums stay exact."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    n = 4_500_000_000
    if free < n * 12 + (4 << 30):
        pytest.skip("not enough free HBM for 54 GB of columns")
    e = aqe.Engine(0).generate(n, seed=7, columns=("amount", "region"))
    r = e.sql("SELECT SUM(amount) FROM sales")[0]
    p = e.scan("amount")
    assert r.count == n == p.count and abs(r.value - p.sum) <= 4 * math.ulp(p.sum)
    w = e.sql("SELECT COUNT(amount) FROM sales WHERE amount BETWEEN 100 AND 500")[0]
    assert w.count == e.scan("amount", "amount", 100.0, 500.0).count and w.count > 2**30
    a, b = e.sql("SELECT SUM(amount) FROM sales WHERE amount < 300")[0], e.sql("SELECT SUM(amount) FROM sales WHERE amount >= 300")[0]
    assert a.count + b.count == n and abs((a.value + b.value) - r.value) <= 2 * math.ulp(r.value)
    # integer aggregates: the exact-scan kernel goes in 2^32-row segments, the SQL kernel in one launch; same int128
    by_region = e.sql("SELECT SUM(region) FROM sales GROUP BY region")
    assert e.sum_int("region") == e.sql("SELECT SUM(region) FROM sales")[0].isum == sum(g.isum for g in by_region)
    assert sum(g.count for g in by_region) == n and all(g.isum == g.key * g.count for g in by_region)
    e.close()


def test_sql_corrected_interval_covers_the_truth():
    """AQE_SQL_CI_CORRECT (additive): over 300 seeded tables the 95 % interval of the 1-in-10 sampled SUM and AVG contains
    the exact answer at about the nominal rate (the reference's own SUM interval reports mean * 100/p, not a total)."""
    hits_sum = hits_avg = 0
    trials = 300
    e = aqe.Engine(0)
    for seed in range(trials):
        e.generate(200_000, seed=1000 + seed, columns=("id", "amount"))
        exact = e.sql("SELECT SUM(amount) FROM sales")[0]
        s = e.sql("SELECT SUM(amount) FROM sales", 10, "ci_correct")[0]
        a = e.sql("SELECT AVG(amount) FROM sales", 10, "ci_correct")[0]
        hits_sum += s.ci_lower <= exact.value <= s.ci_upper
        hits_avg += a.ci_lower <= exact.value / exact.count <= a.ci_upper
        ref = e.sql("SELECT SUM(amount) FROM sales", 10, "ci_reference")[0]
        assert not (ref.ci_lower <= exact.value <= ref.ci_upper)        # executor.cpp:225-241: a mean scaled by 100/p
    e.close()
    sigma = math.sqrt(0.95 * 0.05 / trials)
    assert hits_sum / trials >= 0.95 - 3 * sigma and hits_avg / trials >= 0.95 - 3 * sigma, (hits_sum, hits_avg)


def test_sql_small_values_in_a_wide_range_column(oracle):
    """amount spans 1e-7 .. 1e6; clauses that bound amount itself set the fixed-point scale (sql_layout), so sums of the
    1e-6-sized rows stay within 1e-12 of the reference's -- and bit-identical between the two kernels."""
    rows = wide_range_rows(oracle, n=200003)
    e = aqe.Engine(0).from_rows(rows)
    for sql, p, mode in WIDE_RANGE_QUERIES + (("SELECT SUM(amount) FROM sales", 0, "run_query"),
                                              ("SELECT AVG(amount) FROM sales WHERE amount >= 1000 GROUP BY region", 10, "run_query_groupby_with_ci")):
        got = run_engine(e, sql, p, mode)
        assert rows_close(got, oracle.sql(rows, sql, p, mode), REL) is None, (sql, p, rows_close(got, oracle.sql(rows, sql, p, mode), REL))
    x = rows["amount"]
    r = e.sql("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 0 AND 0.00001", 0)[0]
    assert r.value == math.fsum(x[(x >= 0) & (x <= 1e-5)])   # 2^-78 grid: every selected double is on it
    e.close()


def test_sql_negative_values_and_keys_and_the_group_limit(oracle):
    """Signed fixed point (amount ~ U(-1000, 1000)), negative group keys, exactly SQL_MAX_GROUPS keys, one more than that, and
    groups of a single row (GROUP BY timestamp: n - 1 = 0 in the reference's variance)."""
    from sql_helpers import SIGNED_QUERIES, signed_rows
    rows = signed_rows(oracle)
    e = aqe.Engine(0).from_rows(rows)
    for sql, p, mode in SIGNED_QUERIES:
        try:
            want = oracle.sql(rows, sql, p, mode)
        except SqlError as ex:
            assert ex.kind in ("stod", "terminate"), (sql, ex)
            with pytest.raises(ValueError):
                run_engine(e, sql, p, mode)
            continue
        assert rows_close(run_engine(e, sql, p, mode), want, REL) is None, (sql, p, rows_close(run_engine(e, sql, p, mode), want, REL))
    x = rows["amount"]
    assert e.sql("SELECT SUM(amount) FROM sales")[0].value == math.fsum(x)                       # cancellation: still the exactly rounded sum
    assert e.sql("SELECT SUM(amount) FROM sales WHERE amount < 0")[0].value == math.fsum(x[x < 0])
    e.close()
    rows["product_id"][2] = 2048
    e = aqe.Engine(0).from_rows(rows)
    with pytest.raises(RuntimeError, match="key range wider"):
        e.sql("SELECT SUM(amount) FROM sales GROUP BY product_id")
    e.close()
    few = signed_rows(oracle, n=4096, seed=92)
    e = aqe.Engine(0).from_rows(few)
    for sql, p, mode in (("SELECT SUM(amount) FROM sales GROUP BY timestamp", 0, "run_query_groupby"),
                         ("SELECT AVG(amount) FROM sales WHERE amount > -500 GROUP BY id", 0, "run_query_groupby"),
                         ("SELECT SUM(amount) FROM sales GROUP BY timestamp", 0, "run_query_groupby_with_ci")):
        try:
            want = oracle.sql(few, sql, p, mode)
        except SqlError as ex:
            assert ex.kind in ("stod", "terminate"), (sql, ex)
            with pytest.raises(ValueError):
                run_engine(e, sql, p, mode)
            continue
        got = run_engine(e, sql, p, mode)
        assert len(got) == len(want) and all(a[0] == b[0] and a[1] == pytest.approx(b[1], rel=REL) for a, b in zip(got, want)), sql
        for a, b in zip(got, want):     # one row per group: the reference divides by n - 1 = 0; whatever it reports (nan / inf), the engine reports too
            for u, v in zip(a[2:], b[2:]):
                assert (math.isnan(u) and math.isnan(v)) or u == pytest.approx(v, rel=REL), (sql, a, b)
    e.close()


def _skewed_rows(oracle, n=120_000, seed=33):
    """Nine rows in ten carry ONE product_id, the rest spread over 1000 keys: a CTA of the grouped scan adds far more than
    kSqlSharedPackedLimit (1024) rows of a tile to one packed shared bin, so bins are emptied in mid-scan by the threads that see it."""
    rng = np.random.default_rng(seed)
    rows = oracle.synth(n, seed=seed)
    rows["product_id"] = np.where(rng.random(n) < 0.9, 417, rng.integers(0, 1000, n))
    rows["product_id"][:2] = (0, 999)
    return rows


PACKED_QUERIES = (("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, 0),
                  ("SELECT AVG(amount) FROM sales GROUP BY product_id", 0, aqe.SQL_MOMENTS),
                  ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY product_id", 0, 0),                 # few rows pass: the sparse walk
                  ("SELECT SUM(amount) FROM sales WHERE region = 3 AND amount < 120.5 GROUP BY product_id", 0, aqe.SQL_MOMENTS),
                  ("SELECT SUM(timestamp) FROM sales GROUP BY product_id", 0, 0),                                 # integer aggregate, values above 2^30
                  ("SELECT AVG(id) FROM sales WHERE id > 500 GROUP BY product_id", 0, aqe.SQL_MOMENTS),
                  ("SELECT SUM(region) FROM sales WHERE region >= 3 GROUP BY product_id", 0, 0),                  # pieces above bit 32 are all zero
                  ("SELECT SUM(amount) FROM sales GROUP BY product_id", 50, aqe.SQL_MOMENTS),                      # strided visit (k_sql_agg)
                  ("SELECT AVG(amount) FROM sales WHERE amount > 950 GROUP BY product_id", 10, 0),
                  ("SELECT SUM(product_id) FROM sales GROUP BY product_id", 0, aqe.SQL_MOMENTS))


def test_sql_packed_shared_bins_leave_the_words_of_the_general_form(tables, oracle, monkeypatch):
    """Mid-sized GROUP BY (17..4096 keys): the packed shared bins (three atomics per row, SqlBins MODE 3 -- the default whenever the
    aggregate column's fixed-point range allows) leave exactly the accumulator words of the general form (count word + limbs with
    carry chains, AQE_SQL_PACKED=0), through both kernels: on the uniform table, on a table whose keys are skewed enough that bins
    fill up and are emptied in mid-scan, and on signed data (bins hold value - column minimum)."""
    from sql_helpers import signed_rows
    uniform = [t for t in tables if t[0]["n"] == 100000][0][1]
    signed = signed_rows(oracle, n=150_001)
    signed["product_id"] = signed["product_id"] // 4          # 1024 keys, negative ones included
    for name, rows in (("uniform", uniform), ("skewed", _skewed_rows(oracle)), ("signed", signed)):
        e = aqe.Engine(0).from_rows(rows)
        for sql, p, flags in PACKED_QUERIES:
            q = aqe.sql_parse(sql, p)
            layout = aqe.sql_layout(q, [e.sql_facts(q)])
            assert 16 < layout.n_groups <= aqe.SQL_MAX_GROUPS
            got = {}
            for packed in ("1", "0"):
                monkeypatch.setenv("AQE_SQL_PACKED", packed)
                for variant in ("0", "1", "2"):
                    monkeypatch.setenv("AQE_SQL_VARIANT", variant)
                    got[packed, variant] = e.sql_scan(q, layout, flags)
            monkeypatch.delenv("AQE_SQL_PACKED", raising=False)
            monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
            ref = got["0", "1"]
            for k, v in got.items():
                assert (v == ref).all(), (name, sql, p, k)
            assert int(ref.reshape(-1, 5)[:, 0].sum()) > 0, (name, sql)
        e.close()


def test_sql_skewed_keys_against_oracle(oracle):
    """The same skewed table end to end against the oracle's restatement of the reference executor."""
    rows = _skewed_rows(oracle, n=60_000, seed=34)
    e = aqe.Engine(0).from_rows(rows)
    for sql, p, mode in (("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, "run_query_groupby"),
                         ("SELECT AVG(amount) FROM sales WHERE amount > 300 GROUP BY product_id", 0, "run_query_groupby_with_ci"),
                         ("SELECT SUM(timestamp) FROM sales GROUP BY product_id", 0, "run_query_groupby"),
                         ("SELECT AVG(amount) FROM sales WHERE product_id = 417 GROUP BY product_id", 50, "run_query_groupby_with_ci"),
                         ("SELECT SUM(amount) FROM sales GROUP BY product_id", 50, "run_query_groupby_with_ci")):
        try:
            want = oracle.sql(rows, sql, p, mode)
        except SqlError as ex:      # a key none of whose few rows was sampled: NULL -> stod in the reference (the last query)
            assert ex.kind in ("stod", "terminate"), (sql, ex)
            with pytest.raises(ValueError):
                run_engine(e, sql, p, mode)
            continue
        assert rows_close(run_engine(e, sql, p, mode), want, REL) is None, (sql, p, rows_close(run_engine(e, sql, p, mode), want, REL))
    x, k = rows["amount"], rows["product_id"]
    r = {row.key: row.value for row in e.sql("SELECT SUM(amount) FROM sales GROUP BY product_id")}
    assert r[417] == math.fsum(x[k == 417]) and r[999] == math.fsum(x[k == 999])      # exactly rounded, emptied bins included
    e.close()


def test_sql_wide_int64_sums_over_many_groups_take_the_general_form(oracle):
    """An int64 aggregate column whose values span more than 2^62 cannot use the packed shared bins (62-bit values): the general form
    (four limbs with carries) answers, exactly -- checked against Python integers per group."""
    n = 90_000
    rows = oracle.synth(n, seed=21)
    rng = np.random.default_rng(21)
    rows["timestamp"] = rng.integers(-(2 ** 62) - 12345, 2 ** 62 + 999, n)
    e = aqe.Engine(0).from_rows(rows)
    got = e.sql("SELECT SUM(timestamp) FROM sales GROUP BY product_id")
    want = {}
    for k, v in zip(rows["product_id"].tolist(), rows["timestamp"].tolist()):
        want[k] = want.get(k, 0) + v
    assert len(got) == len(want)
    for r in got:
        assert (r.isum_hi << 64) + r.isum_lo == want[r.key] and r.count == int((rows["product_id"] == r.key).sum()), r.key
    narrow = e.sql("SELECT SUM(id) FROM sales WHERE id > 1000 GROUP BY product_id")        # the packed form next to it, same table
    for r in narrow:
        m = (rows["product_id"] == r.key) & (rows["id"] > 1000)
        assert (r.isum_hi << 64) + r.isum_lo == int(rows["id"][m].astype(object).sum())
    e.close()


def test_sql_keys_outside_a_narrower_layout_are_dropped_by_every_bin_form(tables, monkeypatch):
    """aqe_sql_scan with a caller-made layout that covers only part of the key range: rows whose key falls outside leave no trace, and
    the groups inside hold exactly the words the full layout gives them -- thread-private bins (the branch-free row add sends such rows to
    a bin of their own that no drain reads), the same bins through SqlBins::add (AQE_SQL_PAIR_BINS=2 / 0), CTA-shared bins, both kernels."""
    g, rows, e = [t for t in tables if t[0]["n"] == 100000][0]
    for sql, p, flags, lo, cnt in (("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", 0, aqe.SQL_MOMENTS, 2, 3),
                                   ("SELECT SUM(amount) FROM sales GROUP BY region", 0, aqe.SQL_MOMENTS, 1, 6),
                                   ("SELECT SUM(timestamp) FROM sales WHERE amount > 5 GROUP BY region", 0, aqe.SQL_MOMENTS, 5, 1),
                                   ("SELECT AVG(amount) FROM sales GROUP BY region", 0, 0, 3, 4),
                                   ("SELECT COUNT(amount) FROM sales GROUP BY region", 0, 0, 0, 2),
                                   ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY product_id", 0, 0, 100, 50),
                                   ("SELECT AVG(amount) FROM sales GROUP BY product_id", 0, aqe.SQL_MOMENTS, 7, 300)):
        q = aqe.sql_parse(sql, p)
        full = aqe.sql_layout(q, [e.sql_facts(q)])
        whole = e.sql_scan(q, full, flags).reshape(-1, 5)
        part = aqe.SqlLayout.from_buffer_copy(bytes(full))
        first = full.key_min + lo
        part.key_min, part.n_groups = first, cnt
        want = whole[lo:lo + cnt]
        for pair_bins in ("1", "2", "0"):
            monkeypatch.setenv("AQE_SQL_PAIR_BINS", pair_bins)
            for variant in (None, "1"):
                if variant:
                    monkeypatch.setenv("AQE_SQL_VARIANT", variant)
                got = e.sql_scan(q, part, flags).reshape(-1, 5)
                monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
                assert (got == want).all(), (sql, pair_bins, variant)
        monkeypatch.delenv("AQE_SQL_PAIR_BINS", raising=False)


def test_sql_four_byte_rows_take_sixteen_rows_per_thread_and_agree(oracle, monkeypatch):
    """Ungrouped queries that read ONE int32 column run the ring with 16 rows per consumer thread and tile (K = 16; the grouped ones keep
    K = 8).  Same accumulator words as K = 8 (AQE_SQL_K16=0) and as the register-staged kernel, on whole tiles, ragged last tiles and
    tables smaller than a tile."""
    for n in (1, 255, 4095, 4096, 4097, 12_289, 100_003):
        rows = oracle.synth(n, seed=11 + n % 7)
        e = aqe.Engine(0).from_rows(rows)
        for sql, p in (("SELECT COUNT(*) FROM sales WHERE region = 1", 0), ("SELECT COUNT(*) FROM sales WHERE region IN (1, 3, 5, 7)", 0),
                       ("SELECT COUNT(*) FROM sales WHERE (region = 1 OR region = 3)", 0), ("SELECT COUNT(amount) FROM sales GROUP BY region", 0),
                       ("SELECT COUNT(amount) FROM sales GROUP BY product_id", 0), ("SELECT COUNT(*) FROM sales WHERE product_id NOT IN (1, 2, 3) GROUP BY product_id", 0),
                       ("SELECT SUM(region) FROM sales GROUP BY region", 0), ("SELECT SUM(product_id) FROM sales WHERE product_id > 500", 0),
                       ("SELECT COUNT(*) FROM sales WHERE region != 2 GROUP BY region", 3)):
            q = aqe.sql_parse(sql, p)
            layout = aqe.sql_layout(q, [e.sql_facts(q)])
            monkeypatch.setenv("AQE_SQL_VARIANT", "2")
            k16 = e.sql_scan(q, layout, 0)
            monkeypatch.setenv("AQE_SQL_K16", "0")
            k8 = e.sql_scan(q, layout, 0)
            monkeypatch.delenv("AQE_SQL_K16", raising=False)
            monkeypatch.setenv("AQE_SQL_DRAIN_ROWS", "16")
            drained = e.sql_scan(q, layout, 0)
            monkeypatch.delenv("AQE_SQL_DRAIN_ROWS", raising=False)
            monkeypatch.setenv("AQE_SQL_VARIANT", "1")
            regs = e.sql_scan(q, layout, 0)
            monkeypatch.delenv("AQE_SQL_VARIANT", raising=False)
            assert (k16 == regs).all() and (k8 == regs).all() and (drained == regs).all(), (n, sql)
        e.close()
