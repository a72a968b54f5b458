"""The oracle (oracle/aqe_oracle.c) against the UNMODIFIED reference compiled in place (oracle/_ref/libaqe_ref.so).
Runs only where that library exists (this container: /root/reference is present and `make -C oracle ref` built it;
elsewhere the golden vectors minted from it carry the pin -- tests/test_oracle_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import Ref, RefScheduler, make_params

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libaqe_ref.so not built (no /root/reference here)")

DETERMINISTIC = ["memory_stride", "slow_pointer", "fast_pointer", "dual_pointer", "parallel_pointer", "random_pointer",
                 "optimized_address_arithmetic", "index_based", "byte_offset", "optimized_clt", "block", "page", "parallel_block",
                 "node_skip", "balanced_tree", "direct_access", "adaptive_block", "stratified_block"]


@pytest.mark.parametrize("n", [254, 255, 256, 381, 382, 1000, 12345, 65537, 99999])
def test_index_sets_and_exact_sums_match_reference(oracle, n):
    rows = oracle.synth(n, seed=11)
    R = Ref(rows)
    assert R.total() == n and R.sum_amount() == oracle.sum_amount(rows)          # bit-equal to the serial id-order sum
    assert R.avg_amount() == oracle.avg_amount(rows)
    for lo, hi in ((100.0, 500.0), (0.0, 1.0), (999.0, 1e9), (500.0, 100.0)):
        assert R.sum_amount_where(lo, hi) == oracle.sum_amount_where(rows, lo, hi)[0]
    assert R.tree_height() == oracle.tree_height(n) and R.node_count() == oracle.node_count(n)
    order = None
    for m in DETERMINISTIC:
        if m == "memory_stride" and 255 <= n < 1000:
            continue                           # reference result depends on earlier calls there (stale subtree count)
        for p in (0.1, 1.0, 7.5, 20.0, 50.0, 100.0):
            for kw in ({}, {"num_threads": 3, "step_size": 3, "block_size": 1024 if m == "page" else 100,
                            "block_size_max": 5 if m == "stratified_block" else 900, "seed": 7}):
                if m == "dual_pointer" and int(n * p / 100.0) < 3:
                    continue                   # reference divides by zero
                prm = make_params(m, p, **kw)
                want = R.sample(m, prm)["id"] - 1
                got = oracle.indices(rows, m, prm)
                if m == "stratified_block":
                    if order is None:
                        order = np.argsort(rows["amount"], kind="stable")
                    got = order[got]
                assert len(got) == len(want) and np.array_equal(got, want), (n, m, p, kw)


def test_racy_clt_sampler_is_a_prefix_family_of_the_lockstep_one(oracle):
    """clt_validated_dual_pointer_sample is racy in the reference; every run returns, per thread, a prefix of the same
    stride sequence the lock-step schedule walks, and about as many rows."""
    rows = oracle.synth(200000, seed=5)
    R = Ref(rows)
    prm = make_params("clt_validated_dual_pointer", 20.0, max_error_percent=1.0)
    lock = oracle.indices(rows, "clt_validated_dual_pointer", prm)
    N, T = len(rows), int(len(rows) * 20 / 100.0)
    for _ in range(3):
        got = R.sample("clt_validated_dual_pointer", prm)["id"] - 1
        assert 0.5 * len(lock) <= len(got) <= 2.0 * len(lock) + T
        # each returned row lies on one of the four stride sequences (fast: a + 5k, slow: a + 2 + 5k; regions of N/2)
        half = N // 2
        on_seq = ((got % half) % 5 == 0) | ((got % half) % 5 == 2) | (got % max(1, N // (T // 4)) == 0)
        assert on_seq.all()


def test_scheduler_constants_match_reference(oracle):
    rows = oracle.synth(20000, seed=2)
    S = RefScheduler(rows)
    r = S.run(0)
    assert r.value == oracle.sum_amount(rows) and (r.confidence_level, r.error_margin, r.samples_used, r.status) == (1.0, 0.0, 20000, 0)
    for p, conf in ((10.0, 0.95), (3.0, 0.90), (1.0, 0.85), (0.3, 0.80), (0.1, 0.70)):
        q = S.run(3, "SELECT SUM(amount) FROM sales", p, 4)
        assert (q.confidence_level, q.error_margin, q.samples_used) == (conf, p / 100.0, int(20000 * p / 100.0))


# ---- SQL-string path: oracle/aqe_oracle_sql.c against the compiled executor.cpp + parser.cpp + core/db.cpp ----------------
def test_sql_restatement_against_compiled_reference(oracle, tmp_path):
    """Seeded random queries from the restated grammar, every entry point, random sample percentages, on a SQLite copy
    of a seeded table (signed amounts, ids with gaps).  Grouped calls for which the restatement predicts a throw inside
    one of the reference's worker threads are not issued: the reference process would abort."""
    import math

    import numpy as np

    from oracle import RefSql, SqlError
    if not RefSql.available():
        pytest.skip("oracle/_ref/libaqe_refsql.so not built")
    rng = np.random.default_rng(77)
    n = 3000
    rows = oracle.synth(n, seed=31)
    rows["amount"] = np.round(np.exp(rng.normal(3.0, 2.0, size=n)) * rng.choice([-1.0, 1.0], size=n, p=[0.15, 0.85]), 6)
    rows["id"] = np.sort(rng.choice(5 * n, size=n, replace=False)) + 1
    db = str(tmp_path / "t.db")
    RefSql.make_sqlite(db, rows)
    R = RefSql(db)
    t0 = 1700000000
    checked = 0
    for _ in range(150):
        agg = rng.choice(["SUM", "AVG", "COUNT", "sum"])
        col = rng.choice(["amount", "region", "product_id", "timestamp", "id"])
        terms = []
        for _k in range(rng.integers(0, 3)):
            a, b = sorted(rng.uniform(-100, 1500, size=2))
            k = int(rng.integers(-1, 9))
            terms.append(rng.choice([f"amount BETWEEN {a:.2f} AND {b:.2f}", f"amount >= {a:.1f}", f"region != {k}", f"region < {k}.5", f"'{k}' = region",
                                     f"product_id BETWEEN {int(a)} AND {int(b)}", f"timestamp > {t0 + int(a)}", f"(id <= {int(b) * 4})", f"rowid <> {int(a)}"]))
        group = rng.choice(["", "", " GROUP BY region", " GROUP BY product_id"])
        sql = f"SELECT {agg}({col}) FROM sales" + (" WHERE " + " AND ".join(terms) if terms else "") + group
        p = int(rng.choice([0, 1, 4, 10, 33, 50, 99, 100]))
        for mode in (("run_query_groupby", "run_query_groupby_with_ci") if group else ("run_query", "run_query_with_ci")):
            try:
                got = oracle.sql(rows, sql, p, mode)
            except SqlError as e:
                if group and (e.kind == "stod" or "integer overflow" in e.msg):
                    continue
                with pytest.raises(SqlError) as ei:
                    R.run(sql, p, mode)
                assert ei.value.kind == e.kind, (sql, p, mode)
                continue
            want = R.run(sql, p, mode)
            assert len(got) == len(want), (sql, p, mode)
            for g, w in zip(got, want):
                assert g[0] == w[0]
                for x, y in zip(g[1:], w[1:]):
                    assert x == y or (math.isnan(x) and math.isnan(y)) or abs(x - y) <= 4e-15 * max(abs(x), abs(y)) or abs(x - y) <= 1e-9 * abs(w[1]), (sql, p, mode, g, w)
            checked += 1
    assert checked > 150
