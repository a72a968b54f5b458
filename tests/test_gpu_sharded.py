"""The whole table behind ONE handle, range-sharded over several GPUs of this process (aqe_create_sharded; csrc/aqe_group.inl).

On a 1-GPU box the shards are colocated on device 0 (host merge); with >= 2 GPUs the same checks run in peer mode (the kernels
exchange through peer-mapped mailboxes) -- both must return what ONE GPU holding the whole table returns: counts, integer sums,
index sets and rows bit for bit, fp64 sums to <= 1 ulp (the partition of a compensated sum), and the oracle's answers under the
stated tolerances.  Reference: the region split custom_bplus_db.cpp:925-926, the per-thread sampler loops :814-854, :1218-1271,
:1880-1960."""
import math
import os

import numpy as np
import pytest

import approximatequeryengine_b200 as aqe
from oracle import make_params as orc_params

pytestmark = pytest.mark.gpu

N = 300_007


def n_gpus():
    c = aqe.C.c_int()
    aqe.lib().aqe_device_count(aqe.C.byref(c))
    return c.value


def layouts():
    out = [("colocated3", [0, 0, 0])]
    if n_gpus() >= 2:
        out.append(("peer%d" % min(n_gpus(), 8), list(range(min(n_gpus(), 8)))))
    return out


@pytest.fixture(scope="module")
def small_shards():
    old = os.environ.get("AQE_MIN_SHARD_ROWS")
    os.environ["AQE_MIN_SHARD_ROWS"] = "1"     # tables of a few hundred thousand rows are sharded too
    yield
    if old is None:
        os.environ.pop("AQE_MIN_SHARD_ROWS", None)
    else:
        os.environ["AQE_MIN_SHARD_ROWS"] = old


@pytest.fixture(scope="module")
def world(oracle, small_shards):
    rows = oracle.synth(N, seed=7)
    one = aqe.Engine(0).from_rows(rows)
    groups = [(name, aqe.Engine(devices=devs).from_rows(rows)) for name, devs in layouts()]
    return rows, one, groups


def ulps(a, b):
    return abs(a - b) / math.ulp(max(abs(a), abs(b), 1e-300))


def test_layout(world):
    rows, one, groups = world
    for name, g in groups:
        G = g.shard_count
        assert G == (3 if name.startswith("colocated") else min(n_gpus(), 8)) and g.count == N
        assert [g.shard_first_row(i) for i in range(G + 1)] == [N * i // G for i in range(G + 1)]   # custom_bplus_db.cpp:925-926
        assert g.fused == name.startswith("peer")
        assert g.read_rows().tobytes() == rows.tobytes()
        assert g.read_rows(N // 3 - 5, 11).tobytes() == rows[N // 3 - 5:N // 3 + 6].tobytes()
        assert np.array_equal(g.read_column("timestamp", N // 2 - 3, 9), rows["timestamp"][N // 2 - 3:N // 2 + 6])


def test_exact_scans_equal_one_gpu(world, oracle):
    rows, one, groups = world
    t0 = 1700000000
    cases = [("amount", None, 0, 0), ("amount", "amount", 100.0, 500.0), ("amount", "timestamp", t0 + 5000, t0 + 170000), ("amount", "region", 2, 5),
             ("id", None, 0, 0), ("id", "amount", 1.0, 300.0), ("timestamp", "region", 0, 0), ("region", None, 0, 0),
             ("product_id", "product_id", 100, 899), ("amount", "amount", 5000.0, 6000.0)]
    for name, g in groups:
        for agg, pred, lo, hi in cases:
            a, b = g.scan(agg, pred, lo, hi), one.scan(agg, pred, lo, hi)
            o = oracle.scan(rows, agg, pred, lo, hi)
            assert a.count == b.count == o.count and a.isum == b.isum, (name, agg, pred)
            if agg == "amount":
                assert ulps(a.sum, b.sum) <= 1 and (abs(a.sum - o.sum) <= 1e-12 * abs(o.sum) or o.count == 0), (name, agg, pred, a.sum, b.sum)
                if a.count:
                    assert (a.minv, a.maxv) == (b.minv, b.maxv)
            else:
                assert a.isum == o.isum and a.sum == b.sum
        assert g.sum_amount() == g.scan("amount").sum and g.sum_int("id") == N * (N + 1) // 2
        s, c = g.sum_amount_where(100.0, 500.0)
        assert (s, c) == (g.scan("amount", "amount", 100.0, 500.0).sum, one.sum_amount_where(100.0, 500.0)[1])
        assert g.scan("amount", "amount", 100.0, 500.0).sum == s     # run-to-run bit stable


METHOD_KW = {"fast_pointer": dict(step_size=3), "parallel_pointer": dict(num_threads=6), "random_pointer": dict(seed=9), "block": dict(block_size=700),
             "page": dict(block_size=4096), "parallel_block": dict(block_size=500, num_threads=3), "node_skip": dict(step_size=3),
             "adaptive_block": dict(block_size=300, block_size_max=900), "stratified_block": dict(block_size=400, block_size_max=5),
             "sample_records": dict(seed=5), "optimized_sequential": dict(seed=5), "random_start_nth": dict(step_size=7, seed=5),
             "address_arithmetic": dict(seed=5), "random_start_memory_stride": dict(seed=5), "multithreaded_memory_stride": dict(num_threads=5, seed=5),
             "clt_validated_dual_pointer": dict(max_error_percent=1.0), "optimized_clt": dict(num_threads=6), "signal_based_clt": dict(check_interval=20)}


def test_every_sampler_equals_one_gpu(world, oracle):
    """Config[3] 'block sampling + parallel fast/slow method' across shards: same index set, same rows, same moments."""
    rows, one, groups = world
    for name, g in groups:
        for method in aqe.METHODS:
            for pct in (1.0, 7.5):
                p = aqe.make_params(method, pct, **METHOD_KW.get(method, {}))
                pg, p1 = g.plan(method, p), one.plan(method, p)
                assert pg.count == p1.count and np.array_equal(pg.indices(), p1.indices()), (name, method, pct)
                want = oracle.indices(rows, method, orc_params(method, pct, **METHOD_KW.get(method, {})))
                if pg.by_amount_order:    # the oracle answers in positions of the amount-sorted table (custom_bplus_db.cpp:1343)
                    want = np.argsort(rows["amount"], kind="stable")[want]
                rg, r1 = g.gather(pg), one.gather(p1)
                assert rg.tobytes() == r1.tobytes(), (name, method, pct)
                assert np.array_equal(rg["id"] - 1, want)
                sg, s1 = g.stats(pg), one.stats(p1)
                assert sg.n == s1.n == len(want)
                if sg.n:
                    assert abs(sg.sum - s1.sum) <= 1e-12 * abs(s1.sum) and abs(sg.m2 - s1.m2) <= 1e-9 * max(s1.m2, 1e-300), (name, method, pct)
                    so = oracle.stats(rows, want)
                    assert abs(sg.sum - so.sum) <= 1e-12 * abs(so.sum) and abs(sg.m2 - so.m2) <= 1e-9 * max(so.m2, 1e-300)
        # sampled rows failing a predicate contribute 0 (parallel_sum_where_sample :317-343), another column, an index list
        p = g.plan("memory_stride", aqe.make_params("memory_stride", 5.0))
        a, b = g.stats(p, where=(100.0, 500.0)), one.stats(one.plan("memory_stride", aqe.make_params("memory_stride", 5.0)), where=(100.0, 500.0))
        assert a.n == b.n and abs(a.sum - b.sum) <= 1e-12 * abs(b.sum)
        a, b = g.stats(p, col="timestamp"), one.stats(one.plan("memory_stride", aqe.make_params("memory_stride", 5.0)), col="timestamp")
        assert a.n == b.n and abs(a.sum - b.sum) <= 1e-12 * abs(b.sum) and abs(a.m2 - b.m2) <= 1e-9 * b.m2
        idx = np.array([N - 1, 0, 5, N // 2, 5, N // 3, N // 3 - 1], dtype=np.int64)
        assert g.gather_indices(idx).tobytes() == rows[idx].tobytes()
        assert g.stats_from_indices(idx).n == len(idx) and abs(g.stats_from_indices(idx).sum - one.stats_from_indices(idx).sum) <= 1e-9
        fa, fb = g.fast_aggregated(aqe.make_params("multithreaded_memory_stride", 2.0, seed=5)), one.fast_aggregated(aqe.make_params("multithreaded_memory_stride", 2.0, seed=5))
        assert fa[1] == fb[1] and abs(fa[0] - fb[0]) <= 1e-12 * abs(fb[0])


def test_sql_equals_one_gpu(world):
    rows, one, groups = world
    queries = [("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10, "ci_reference"), ("SELECT AVG(amount) FROM sales GROUP BY region", 0, "value"),
               ("SELECT SUM(amount) FROM sales WHERE region >= 2 GROUP BY region", 20, "ci_reference"), ("SELECT SUM(amount) FROM sales GROUP BY product_id", 50, "value"),
               ("SELECT COUNT(*) FROM sales", 0, "value"), ("SELECT COUNT(amount) FROM sales WHERE id <= 400 GROUP BY product_id", 10, "value"),
               ("SELECT SUM(timestamp) FROM sales WHERE (region = 1 OR region = 5)", 7, "value"), ("SELECT SUM(amount) FROM sales WHERE region IN (1, 3, 5)", 0, "ci_correct")]
    for name, g in groups:
        for it in range(2):
            for sql, p, mode in queries:
                a = [(r.key, r.count, r.value, r.ci_lower, r.ci_upper, r.isum) for r in g.sql(sql, p, mode)]
                b = [(r.key, r.count, r.value, r.ci_lower, r.ci_upper, r.isum) for r in one.sql(sql, p, mode)]
                assert a == b, (name, sql)         # 128-bit fixed-point accumulators: shards merge exactly


def test_approx_over_shards(world):
    rows, one, groups = world
    truth = math.fsum(rows["amount"])
    sel = (rows["amount"] >= 100.0) & (rows["amount"] <= 500.0)
    tw, cw = math.fsum(rows["amount"][sel]), int(sel.sum())
    for name, g in groups:
        for agg, where, want in (("sum", None, truth), ("avg", None, truth / N), ("count", None, N), ("sum", (100.0, 500.0), tw),
                                 ("count", (100.0, 500.0), cw), ("avg", (100.0, 500.0), tw / cw)):
            hit = 0
            for seed in range(40):
                r = g.approx(agg, error_percent=1.0, seed=seed, where=where)
                assert r.status == 0 and r.population == N and r.error_margin <= 0.0100001, (name, agg, where)
                hit += r.ci_lower <= want <= r.ci_upper
            assert hit >= 34, (name, agg, where, hit)
        r = g.approx("sum", design="block", block_size=500, min_samples=64, seed=1)      # 600 tiles in all: the budget may run out (DRIFTING)
        assert r.status in (0, 1) and abs(r.estimate - truth) / truth < 0.03 and r.ci_lower <= r.estimate <= r.ci_upper


def test_files_appends_and_reload(world, oracle, tmp_path):
    rows, one, groups = world
    path = str(tmp_path / "t.aqe")
    oracle.save_file(path, rows)
    for name, devs in layouts():
        g = aqe.Engine(devices=devs).load_file(path)
        assert g.count == N and g.shard_count > 1 and g.read_rows().tobytes() == rows.tobytes()
        out = str(tmp_path / (name + ".aqe"))
        g.save_file(out)
        assert open(out, "rb").read() == open(path, "rb").read()          # byte-identical to save_to_file (custom_bplus_db.cpp:665-683)
        # unsorted input: one stable order by id across the shards (custom_bplus_db.cpp:198-200)
        perm = np.random.default_rng(3).permutation(N)
        g.from_rows(rows[perm])
        assert g.read_rows().tobytes() == rows.tobytes()
        shuffled = str(tmp_path / (name + "_shuffled.aqe"))
        oracle.save_file(shuffled, rows[perm])
        assert aqe.Engine(devices=devs).load_file(shuffled).read_rows().tobytes() == rows.tobytes()
        # a slice of the file, then appended rows (insert_record / insert_batch): re-sharded on the next query
        g.load_file(path, first_row=1000, n_rows=50_000)
        assert g.count == 50_000 and g.sum_int("id") == sum(range(1001, 51001))
        g.append(rows[:1000])
        assert g.count == 51_000 and g.sum_int("id") == 51_000 * 51_001 // 2 and g.read_rows(0, 3).tobytes() == rows[:3].tobytes()
        g.generate(123_457, seed=11)
        assert g.read_rows().tobytes() == oracle.synth(123_457, seed=11).tobytes()
        g.close()


def test_small_tables_stay_on_one_gpu(oracle):
    os.environ.pop("AQE_MIN_SHARD_ROWS", None)
    try:
        g = aqe.Engine(devices=[0, 0]).from_rows(oracle.synth(5000, seed=3))
        assert g.shard_count == 1 and g.sum_int("id") == 5000 * 5001 // 2
        g.generate(40_000_000, seed=7, columns=("amount",))                   # 2^24 rows per shard: two shards
        assert g.shard_count == 2 and g.scan("amount", "amount", 100.0, 500.0).count > 0
    finally:
        os.environ["AQE_MIN_SHARD_ROWS"] = "1"


def test_dropin_module_over_shards(world, oracle, tmp_path, small_shards):
    """The call sequence of enhanced_aqe_cli.py:165-186, 327-346 on a sharded CustomBPlusDB equals the one-GPU module."""
    rows, one, groups = world
    path = str(tmp_path / "sales.aqe")
    oracle.save_file(path, rows)
    b = aqe.backend()
    ref = b.CustomBPlusDB(0)
    assert ref.open_database(path) and ref.shard_count == 1
    for name, devs in layouts():
        db = b.CustomBPlusDB(devs)
        assert db.open_database(path) and db.get_total_records() == N and db.shard_count == len(devs)
        assert db.shards_fused == name.startswith("peer")
        assert ulps(db.sum_amount(), ref.sum_amount()) <= 1 and ulps(db.sum_amount_where(100.0, 500.0), ref.sum_amount_where(100.0, 500.0)) <= 1
        for call in (lambda d: d.memory_stride_sample(1.0, 0), lambda d: d.block_sample(2.0), lambda d: d.parallel_block_sample(2.0),
                     lambda d: d.parallel_pointer_sample(2.0), lambda d: d.clt_validated_dual_pointer_sample(20, 0.95, 10, 4, 1.0),
                     lambda d: d.stratified_block_sample(2.0), lambda d: d.adaptive_block_sample(2.0), lambda d: d.random_pointer_sample(0.5)):
            a, r = call(db), call(ref)
            assert [(x.id, x.amount, x.region, x.product_id, x.timestamp) for x in a] == [(x.id, x.amount, x.region, x.product_id, x.timestamp) for x in r]
        assert db.query_groupby("SELECT SUM(amount) FROM sales GROUP BY region") == ref.query_groupby("SELECT SUM(amount) FROM sales GROUP BY region")
        r = db.approx_avg(error_percent=1.0, seed=4)
        assert r.status == b.CustomApproximationStatus.STABLE and r.ci_lower <= ref.sum_amount() / N <= r.ci_upper
        s = b.CustomApproximateScheduler()
        assert s.open_database(path) and s.execute_exact_sum().value > 0
        db.close_database()


def test_plan_built_for_a_larger_table_is_refused(oracle):
    """a plan's segments expand against whatever table the handle holds -- used on a smaller table they read out of bounds."""
    rows = oracle.synth(20_000, seed=5)
    big, small = aqe.Engine(0).from_rows(rows), aqe.Engine(0).from_rows(rows[:10_000])
    p = big.plan("memory_stride", aqe.make_params("memory_stride", 5.0))
    assert big.stats(p).n == p.count
    for call in (lambda: small.stats(p), lambda: small.gather(p), lambda: small.stats(aqe.build_plan(20_000, "block", aqe.make_params("block", 5.0)))):
        with pytest.raises(aqe.AqeError) as ei:
            call()
        assert ei.value.code == 1
    assert small.stats(aqe.build_plan(10_000, "block", aqe.make_params("block", 5.0))).n > 0
    big.from_rows(rows[:5_000])                     # the table shrank under a plan built before
    with pytest.raises(aqe.AqeError):
        big.stats(p)


def test_negative_integer_totals_convert_exactly(oracle):
    """(double)hi * 2^64 + (double)lo loses a small negative total; the device converts through the magnitude."""
    rows = oracle.synth(1000, seed=1)
    rows["region"] = 0
    rows["region"][:5] = -1                          # total -5
    rows["timestamp"] = -(2**40) - np.arange(1000)
    for e in (aqe.Engine(0).from_rows(rows), aqe.Engine(devices=[0, 0]).from_rows(rows)):
        p = e.scan("region")
        assert p.isum == -5 and p.sum == -5.0
        p = e.scan("timestamp")
        want = int(rows["timestamp"].astype(object).sum())
        assert p.isum == want and p.sum == float(want)
    big = oracle.synth(300_000, seed=2)
    big["timestamp"] = -(2**62) + np.arange(300_000)             # |total| > 2^64: the 128-bit path
    want = int(big["timestamp"].astype(object).sum())
    for e in (aqe.Engine(0).from_rows(big), aqe.Engine(devices=[0, 0, 0]).from_rows(big)):
        p = e.scan("timestamp")
        assert p.isum == want and p.sum == float(want)
