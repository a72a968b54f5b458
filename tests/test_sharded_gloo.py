"""The N>1 host path on CPU: world_size-2 (and 3) gloo process groups.  Each rank computes the partial of its
contiguous shard with the ORACLE (no GPU here), the product's host layer (approximatequeryengine_b200.sharded)
all-gathers the 64-byte partials and folds them in rank order with the C-ABI's aqe_merge_partials /
aqe_approx_merge, and every rank must end up with the oracle's whole-table answer."""
import math
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    import approximatequeryengine_b200 as aqe
    from approximatequeryengine_b200 import sharded
    from oracle import ApproxSpec, Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = Oracle()
        a, b = sharded.shard_range(n, rank, world)
        rows = O.synth(b - a, seed=7, first_row=a)          # this rank's shard of the synthetic table
        out = {}
        for name, (agg, pred, lo, hi) in {"sum": ("amount", None, 0, 0), "where": ("amount", "amount", 100.0, 500.0),
                                          "ts": ("timestamp", "region", 1, 6), "id": ("id", None, 0, 0)}.items():
            o = O.scan(rows, agg, pred, lo, hi)
            local = aqe.Partial(count=o.count, sum=o.sum, comp=0.0, isum_lo=o.isum_lo, isum_hi=o.isum_hi, sumsq=o.sumsq, minv=o.minv, maxv=o.maxv)
            parts = sharded.allgather_struct(local, aqe.Partial)
            assert len(parts) == world and parts[rank].count == o.count and parts[rank].sum == o.sum
            m = sharded.merge_partials(parts, is_integer=agg != "amount")
            out[name] = (m.count, m.sum, m.comp, m.isum, m.minv, m.maxv)
        # sampled estimates: shards are strata
        r = O.approx(rows, ApproxSpec(agg=0, design=0, agg_col=1, pred_col=-1, error_percent=1.0, confidence_level=0.95, seed=(5 << 8) + rank))
        local = aqe.ApproxResult.from_buffer_copy(bytes(r))
        parts = sharded.allgather_struct(local, aqe.ApproxResult)
        m = sharded.merge_approx(parts, "sum", 0.95)
        out["approx"] = (m.estimate, m.ci_lower, m.ci_upper, m.n_samples, m.population, m.status)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_over_gloo(oracle, world):
    n = 300_007
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = oracle.synth(n, seed=7)
    assert all(got[r] == got[0] for r in range(world))       # every rank computed identical bits
    g = got[0]
    whole = oracle.scan(rows, "amount")
    assert g["sum"][0] == n and abs(g["sum"][1] - math.fsum(rows["amount"])) <= 1e-12 * whole.sum
    assert (g["sum"][4], g["sum"][5]) == (whole.minv, whole.maxv)
    w = oracle.scan(rows, "amount", "amount", 100.0, 500.0)
    assert g["where"][0] == w.count and abs(g["where"][1] - w.sum) <= 1e-12 * w.sum
    assert g["ts"][3] == oracle.scan(rows, "timestamp", "region", 1, 6).isum and g["ts"][0] == oracle.scan(rows, "timestamp", "region", 1, 6).count
    assert g["id"][3] == n * (n + 1) // 2
    est, lo, hi, ns, pop, status = g["approx"]
    assert pop == n and status == 0 and lo < est < hi
    assert abs(est - whole.sum) / whole.sum < 0.02 and (hi - lo) / 2 / est <= 0.011


def test_shard_ranges_tile_the_table():
    from approximatequeryengine_b200.sharded import shard_range
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


class _HostShard:
    """Stands in for an Engine on a box without a GPU: column facts and the grouped-scan accumulators of this rank's
    rows come from numpy (tests/sql_helpers.py); everything downstream is the product's host layer."""

    def __init__(self, rows):
        self.rows = rows

    def sql_facts(self, q):
        import approximatequeryengine_b200 as aqe
        from sql_helpers import COL_NAMES
        f = aqe.SqlFacts()
        f.key_min, f.key_max = (0, 0) if len(self.rows) else (0, -1)
        if q.group_col >= 0 and len(self.rows):
            c = self.rows[COL_NAMES[q.group_col]]
            f.key_min, f.key_max = int(c.min()), int(c.max())
        if q.agg_col >= 0:
            f.agg_is_integer = int(q.agg_col != 1)
            if len(self.rows):
                c = self.rows[COL_NAMES[q.agg_col]]
                f.agg_absmax = float(max(abs(float(c.min())), abs(float(c.max()))))
        return f

    def sql_scan(self, q, layout, flags=0):
        from sql_helpers import emulate_scan
        return emulate_scan(self.rows, q, layout, flags)


SQL_CASES = [("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10, "run_query_with_ci"),
             ("SELECT AVG(amount) FROM sales GROUP BY region", 0, "run_query_groupby"),
             ("SELECT SUM(amount) FROM sales WHERE region >= 2 GROUP BY region", 20, "run_query_groupby_with_ci"),
             ("SELECT COUNT(amount) FROM sales WHERE id <= 400 GROUP BY product_id", 10, "run_query_groupby"),
             ("SELECT SUM(timestamp) FROM sales", 7, "run_query")]


def _sql_worker(rank, world, port, n, q):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from approximatequeryengine_b200 import sharded
    from oracle import Oracle
    from sql_helpers import MODE_OF, engine_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharded.shard_range(n, rank, world)
        t = sharded.ShardedTable(_HostShard(Oracle().synth(b - a, seed=7, first_row=a)), n, a)
        out = []
        for sql, p, mode in SQL_CASES:
            try:
                out.append(engine_rows(t.sql(sql, p, MODE_OF[mode])))
            except ValueError:
                out.append("stod")
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sql_over_gloo(oracle, world):
    """SQL path across ranks: facts / accumulators / existence counts travel over gloo, every rank ends with the
    oracle's whole-table answer and all ranks agree bit for bit."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import SqlError
    from sql_helpers import REL, rows_close
    n = 20_011
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sql_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = oracle.synth(n, seed=7)
    for i, (sql, p, mode) in enumerate(SQL_CASES):
        assert all(got[r][i] == got[0][i] for r in range(world)), sql
        try:
            want = oracle.sql(rows, sql, p, mode)
        except SqlError as e:
            assert e.kind == "stod" and got[0][i] == "stod", (sql, e)
            continue
        assert rows_close(got[0][i], want, REL) is None, (sql, rows_close(got[0][i], want, REL))


# ---------------------------------------------------------------------------------------------------------------------------
# legacy sampler families across ranks (BASELINE configs[3]) and the AVG ... WHERE merge of independent shard estimates
# ---------------------------------------------------------------------------------------------------------------------------
class _HostSamplerShard:
    """Stands in for an Engine on a box without a GPU: the window filter and partial sums of k_plan_stats / k_plan_gather in numpy
    (positions come from the product's own host plan expansion); everything downstream is the product's host layer."""

    def __init__(self, rows):
        self.rows = rows

    def _local(self, plan, window_first):
        pos = plan.indices()
        k = np.nonzero((pos >= window_first) & (pos < window_first + len(self.rows)))[0]
        return k, pos[k] - window_first

    def stats_window(self, plan, window_first, col="amount", where=None, where_col="amount"):
        import approximatequeryengine_b200 as aqe
        _, loc = self._local(plan, window_first)
        x = self.rows[col][loc].astype(np.float64)
        if where is not None:
            pv = self.rows[where_col][loc].astype(np.float64)
            x = np.where((pv >= where[0]) & (pv <= where[1]), x, 0.0)
        K = float(self.rows[col][0]) if len(self.rows) else 0.0
        d = x - K
        return aqe.StatsPartial(n=len(x), sum=math.fsum(x), sum_c=0.0, shift=K, sd=math.fsum(d), sd_c=0.0, sdd=math.fsum(d * d), sdd_c=0.0)

    def gather_window(self, plan, window_first, k_first=0, k_count=None):
        import approximatequeryengine_b200 as aqe
        out = np.zeros(plan.count, dtype=aqe.RECORD_DTYPE)
        k, loc = self._local(plan, window_first)
        out[k] = self.rows[loc]
        return out, len(k)


SAMPLER_CASES = [("parallel_pointer", 2.0, dict(num_threads=6)), ("parallel_block", 3.0, dict(block_size=500, num_threads=3)), ("block", 1.0, {}),
                 ("memory_stride", 1.0, {}), ("optimized_clt", 5.0, dict(num_threads=5)), ("sample_records", 2.5, dict(seed=11)),
                 ("address_arithmetic", 1.5, dict(seed=4)), ("random_pointer", 0.5, dict(seed=9)), ("index_based", 1.0, {}),
                 ("multithreaded_memory_stride", 2.0, dict(num_threads=5, seed=5))]


def _sampler_worker(rank, world, port, n, q):
    import hashlib

    import torch.distributed as dist

    import approximatequeryengine_b200 as aqe
    from approximatequeryengine_b200 import sharded
    from oracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharded.shard_range(n, rank, world)
        t = sharded.ShardedTable(_HostSamplerShard(Oracle().synth(b - a, seed=7, first_row=a)), n, a)
        out = []
        for method, pct, kw in SAMPLER_CASES:
            p = aqe.make_params(method, pct, **kw)
            s = t.stats(method, p)
            w = t.stats(method, p, where=(100.0, 500.0))
            rows = t.gather(method, p)
            out.append((s.n, s.sum, s.mean, s.m2, w.n, w.sum, hashlib.sha256((rows["id"] - 1).astype("<i8").tobytes()).hexdigest(),
                        hashlib.sha256(rows.tobytes()).hexdigest()))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_samplers_over_gloo(oracle, world):
    """n and the index-set hash identical to the one-table plan, sums to 1e-12, on every rank alike."""
    import hashlib

    from oracle import make_params as orc_params
    n = 100_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = oracle.synth(n, seed=7)
    for i, (method, pct, kw) in enumerate(SAMPLER_CASES):
        assert all(got[r][i] == got[0][i] for r in range(world)), method
        cnt, s, mean, m2, wn, wsum, idx_sha, rows_sha = got[0][i]
        idx = oracle.indices(rows, method, orc_params(method, pct, **kw))
        o = oracle.stats(rows, idx)
        assert cnt == len(idx) == o.n == wn and idx_sha == hashlib.sha256(idx.astype("<i8").tobytes()).hexdigest(), method
        assert rows_sha == hashlib.sha256(rows[idx].tobytes()).hexdigest(), method
        assert abs(s - o.sum) <= 1e-12 * abs(o.sum) and abs(mean - o.mean) <= 1e-12 * abs(o.mean) and abs(m2 - o.m2) <= 1e-9 * o.m2, method
        x = rows["amount"][idx]
        assert abs(wsum - math.fsum(x[(x >= 100.0) & (x <= 500.0)])) <= 1e-12 * abs(wsum)


def test_stats_merge_is_exact_about_differing_shifts():
    """aqe_stats_merge (host code): shards with different shifts K and sizes fold to the moments of the union."""
    import approximatequeryengine_b200 as aqe
    rng = np.random.default_rng(2)
    parts, allx = [], []
    for n, loc in ((1000, 5.0), (0, 0.0), (37, 1e6), (50000, -300.0), (1, 42.0)):
        x = rng.normal(loc, 10.0, size=n)
        K = float(x[0]) if n else 0.0
        d = x - K
        parts.append(aqe.StatsPartial(n=n, sum=math.fsum(x), sum_c=0.0, shift=K, sd=math.fsum(d), sd_c=0.0, sdd=math.fsum(d * d), sdd_c=0.0))
        allx.append(x)
    x = np.concatenate(allx)
    s = aqe.merge_stats(parts)
    mean = math.fsum(x) / len(x)
    assert s.n == len(x) and abs(s.sum - math.fsum(x)) <= 1e-12 * abs(math.fsum(x)) and abs(s.mean - mean) <= 1e-12 * abs(mean)
    assert abs(s.m2 - math.fsum((x - mean) ** 2)) <= 1e-9 * s.m2
    assert aqe.merge_stats([]).n == 0 and aqe.merge_stats(parts[1:2]).n == 0


def _avg_where_worker(rank, world, port, n, q):
    import torch.distributed as dist

    import approximatequeryengine_b200 as aqe
    from approximatequeryengine_b200 import sharded
    from oracle import ApproxSpec, Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = Oracle()
        a, b = sharded.shard_range(n, rank, world)
        rows = _skewed_rows(O, n)[a:b]
        out = []
        for seed in range(12):
            # every rank's own estimate of AVG(amount) WHERE timestamp <= cut (the restated persistent kernel), merged as strata
            r = O.approx(rows, ApproxSpec(agg=1, design=0, agg_col=1, pred_col=4, lo=0.0, hi=float(1700000000 + n * 6 // 10), error_percent=1.0,
                                          confidence_level=0.95, seed=(seed << 8) + rank, min_samples=4096))
            parts = sharded.allgather_struct(aqe.ApproxResult.from_buffer_copy(bytes(r)), aqe.ApproxResult)
            m = sharded.merge_approx(parts, "avg", 0.95)
            out.append((m.estimate, m.ci_lower, m.ci_upper, m.pass_fraction, [p.pass_fraction for p in parts]))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def _skewed_rows(O, n):
    rows = O.synth(n, seed=7)
    rows["amount"][n // 2:] += 2000.0          # the second shard's matching rows are few AND much larger
    return rows


def test_avg_where_merge_weights_by_matching_rows(oracle):
    """AVG ... WHERE over shards of different selectivity.  `timestamp <= cut` keeps all of shard 0 and a fifth of
    shard 1; weighting the shard means by shard size (the old merge) lands near the midpoint of the two means, far outside the
    interval; weighting by the estimated matching rows (population x pass_fraction) reproduces the table-level conditional mean."""
    n, world = 400_000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_avg_where_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]
    rows = _skewed_rows(oracle, n)
    sel = rows["timestamp"] <= 1700000000 + n * 6 // 10
    truth = math.fsum(rows["amount"][sel]) / int(sel.sum())
    wrong = (rows["amount"][:n // 2][sel[:n // 2]].mean() + rows["amount"][n // 2:][sel[n // 2:]].mean()) / 2     # what equal shard weights give
    hits = 0
    for est, lo, hi, pf, pfs in got[0]:
        assert abs(pfs[0] - 1.0) < 1e-12 and abs(pfs[1] - 0.2) < 0.03 and abs(pf - sel.mean()) < 0.02
        assert abs(est - truth) / truth < 0.04 and abs(est - wrong) / wrong > 0.2 and (hi - lo) / 2 / est < 0.05
        hits += lo <= truth <= hi
    assert hits >= 10
