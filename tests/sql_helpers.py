"""Shared helpers of the SQL-path tests: golden access, comparison rules, and a numpy stand-in for the grouped-scan
kernel (so the host half -- parser, WHERE compilation, layout, merge, finish -- is testable without a GPU)."""
import glob
import json
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Relative tolerance on floating-point results of the SQL path.  The reference reads every number back from SQLite
# as TEXT with 15 significant digits (and SQLite's REAL->TEXT is not always correctly rounded), so two correct
# evaluations can differ by a few units in the 15th digit; north_star's fp64 gate is 1e-12.
REL = 1e-12
REL_ORACLE = 4e-15   # restatement vs reference: same arithmetic, same 15-digit text round trip


def sql_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "sql_n*_s*.json")))


def load(path):
    with open(path) as f:
        return json.load(f)


def golden_rows(case):
    out = []
    for r in case["rows"]:
        v = float.fromhex(r[1])
        lo, hi = (float.fromhex(r[2]), float.fromhex(r[3])) if len(r) == 4 else (v, v)
        out.append((r[0], v, lo, hi))
    return out


def close(a, b, rel, scale=None):
    if a == b or (math.isnan(a) and math.isnan(b)):
        return True
    if math.isinf(a) or math.isinf(b) or math.isnan(a) or math.isnan(b):
        return False
    ref = max(abs(a), abs(b)) if scale is None else max(abs(a), abs(b), scale)
    return abs(a - b) <= rel * ref


def rows_close(got, want, rel):
    """[(key, value, lo, hi)] lists agree: keys exact, value to `rel`; interval ends to `rel` of the value's
    magnitude plus the cancellation the reference's own (sum_sq - sum^2/n) suffers from its 15-digit inputs."""
    if len(got) != len(want):
        return f"{len(got)} rows, want {len(want)}"
    for g, w in zip(got, want):
        if g[0] != w[0]:
            return f"key {g[0]} != {w[0]}"
        if not close(g[1], w[1], rel):
            return f"key {g[0]}: value {g[1]!r} != {w[1]!r}"
        for i in (2, 3):
            if not close(g[i], w[i], rel, scale=abs(w[1])) and not close(g[i] - g[1], w[i] - w[1], 1e-6):
                return f"key {g[0]}: bound {g[i]!r} != {w[i]!r}"
    return None


COL_NAMES = ["id", "amount", "region", "product_id", "timestamp"]


def emulate_scan(rows, q, layout, flags=0):
    """What k_sql_agg leaves in its accumulators, computed with numpy / Python integers from the compiled query."""
    import approximatequeryengine_b200 as aqe
    G = layout.n_groups
    acc = np.zeros(G * 5, dtype=np.uint64)
    if q.always_false or len(rows) == 0:
        return acc
    unsampled = bool(flags & aqe.SQL_UNSAMPLED)
    moments = bool(flags & aqe.SQL_MOMENTS) and not unsampled
    mask = np.ones(len(rows), dtype=bool) if q.n_alt == 0 else np.zeros(len(rows), dtype=bool)
    for branch in q.branches():
        m = np.ones(len(rows), dtype=bool)
        for t in branch:
            col = rows[COL_NAMES[t.col]]
            if t.col == 1:
                m &= (col >= t.lo) & (col <= t.hi)
                if t.has_ne:
                    m &= col != t.ne
            else:
                c = col.astype(np.int64)
                m &= (c >= t.ilo) & (c <= t.ihi)
                if t.has_ne:
                    m &= c != t.ine
        mask |= m
    p = q.sample_percent
    step = 0 if (unsampled or p <= 0 or p >= 100) else max(1, 100 // p)
    if step > 1:
        mask &= (rows["id"] % step) == 0
    sums = q.agg_col >= 0 and not unsampled and (q.agg != 2 or moments)
    grp = np.zeros(len(rows), dtype=np.int64) if q.group_col < 0 else rows[COL_NAMES[q.group_col]].astype(np.int64) - layout.key_min
    sel = np.nonzero(mask)[0]
    g = grp[sel]
    assert ((g >= 0) & (g < G)).all()
    cnt = np.bincount(g, minlength=G)
    M = (1 << 64) - 1
    if sums:
        x = rows[COL_NAMES[q.agg_col]][sel]
        if q.agg_col == 1:
            fx = np.rint(x * math.ldexp(1.0, layout.sum_shift)).astype(np.int64)
            d = x
        else:
            fx = x.astype(np.int64)
            d = x.astype(np.float64)
        fq = np.rint((d * d) * math.ldexp(1.0, layout.sq_shift)).astype(np.int64) if moments else None
    for k in range(G):
        acc[5 * k] = int(cnt[k])
        if sums and cnt[k]:
            m = g == k
            s = sum(int(v) for v in fx[m])
            acc[5 * k + 1] = s & M
            acc[5 * k + 2] = (s >> 64) & M
            if moments:
                s2 = sum(int(v) for v in fq[m])
                acc[5 * k + 3] = s2 & M
                acc[5 * k + 4] = (s2 >> 64) & M
    return acc


def host_execute(rows, sql, p, mode):
    """aqe_sql_execute with the kernel replaced by emulate_scan: parse -> facts -> layout -> scan -> finish."""
    import approximatequeryengine_b200 as aqe
    q = aqe.sql_parse(sql, p)
    f = aqe.SqlFacts()
    f.key_min, f.key_max = (0, 0) if len(rows) else (0, -1)
    if q.group_col >= 0 and len(rows):
        c = rows[COL_NAMES[q.group_col]]
        f.key_min, f.key_max = int(c.min()), int(c.max())
    if q.agg_col >= 0:
        f.agg_is_integer = int(q.agg_col != 1)
        if len(rows):
            c = rows[COL_NAMES[q.agg_col]]
            f.agg_absmax = float(max(abs(float(c.min())), abs(float(c.max()))))
    L = aqe.sql_layout(q, [f])
    grouped = q.group_col >= 0
    step = 0 if (p <= 0 or p >= 100) else max(1, 100 // p)
    if mode == "value":
        moments = False
    elif grouped:
        moments = mode == "ci_reference" or q.agg != 2
    else:
        moments = q.agg != 2 and step > 0
    acc = emulate_scan(rows, q, L, aqe.SQL_MOMENTS if moments else 0)
    exists = None
    if grouped and step > 1 and any(acc[5 * k] == 0 for k in range(L.n_groups)):
        exists = emulate_scan(rows, q, L, aqe.SQL_UNSAMPLED)
    return aqe.sql_finish(q, L, acc, mode, exists)


MODE_OF = {"run_query": "value", "run_query_with_ci": "ci_reference", "run_query_groupby": "value",
           "run_query_groupby_with_ci": "ci_reference"}


def engine_rows(sql_rows):
    """aqe SqlRow list -> [(key, value, lo, hi)], raising the error the binding raises for a NULL aggregate."""
    for r in sql_rows:
        if r.is_null:
            raise ValueError("stod")
    return [(r.key, r.value, r.ci_lower, r.ci_upper) for r in sql_rows]


def wide_range_rows(oracle, n=20000, seed=77):
    """amount spans 13 decades: half the rows near 1e-6, half up to 1e6."""
    rng = np.random.default_rng(seed)
    rows = oracle.synth(n, seed=seed)
    rows["amount"] = np.where(rng.random(n) < 0.5, 10.0 ** rng.uniform(-7, -5, n), 10.0 ** rng.uniform(3, 6, n))
    return rows


WIDE_RANGE_QUERIES = (("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 0 AND 0.00001", 0, "run_query"),
                      ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 0 AND 0.00001", 10, "run_query_with_ci"),
                      ("SELECT AVG(amount) FROM sales WHERE amount > 0 AND amount <= 0.000002 GROUP BY region", 0, "run_query_groupby"),
                      ("SELECT SUM(amount) FROM sales WHERE (amount BETWEEN 0 AND 0.000001 OR amount BETWEEN 0.000005 AND 0.00001) GROUP BY region", 20,
                       "run_query_groupby_with_ci"))



def signed_rows(oracle, n=20000, seed=91):
    """Negative amounts and keys: amount ~ U(-1000, 1000), product_id in -2048..2047 (exactly SQL_MAX_GROUPS keys), region in -3..4."""
    rng = np.random.default_rng(seed)
    rows = oracle.synth(n, seed=seed)
    rows["amount"] = rng.uniform(-1000.0, 1000.0, n)
    rows["product_id"] = rng.integers(-2048, 2048, n)
    rows["product_id"][:2] = (-2048, 2047)
    rows["region"] = rng.integers(-3, 5, n)
    return rows


SIGNED_QUERIES = (("SELECT SUM(amount) FROM sales", 0, "run_query"), ("SELECT AVG(amount) FROM sales WHERE amount < 0", 10, "run_query_with_ci"),
                  ("SELECT SUM(amount) FROM sales WHERE region < 0 GROUP BY region", 0, "run_query_groupby_with_ci"),
                  ("SELECT SUM(amount) FROM sales WHERE product_id BETWEEN -5 AND 5 GROUP BY region", 20, "run_query_groupby_with_ci"),
                  ("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, "run_query_groupby"),
                  ("SELECT AVG(amount) FROM sales WHERE amount > -900.5 GROUP BY product_id", 50, "run_query_groupby"),
                  ("SELECT COUNT(*) FROM sales WHERE region != -1 GROUP BY product_id", 0, "run_query_groupby"),
                  ("SELECT SUM(product_id) FROM sales WHERE amount <= -1 GROUP BY region", 0, "run_query_groupby"),
                  ("SELECT AVG(region) FROM sales WHERE region IN (-3, -1, 4)", 25, "run_query_with_ci"),
                  ("SELECT SUM(amount) FROM sales WHERE region IN (-3, 0, 4) AND amount > 0", 0, "run_query"),
                  ("SELECT COUNT(amount) FROM sales WHERE region NOT IN (-1, 2) AND product_id < 100 GROUP BY region", 10, "run_query_groupby"),
                  ("SELECT SUM(amount) FROM sales WHERE region != -3 AND region != 0 AND region != 4 AND region != 3", 34, "run_query_with_ci"),
                  ("SELECT SUM(amount) FROM sales WHERE region IN (5, 6)", 0, "run_query_with_ci"))
