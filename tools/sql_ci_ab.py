#!/usr/bin/env python3
"""Grouped scans over thread-private bins (GROUP BY region, with and without the squares of the reference's run_query_groupby_with_ci),
1 B rows: median ms per query under each setting of AQE_SQL_PAIR_BINS (1: the branch-free row add, 2: the same bins through
SqlBins::add, 0: unpaired words).  python tools/sql_ci_ab.py [rows] [reps] > out.json"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = 1700000000


def child(n, reps):
    import approximatequeryengine_b200 as aqe
    e = aqe.Engine(0).generate(n, seed=7)
    out = []
    for sql, mode in (("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", "ci_reference"),
                      ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", "value"),
                      ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY region", "ci_reference"),
                      ("SELECT AVG(amount) FROM sales WHERE product_id < 500 GROUP BY region", "ci_reference"),
                      ("SELECT SUM(timestamp) FROM sales WHERE amount > 1 GROUP BY region", "ci_reference"),
                      (f"SELECT AVG(amount) FROM sales WHERE timestamp BETWEEN {T0 + n // 4} AND {T0 + n // 2} GROUP BY region", "value"),
                      ("SELECT SUM(amount) FROM sales GROUP BY region", "ci_reference"),
                      ("SELECT SUM(amount) FROM sales GROUP BY region", "value"),
                      ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", "value"),
                      ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", "ci_reference")):
        e.sql(sql, 0, mode)
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            r = e.sql(sql, 0, mode)
            ts.append(time.perf_counter() - t)
        ts.sort()
        out.append({"sql": sql, "mode": mode, "ms": ts[len(ts) // 2] * 1e3, "first_group": [int(r[0].key), int(r[0].count), r[0].sum, r[0].sumsq]})
    print(json.dumps(out))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    if os.environ.get("AQE_CI_AB_CHILD"):
        return child(n, reps)
    res = {}
    for pb in ("1", "2", "0"):
        env = dict(os.environ, AQE_SQL_PAIR_BINS=pb, AQE_CI_AB_CHILD="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(reps)], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            res[pb] = {"error": r.stderr[-2000:]}
            continue
        res[pb] = json.loads(r.stdout.strip().splitlines()[-1])
    for pb, rows in res.items():
        print(f"== AQE_SQL_PAIR_BINS={pb}", file=sys.stderr)
        if isinstance(rows, dict):
            print(rows["error"], file=sys.stderr)
            continue
        for q in rows:
            print(f"  {q['ms']:8.3f} ms  {q['mode']:13s} {q['sql'][:100]}  -> {q['first_group']!r}", file=sys.stderr)
    print(json.dumps({"rows": n, "reps": reps, "by_pair_bins": res}))


if __name__ == "__main__":
    main()
