#!/usr/bin/env python3
"""Timings of the SQL-string path (k_sql_agg) on a device-generated table: ms per query (median of reps, host wall clock
around the synchronous C-ABI call aqe_sql_run), rows/s and achieved GB/s against the ALGORITHMIC bytes of the query (the widths of
the distinct columns it reads x rows visited).  python tools/sql_bench.py [rows] [reps] [sampled|or|groups] > out.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approximatequeryengine_b200 as aqe  # noqa: E402

WIDTH = {"id": 8, "amount": 8, "region": 4, "product_id": 4, "timestamp": 8}
T0 = 1700000000


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    e = aqe.Engine(0).generate(n, seed=7)
    cases = [
        ("SELECT SUM(amount) FROM sales", 0, "value", ["amount"]),
        ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 0, "value", ["amount"]),
        ("SELECT COUNT(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 0, "value", ["amount"]),
        ("SELECT SUM(timestamp) FROM sales WHERE region = 3", 0, "value", ["timestamp", "region"]),
        ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "value", ["amount", "region"]),
        ("SELECT COUNT(amount) FROM sales GROUP BY region", 0, "value", ["region"]),
        (f"SELECT AVG(amount) FROM sales WHERE timestamp BETWEEN {T0 + n // 4} AND {T0 + n // 2} GROUP BY region", 0, "value", ["amount", "region", "timestamp"]),
        ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "ci_reference", ["amount", "region"]),
        ("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, "value", ["amount", "product_id"]),
        ("SELECT COUNT(amount) FROM sales GROUP BY product_id", 0, "value", ["product_id"]),
        ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY product_id", 0, "value", ["amount", "product_id"]),
        ("SELECT SUM(amount) FROM sales", 50, "ci_reference", ["amount"]),
        ("SELECT SUM(amount) FROM sales", 10, "ci_reference", ["amount"]),
        ("SELECT SUM(amount) FROM sales", 1, "ci_reference", ["amount"]),
        ("SELECT AVG(amount) FROM sales GROUP BY region", 10, "ci_reference", ["amount", "region"]),
        ("SELECT SUM(amount) FROM sales GROUP BY product_id", 1, "value", ["amount", "product_id"]),
    ]
    if len(sys.argv) > 3 and sys.argv[3] == "sampled":   # small steps: the ring's row-number filter vs the strided visit (AQE_SQL_VARIANT=1)
        shapes = [("SELECT SUM(amount) FROM sales", "ci_reference", ["amount"]),
                  ("SELECT SUM(amount) FROM sales", "value", ["amount"]),
                  ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", "ci_reference", ["amount"]),
                  ("SELECT SUM(timestamp) FROM sales WHERE region = 3", "value", ["timestamp", "region"]),
                  ("SELECT AVG(amount) FROM sales GROUP BY region", "ci_reference", ["amount", "region"]),
                  ("SELECT SUM(amount) FROM sales GROUP BY region", "value", ["amount", "region"]),
                  (f"SELECT AVG(amount) FROM sales WHERE timestamp BETWEEN {T0 + n // 4} AND {T0 + n // 2} GROUP BY region", "value", ["amount", "region", "timestamp"]),
                  ("SELECT SUM(amount) FROM sales GROUP BY product_id", "value", ["amount", "product_id"])]
        cases = [(sql, p, mode, cols) for p in (50, 33, 25, 20, 15) for sql, mode, cols in shapes]
    if len(sys.argv) > 3 and sys.argv[3] == "or":   # cost of OR branches
        cases = [(f"SELECT {agg} FROM sales WHERE {w}", 0, "value", cols) for agg, cols in (("SUM(amount)", ["amount", "region"]), ("COUNT(*)", ["region"]))
                 for w in ("region = 1", "(region = 1 OR region = 3)", "region IN (1, 3, 5, 7)", "region IN (0, 1, 2, 3, 4, 5, 6, 7)", "region != 1 AND region != 3",
                           "(region = 1 OR amount > 900)", "(region = 1 AND amount > 900 OR region = 3 AND amount < 100)", "product_id = 5", "product_id IN (1, 3, 5, 7)",
                           "product_id NOT IN (1, 3, 5, 7, 9, 11)", "amount NOT BETWEEN 100 AND 500", "(amount < 50 OR amount > 950 OR amount BETWEEN 400 AND 410)",
                           f"timestamp NOT BETWEEN {T0 + n // 4} AND {T0 + n // 2}")]
    if len(sys.argv) > 3 and sys.argv[3] == "groups":   # mid-sized GROUP BY: the shared-atomic bins (A/B with AQE_SQL_PACKED=0)
        cases = [(f"SELECT {agg} FROM sales{w} GROUP BY product_id", 0, mode, cols)
                 for agg, mode, cols in (("SUM(amount)", "value", ["amount", "product_id"]), ("AVG(amount)", "value", ["amount", "product_id"]),
                                         ("SUM(amount)", "ci_reference", ["amount", "product_id"]), ("SUM(timestamp)", "value", ["timestamp", "product_id"]),
                                         ("COUNT(amount)", "value", ["product_id"]))
                 for w in ("", " WHERE amount BETWEEN 100 AND 500", " WHERE amount > 900", " WHERE region = 3")]
    out = []
    import ctypes as C
    buf = (aqe.SqlRow * aqe.SQL_MAX_GROUPS)()
    ngot = C.c_uint32()

    def call(sql, p, mode):   # the C-ABI call alone: materialising 1000 ctypes rows in Python costs ~0.3 ms
        aqe.check(e.L.aqe_sql_run(e.h, sql.encode(), p, aqe.SQL_MODE[mode], buf, aqe.SQL_MAX_GROUPS, C.byref(ngot)))
        return range(ngot.value)
    for sql, p, mode, cols in cases:
        call(sql, p, mode)  # warm (column statistics are computed on first use)
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            rows = call(sql, p, mode)
            ts.append(time.perf_counter() - t)
        ts.sort()
        ms = ts[len(ts) // 2] * 1e3
        step = 1 if p <= 0 or p >= 100 else max(1, 100 // p)
        visited = n // step
        # a strided visit of 8-byte values touches one 32-byte sector per row once step >= 4
        alg = sum(WIDTH[c] for c in cols) * visited
        out.append({"sql": sql, "p": p, "mode": mode, "ms": round(ms, 4), "rows_per_s": n / (ms / 1e3), "visited_rows": visited,
                    "algorithmic_GB": alg / 1e9, "GBps": alg / 1e9 / (ms / 1e3), "groups": len(rows)})
        print(f"{ms:9.3f} ms  {alg / 1e9 / (ms / 1e3):8.1f} GB/s  p={p:<3} {mode:13} {sql}", file=sys.stderr)
    json.dump({"rows": n, "reps": reps, "cases": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
