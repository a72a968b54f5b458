#!/usr/bin/env python3
"""Per-call latency of the SQL path at 10 M rows: through the drop-in module (Python results) and through the C-ABI alone."""
import sys, time, statistics
sys.path.insert(0, "/root/repo")
import approximatequeryengine_b200 as aqe
b = aqe.backend()
db = b.CustomBPlusDB(0); db.generate_synthetic(10_000_000, 7)
for name, f in (("query SUM", lambda: db.query("SELECT SUM(amount) FROM sales")), ("query SUM WHERE", lambda: db.query("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500")),
                ("groupby region", lambda: db.query_groupby("SELECT SUM(amount) FROM sales GROUP BY region")),
                ("groupby product_id", lambda: db.query_groupby("SELECT SUM(amount) FROM sales GROUP BY product_id")),
                ("groupby region ci p=10", lambda: db.query_groupby_with_ci("SELECT SUM(amount) FROM sales GROUP BY region", 10)),
                ("sum_amount()", lambda: db.sum_amount())):
    f(); ts = []
    for _ in range(200):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    print(f"{name:28s} p50 {statistics.median(ts) * 1e6:8.1f} us")
# the C-ABI alone (ctypes): parse + facts + layout + one kernel + finish, no Python dict of results
import ctypes as C
e = aqe.Engine(0).generate(10_000_000, seed=7)
rows = (aqe.SqlRow * aqe.SQL_MAX_GROUPS)(); n = C.c_uint32()
for sql in ("SELECT SUM(amount) FROM sales", "SELECT SUM(amount) FROM sales GROUP BY region", "SELECT SUM(amount) FROM sales GROUP BY product_id"):
    ts = []
    for _ in range(220):
        t = time.perf_counter(); aqe.check(e.L.aqe_sql_run(e.h, sql.encode(), 0, 0, rows, aqe.SQL_MAX_GROUPS, C.byref(n))); ts.append(time.perf_counter() - t)
    print(f"aqe_sql_run {sql[7:]:45s} p50 {statistics.median(ts[20:]) * 1e6:8.1f} us")
