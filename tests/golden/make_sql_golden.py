#!/usr/bin/env python3
"""Mint golden vectors for the SQL-string path (run_query*, bindings.cpp:126-136) from the UNMODIFIED reference:
src/aqe_backend/executor.cpp + parser.cpp + core/db.cpp compiled where they lie by `make -C oracle refsql`
(against the system SQLite runtime through oracle/sqlite_shim/sqlite3.h) and run on a SQLite copy
`sales(id INTEGER PRIMARY KEY, amount REAL, region INTEGER, product_id INTEGER, timestamp INTEGER)` of the same
seeded synthetic rows the other goldens use.  Run in the container that has /root/reference:

    python tests/golden/make_sql_golden.py

Writes tests/golden/sql_n<N>_s<seed>.json.  Floats are C99 hex strings (exact).  A case whose reference call
throws records {"error": "stod" | "runtime_error", "msg": ...}.  Grouped queries in which some group has no sampled
row (or whose SUM(col*col) overflows int64) are NOT run against the reference: its worker thread throws
std::invalid_argument("stod") / std::runtime_error("SQL error: integer overflow") outside any handler and the process
aborts (executor.cpp:100, :279); such cases are recorded as {"error": "terminate"} on the strength of the
restatement's prediction.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle, RefSql, SqlError, build  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

QUERIES = [
    "SELECT SUM(amount) FROM sales",
    "SELECT COUNT(amount) FROM sales",
    "SELECT AVG(amount) FROM sales",
    "SELECT COUNT(*) FROM sales",
    "select sum(amount) from sales;",
    "SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500",
    "SELECT COUNT(amount) FROM sales WHERE amount BETWEEN 100 AND 500",
    "SELECT AVG(amount) FROM sales WHERE amount >= 250.5 AND amount < 750",
    "SELECT SUM(amount) FROM sales WHERE region = 3 AND timestamp > 1700000500;",
    "SELECT SUM(amount) FROM sales WHERE region = '3'",
    "SELECT COUNT(*) FROM sales WHERE region != 3 AND region >= 2 AND 6 >= region AND amount < 1e2",
    "SELECT SUM(timestamp) FROM sales",
    "SELECT SUM(id) FROM sales WHERE rowid > 10",
    "SELECT AVG(timestamp) FROM sales WHERE region <> 2",
    "SELECT SUM(region) FROM sales WHERE amount > 990.5",
    "SELECT AVG(product_id) FROM sales WHERE product_id >= 10 AND product_id < 20.5 AND region != 0",
    "SELECT SUM(amount) FROM sales WHERE amount > 5000",
    "SELECT COUNT(*) FROM sales WHERE amount > 5000",
    "SELECT AVG(amount) FROM sales WHERE (amount > 10) AND (region = 1)",
    "SELECT SUM(amount) FROM sales WHERE 1 = 1 AND id <= 600",
    "SELECT SUM(amount) FROM sales GROUP BY region",
    "SELECT COUNT(amount) FROM sales GROUP BY region",
    "SELECT AVG(amount) FROM sales WHERE amount < 300 GROUP BY region;",
    "SELECT SUM(product_id) FROM sales GROUP BY region",
    "SELECT AVG(timestamp) FROM sales GROUP BY region",
    "SELECT SUM(amount) FROM sales WHERE product_id < 48 GROUP BY product_id",
    "SELECT COUNT(amount) FROM sales WHERE id <= 90 GROUP BY product_id",
    "SELECT SUM(amount) FROM sales WHERE product_id BETWEEN 100 AND 120 GROUP BY product_id",
    "SELECT AVG(amount) FROM sales WHERE region = 5 GROUP BY region",
    # OR / parentheses
    "SELECT SUM(amount) FROM sales WHERE region = 1 OR region = 3",   # top-level OR: comparable for exact, ungrouped calls only
    "SELECT SUM(amount) FROM sales WHERE (region = 1 OR region = 3)",
    "SELECT COUNT(amount) FROM sales WHERE (region = 1 OR region = 3) AND amount > 500",
    "SELECT AVG(amount) FROM sales WHERE (amount < 100 OR amount > 900 OR product_id = 7)",
    "SELECT SUM(amount) FROM sales WHERE (region < 2 OR region > 5) AND (product_id < 300 OR timestamp > 1700000900) GROUP BY region",
    "SELECT SUM(amount) FROM sales WHERE region IN (1, 3, 5)",
    "SELECT COUNT(amount) FROM sales WHERE product_id IN (7, 77, 777, 7) AND amount > 100 GROUP BY region",
    "SELECT AVG(amount) FROM sales WHERE NOT (region = 1 OR amount > 500) AND product_id NOT IN (3, 4)",
    "SELECT SUM(amount) FROM sales WHERE amount NOT BETWEEN 100 AND 900 GROUP BY region",
    # what parser.cpp rejects
    "SELECT MAX(amount) FROM sales",
    "SELECT amount FROM sales",
    "SUM(amount) sales",
]
PERCENTS = [0, 1, 7, 10, 30, 50, 60, 100]
MODES = ["run_query", "run_query_with_ci", "run_query_groupby", "run_query_groupby_with_ci"]
PARSE_ONLY = [
    "SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region;",
    "  select   avg( amount )   from   sales   where region=1   group by   product_id ;",
    "SELECT COUNT(*) FROM t",
    "SELECT SUM(x) FROM wherehouse",
    "SELECT SUM(amount) FROM sales GROUP BY region WHERE amount > 1",
]


def enc(rows):
    """[key, value] when the interval is degenerate, else [key, value, ci_lower, ci_upper]."""
    out = []
    for k, v, lo, hi in rows:
        same = (lo == v and hi == v) or (v != v and lo != lo and hi != hi)
        out.append([int(k), float(v).hex()] if same else [int(k), float(v).hex(), float(lo).hex(), float(hi).hex()])
    return out


def mint(n: int, seed: int):
    O = Oracle()
    rows = O.synth(n, seed=seed)
    with tempfile.TemporaryDirectory() as td:
        db = os.path.join(td, "sales.db")
        RefSql.make_sqlite(db, rows)
        R = RefSql(db)
        cases = []
        for sql in QUERIES:
            for p in PERCENTS:
                for mode in MODES:
                    grouped = "groupby" in mode
                    if ("GROUP BY" in sql.upper()) != grouped and "FROM" in sql.upper() and "(" in sql and "MAX" not in sql:
                        continue
                    case = {"sql": sql, "p": p, "mode": mode}
                    predicted = None
                    try:
                        O.sql(rows, sql, p, mode)
                    except SqlError as e:
                        predicted = "stod" if e.kind == "stod" or "integer overflow" in e.msg else e.kind
                    if grouped and predicted == "stod":  # thrown inside a worker thread of the reference
                        case["error"] = "terminate"
                    else:
                        try:
                            case["rows"] = enc(R.run(sql, p, mode))
                        except SqlError as e:
                            case["error"] = e.kind
                            case["msg"] = e.msg
                    cases.append(case)
        parses = [{"sql": s, "parsed": R.parse(s)} for s in PARSE_ONLY + [q for q in QUERIES if "MAX" not in q and "FROM" in q.upper() and "(" in q]]
    out = {"n": n, "seed": seed, "rows_sha256": hashlib.sha256(rows.tobytes()).hexdigest(),
           "table": "sales(id INTEGER PRIMARY KEY, amount REAL, region INTEGER, product_id INTEGER, timestamp INTEGER)",
           "sqlite_version": __import__("sqlite3").sqlite_version, "cases": cases, "parses": parses}
    path = os.path.join(OUT, f"sql_n{n}_s{seed}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print(path, len(cases), "cases", sum("error" in c for c in cases), "errors")


if __name__ == "__main__":
    build(ref=True)
    if not RefSql.available():
        sys.exit("oracle/_ref/libaqe_refsql.so missing: needs /root/reference and libsqlite3.so.0")
    for n, seed in ((1000, 7), (20000, 7), (100000, 42)):
        mint(n, seed)
