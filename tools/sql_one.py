#!/usr/bin/env python3
"""Run a few SQL-path queries once on a device-generated table (the ncu target: `ncu -k regex:k_sql ... python tools/sql_one.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approximatequeryengine_b200 as aqe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
e = aqe.Engine(0).generate(n, seed=7)
for sql, p, mode in (("SELECT SUM(amount) FROM sales", 0, "value"),
                     ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "value"),
                     ("SELECT SUM(amount) FROM sales GROUP BY product_id", 0, "value"),
                     ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY product_id", 0, "value"),
                     ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", 0, "ci_reference"),
                     ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "ci_reference"),
                     ("SELECT SUM(amount) FROM sales WHERE product_id IN (1, 3, 5, 7)", 0, "value"),
                     ("SELECT AVG(amount) FROM sales GROUP BY region", 50, "ci_reference")):
    r = e.sql(sql, p, mode)
    print(sql, p, mode, len(r), r[0].value)
