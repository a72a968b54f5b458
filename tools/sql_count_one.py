#!/usr/bin/env python3
"""COUNT queries over one int32 column once on a device-generated table (ncu target for the K = 16 instantiations of k_sql_ring:
`ncu -k regex:k_sql_ring --set full ... python tools/sql_count_one.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approximatequeryengine_b200 as aqe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
e = aqe.Engine(0).generate(n, seed=7)
for sql in ("SELECT COUNT(*) FROM sales WHERE region = 1",
            "SELECT COUNT(*) FROM sales WHERE region IN (1, 3, 5, 7)",
            "SELECT COUNT(amount) FROM sales GROUP BY region",
            "SELECT COUNT(amount) FROM sales GROUP BY product_id",
            "SELECT COUNT(*) FROM sales WHERE region != 2 GROUP BY region"):
    r = e.sql(sql, 0, "value")
    print(sql, len(r), r[0].count)
