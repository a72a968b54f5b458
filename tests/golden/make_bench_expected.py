#!/usr/bin/env python3
"""Mints tests/golden/bench_expected.json: the exact answers bench.py checks its timed results against (count exact, sum to
1e-12) -- `SUM/COUNT(amount) WHERE amount BETWEEN 100 AND 500` over the first k x 10^9 rows of the synthetic table (seed 7), for
the table sizes bench.py runs (1 B rows in total for strong scaling; 2 / 4 / 8 B for --scaling weak on 2 / 4 / 8 GPUs).
The rows come from the oracle's generator (bit-identical to the device's), the sums from its long-double Neumaier
accumulation (oracle/aqe_oracle.c orc_sum_col_exact_mt); billions are summed chunk by chunk and folded in long double.

    python tests/golden/make_bench_expected.py        # ~1 minute on 8 cores, 8 GB of host memory
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

LO, HI, SEED, CHUNK = 100.0, 500.0, 7, 1_000_000_000
O = Oracle()
out = {"seed": SEED, "predicate": [LO, HI], "tables": {}}
col = np.empty(CHUNK, dtype=np.float64)
cnt_w = 0
s_w = np.longdouble(0)
s_all = np.longdouble(0)
for k in range(8):
    O.synth_amount(CHUNK, seed=SEED, first_row=k * CHUNK, out=col)
    sw, cw = O.sum_col_exact(col, (LO, HI))
    sa, _ = O.sum_col_exact(col)
    cnt_w += cw
    s_w += np.longdouble(sw)
    s_all += np.longdouble(sa)
    if k + 1 in (1, 2, 4, 8):
        out["tables"][str((k + 1) * CHUNK)] = {"count_where": int(cnt_w), "sum_where": float(s_w).hex(), "sum_all": float(s_all).hex()}
        print(k + 1, "B rows:", out["tables"][str((k + 1) * CHUNK)], flush=True)
with open(os.path.join(ROOT, "tests", "golden", "bench_expected.json"), "w") as f:
    json.dump(out, f, indent=1)
