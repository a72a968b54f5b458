#!/usr/bin/env python3
"""Mint golden vectors from the UNMODIFIED reference (oracle/_ref/libaqe_ref.so, compiled in place from
/root/reference by `make -C oracle ref`).  Run here (the container that has /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/golden_n<N>_s<seed>.json.  The reference ships no golden vectors; these
are outputs of the reference itself on seeded synthetic tables, so the pin travels to the GPU box.

Inputs are regenerated from (N, seed) by the Philox generator (oracle.synth == aqe_synth_rows_host ==
the device generator); `rows_sha256` pins the generator.  Floats are stored as C99 hex strings (exact).

Estimates ("est") are the reference CLI's Python formulas applied to the reference's returned rows:
  enhanced_aqe_cli.py:190-195  SUM = sum(amounts) * (N/n);  AVG = sum(amounts)/n      (left-to-right sum())
  enhanced_aqe_cli.py:276-291  s2 = sum((x-mean)^2)/(n-1); MoE = 1.96*sqrt(s2)/sqrt(n);
                               SUM CI = est +- MoE*(N/n) (the reference's interval); AVG CI = est +- MoE
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import METHODS, Oracle, Ref, RefScheduler, make_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

DETERMINISTIC = [
    "slow_pointer", "fast_pointer", "dual_pointer", "parallel_pointer", "random_pointer", "memory_stride",
    "optimized_address_arithmetic", "index_based", "byte_offset", "optimized_clt", "block", "page",
    "parallel_block", "node_skip", "balanced_tree", "direct_access", "adaptive_block", "stratified_block",
]
PARAM_SETS = [
    {},
    {"num_threads": 3, "step_size": 3, "block_size": 128, "block_size_max": 5, "seed": 7},
]
PERCENTS = [0.5, 1.0, 5.0, 20.0]


def cli_estimates(amounts, N):
    """enhanced_aqe_cli.py:190-195 and 276-291, in Python floats exactly as the CLI evaluates them."""
    n = len(amounts)
    if n == 0:
        return None
    vals = [float(a) for a in amounts]
    sample_sum = sum(vals)
    est_sum = sample_sum * (N / n)
    est_avg = sample_sum / n
    out = {"sample_sum": float(sample_sum).hex(), "sum": float(est_sum).hex(), "avg": float(est_avg).hex()}
    if n > 1:
        mean = sum(vals) / n
        var = sum((x - mean) ** 2 for x in vals) / (n - 1)
        std = var ** 0.5
        moe = 1.96 * std / (n ** 0.5)
        scaled = moe * (N / n)
        out.update({
            "m2": float(sum((x - mean) ** 2 for x in vals)).hex(),
            "moe": float(moe).hex(),
            "sum_ci_legacy": [float(est_sum - scaled).hex(), float(est_sum + scaled).hex()],
            "avg_ci": [float(est_avg - moe).hex(), float(est_avg + moe).hex()],
        })
    return out


def sha_idx(idx) -> str:
    return hashlib.sha256(np.ascontiguousarray(idx, dtype="<i8").tobytes()).hexdigest()


def mint(N: int, seed: int) -> dict:
    O = Oracle()
    rows = O.synth(N, seed=seed)
    R = Ref(rows)
    g = {
        "n": N, "seed": seed, "dist": 0,
        "rows_sha256": hashlib.sha256(rows.tobytes()).hexdigest(),
        "first_row": {k: (float(rows[0][k]).hex() if k == "amount" else int(rows[0][k])) for k in rows.dtype.names},
        "total_records": int(R.total()),
        "tree_height": int(R.tree_height()),
        "node_count": int(R.node_count()),
        "sum_amount": float(R.sum_amount()).hex(),
        "avg_amount": float(R.avg_amount()).hex(),
        "sum_amount_where": {},
        "int_sums": {c: int(rows[c].astype(object).sum()) for c in ("id", "region", "product_id", "timestamp")},
        "samplers": [],
    }
    for lo, hi in [(100.0, 500.0), (1.0, 1000.0), (999.5, 2000.0), (500.0, 100.0)]:
        w = R.sum_amount_where(lo, hi)
        cnt = int(((rows["amount"] >= lo) & (rows["amount"] <= hi)).sum())
        g["sum_amount_where"][f"{lo},{hi}"] = {"sum": float(w).hex(), "count": cnt}
    for m in DETERMINISTIC:
        for p in PERCENTS:
            for kw in PARAM_SETS:
                T = int(N * p / 100.0)
                if m == "dual_pointer" and T < 3:
                    continue  # reference divides by zero (custom_bplus_db.cpp:796)
                if m == "memory_stride" and 255 <= N < 1000:
                    continue  # reference result depends on prior calls (stale subtree count)
                prm = make_params(m, p, **kw)
                got = R.sample(m, prm)
                idx = got["id"] - 1
                if m == "stratified_block":
                    pass  # ids identify rows; positions in amount order are recovered by the consumer
                g["samplers"].append({
                    "method": m, "percent": p, "kw": kw, "count": int(len(idx)), "idx_sha256": sha_idx(idx),
                    "head": [int(x) for x in idx[:4]], "tail": [int(x) for x in idx[-4:]],
                    "est": cli_estimates(got["amount"], N),
                })
    # scheduler: exact legs are deterministic (custom_scheduler.cpp:141-205)
    S = RefScheduler(rows)
    g["scheduler"] = {}
    for name, what in (("exact_sum", 0), ("exact_avg", 1), ("exact_count", 2)):
        r = S.run(what)
        g["scheduler"][name] = {"value": float(r.value).hex(), "status": r.status,
                                "confidence_level": r.confidence_level, "error_margin": r.error_margin,
                                "samples_used": r.samples_used}
    r = S.run(3, "SELECT SUM(amount) FROM sales", 10.0, 4)
    g["scheduler"]["sum_query_fields"] = {"status": r.status, "confidence_level": r.confidence_level,
                                          "error_margin": r.error_margin, "samples_used": r.samples_used}
    g["scheduler"]["size_mb"] = float(S.size_mb()).hex()
    return g


def main():
    cases = [(1000, 1), (1000, 7), (1000, 42), (12345, 7), (100000, 7), (100000, 42)]
    if "--big" in sys.argv:
        cases.append((1000000, 7))
    for N, seed in cases:
        g = mint(N, seed)
        path = os.path.join(OUT, f"golden_n{N}_s{seed}.json")
        with open(path, "w") as f:
            json.dump(g, f, indent=0, separators=(",", ":"))
        print(path, len(g["samplers"]), "sampler vectors")


if __name__ == "__main__":
    main()
