#!/usr/bin/env python3
"""aqe-b200 command line: the reference's `enhanced_aqe_cli.py` with its routing repaired.

The reference CLI offers three syntaxes -- `SELECT APPROX(SUM(amount)) ...`, `--s PERCENT`, `--e ERROR` -- but tests
`args.s` / `args.e`, which argparse never sets (dests are `sample` / `error`), so only the embedded form ever reaches
a sampler (enhanced_aqe_cli.py:97-114); its SUM interval is `MoE * N/n` (too narrow by a factor n,
:285) and it dies printing the execution time (:210-212).  This one routes all three, prints the same
result fields (estimate, confidence interval, error margin, sample count, status, time) and computes them on the
device -- no list[Record] ever crosses into Python:

  exact                      sum_amount / get_total_records (enhanced_aqe_cli.py:320-370); queries with WHERE on other
                             columns or GROUP BY go through the SQL path (run_query*, executor.cpp)
  --s P / APPROX(...)        the reference's "random" routing by table size (:172-186: memory_stride / direct_access /
                             optimized_sequential sampler), estimator E1 (:190-200) from device moments, correct interval
  --e E                      CLT early termination: the persistent Philox kernel (`approx_sum/avg/count`); with
                             --method parallel the reference's own clt_validated_dual_pointer_sample rows (:243-255)
  --method block             contiguous 1000-row tiles (block_sample / the BLOCK design of the persistent kernel)

    python -m approximatequeryengine_b200.cli "SELECT APPROX(SUM(amount)) FROM sales" --db sales.aqe
    python -m approximatequeryengine_b200.cli "SELECT AVG(amount) FROM sales" --db sales.aqe --e 1 --compare
    python -m approximatequeryengine_b200.cli "SELECT SUM(amount) FROM sales GROUP BY region" --db sales.aqe --s 10 --json
"""
from __future__ import annotations

import argparse
import json
import math
import re
import sys
import time

METHODS = ("random", "clt", "block", "parallel", "stratified", "adaptive", "revolutionary")


def parse_embedded_approx(query: str):
    """enhanced_aqe_cli.py:83-95: strip one APPROX( ... ) wrapper."""
    m = re.search(r"APPROX\s*\(\s*((?:[^()]|\([^()]*\))+)\s*\)", query, re.IGNORECASE)
    if not m:
        return query, False
    return query[: m.start()] + m.group(1) + query[m.end():], True


def aggregate_of(query: str):
    m = re.search(r"\b(SUM|AVG|COUNT)\s*\(\s*([\w*]+)\s*\)", query, re.IGNORECASE)
    if not m:
        raise ValueError("unsupported query: expected SUM / AVG / COUNT (enhanced_aqe_cli.py:188-200)")
    return m.group(1).upper(), m.group(2).lower()


def plain_amount_query(query: str) -> bool:
    """True when the CustomBPlusDB calls of the reference CLI apply: one aggregate of `amount` (or COUNT), no WHERE / GROUP BY."""
    agg, col = aggregate_of(query)
    return not re.search(r"\b(WHERE|GROUP\s+BY)\b", query, re.IGNORECASE) and (col in ("amount", "*") or agg == "COUNT")


def estimate_from_moments(agg: str, st: dict, N: int, z: float):
    """E1/E2 (enhanced_aqe_cli.py:190-200, 262-291) with the SUM margin scaled as a total (MoE * N)."""
    n = st["n"]
    if n == 0:
        raise RuntimeError("No samples collected")
    var = st["m2"] / (n - 1) if n > 1 else 0.0
    moe = z * math.sqrt(var) / math.sqrt(n)
    if agg == "SUM":
        return st["sum"] * (N / n), moe * N
    if agg == "AVG":
        return st["sum"] / n, moe
    return float(N), 0.0  # COUNT is exact (:196-197)


def run(args) -> dict:
    from . import backend

    b = backend()
    query, embedded = parse_embedded_approx(args.query)
    agg, col = aggregate_of(query)
    db = b.CustomBPlusDB()
    if not db.open_database(args.db):
        raise RuntimeError(f"Could not open database: {args.db}")
    N = db.get_total_records()
    z = b.z_score(args.confidence, 1)
    t0 = time.perf_counter()
    out = {"query": args.query, "records": N, "aggregate": agg}

    def exact_value():
        if plain_amount_query(query):
            return {"SUM": db.sum_amount, "AVG": lambda: db.sum_amount() / N if N else 0.0, "COUNT": lambda: float(N)}[agg]()
        return db.query_groupby(query, 0) if re.search(r"GROUP\s+BY", query, re.IGNORECASE) else db.query(query, 0)

    sample = args.sample if args.sample is not None else (10.0 if embedded and args.error is None else None)
    if args.error is None and sample is None:
        out.update(mode="exact", value=exact_value(), confidence_level=1.0, error_margin=0.0, samples_used=N, status="STABLE")
    elif not plain_amount_query(query):
        # WHERE / GROUP BY / other columns: the SQL path samples with rowid % (100/p) = 0 (executor.cpp:20-42)
        p = int(round(sample if sample is not None else {True: 20, False: 10}[args.error <= 1]))
        grouped = bool(re.search(r"GROUP\s+BY", query, re.IGNORECASE))
        if grouped:
            r = db.query_groupby_with_ci(query, p, correct_ci=True)
            out.update(mode=f"sql sample {p}%", value={k: v.value for k, v in r.items()}, ci={k: [v.ci_lower, v.ci_upper] for k, v in r.items()})
        else:
            r = db.query_with_ci(query, p, correct_ci=True)
            out.update(mode=f"sql sample {p}%", value=r.value, ci=[r.ci_lower, r.ci_upper],
                       error_margin=(r.ci_upper - r.ci_lower) / 2 / abs(r.value) if r.value else 0.0)
        out.update(confidence_level=0.95, status="STABLE")
    elif args.error is not None and args.method != "parallel":
        design = "block" if args.method == "block" else "srs"
        r = {"SUM": db.approx_sum, "AVG": db.approx_avg, "COUNT": db.approx_count}[agg](error_percent=args.error, confidence_level=args.confidence, design=design)
        out.update(mode=f"clt {design} +-{args.error}%", value=r.value, ci=[r.ci_lower, r.ci_upper], confidence_level=r.confidence_level,
                   error_margin=r.error_margin, samples_used=r.samples_used, status=r.status.name, kernel_us=r.kernel_us)
    else:
        if args.error is not None:   # the reference's CLT sampler and its percent map (enhanced_aqe_cli.py:243-255)
            pct = 20.0 if args.error <= 1 else 15.0 if args.error <= 2 else 10.0 if args.error <= 5 else 5.0
            method, kw = "clt_validated_dual_pointer", dict(confidence_level=args.confidence, check_interval=10, num_threads=args.threads, max_error_percent=args.error)
        else:
            pct = sample
            if args.method == "block":
                method, kw = "block", {}
            elif args.method == "parallel":
                method, kw = "parallel_pointer", dict(num_threads=args.threads)
            elif args.method == "stratified":
                method, kw = "stratified_block", {}
            else:                    # :172-186
                method, kw = ("memory_stride" if N > 50000 else "direct_access" if N > 10000 else "optimized_sequential"), {}
        st = db.sample_array(method, pct, stats=True, **kw)
        value, margin = estimate_from_moments(agg, st, N, z)
        out.update(mode=f"{method} {pct:g}%", value=value, ci=[value - margin, value + margin], confidence_level=args.confidence,
                   error_margin=margin / abs(value) if value else 0.0, samples_used=st["n"], status="STABLE")
    out["time_ms"] = (time.perf_counter() - t0) * 1e3
    if args.compare and out.get("mode") != "exact":
        t1 = time.perf_counter()
        ex = exact_value()
        out["exact"] = ex
        out["exact_time_ms"] = (time.perf_counter() - t1) * 1e3
        if isinstance(ex, float) and ex:
            out["actual_error_percent"] = abs(out["value"] - ex) / abs(ex) * 100.0
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="aqe-b200", description=__doc__.split("\n\n")[0], allow_abbrev=False)
    ap.add_argument("query", help="SELECT SUM|AVG|COUNT(col) FROM t [WHERE ...] [GROUP BY g], optionally wrapped in APPROX(...)")
    ap.add_argument("--db", default="custom_demo.db", help="record file (CustomBPlusDB.save_to_file format)")
    ap.add_argument("-s", "--s", "--sample", dest="sample", type=float, metavar="PERCENT", help="sample percentage")
    ap.add_argument("-e", "--e", "--error", dest="error", type=float, metavar="THRESHOLD", help="error threshold in percent (CLT early termination)")
    ap.add_argument("--method", choices=METHODS)
    ap.add_argument("--compare", action="store_true", help="also run the exact query")
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--confidence", type=float, default=0.95)
    ap.add_argument("--ci", action="store_true", help="accepted for compatibility (intervals are always shown)")
    ap.add_argument("--json", action="store_true", help="print one JSON object")
    args = ap.parse_args(argv)
    try:
        out = run(args)
    except (RuntimeError, ValueError) as e:
        print(f"error: {e}", file=sys.stderr)
        return 1
    if args.json:
        print(json.dumps(out))
        return 0
    print(f"{out['mode']}: {out['aggregate']} over {out['records']} records")
    print(f"  estimate          {out['value']}")
    if "ci" in out:
        print(f"  confidence        {out.get('confidence_level')}  interval {out['ci']}")
    if "error_margin" in out:
        print(f"  error margin      {out['error_margin'] * 100:.4g} %")
    if "samples_used" in out:
        print(f"  samples used      {out['samples_used']}")
    print(f"  status            {out.get('status')}")
    print(f"  time              {out['time_ms']:.3f} ms")
    if "exact" in out:
        print(f"  exact             {out['exact']}  ({out['exact_time_ms']:.3f} ms)" + (f"  actual error {out['actual_error_percent']:.4g} %" if "actual_error_percent" in out else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
