#!/usr/bin/env python3
"""BASELINE.json configs[4]: CI-coverage sweep -- seeds x error thresholds on a 100 M-record table, the fused
persistent-kernel estimator (k_approx) vs the reference CLI's estimator formulas on the SAME samples.

    python tools/coverage_sweep.py [--records 100000000] [--seeds 1000] > profiles/rN_coverage.json

For every (distribution, aggregate, threshold): fraction of seeds whose interval covers the exact answer
(ours: z*s/sqrt(n) scaled by N for SUM; reference: enhanced_aqe_cli.py:281-291, SUM margin scaled by N/n),
median samples drawn, median kernel time, achieved relative error.  Exact answers come from the exact scan."""
import argparse
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import approximatequeryengine_b200 as aqe

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=100_000_000)
ap.add_argument("--seeds", type=int, default=1000)
ap.add_argument("--design", default="srs")
args = ap.parse_args()

out = {"records": args.records, "seeds": args.seeds, "design": args.design, "confidence_level": 0.95, "rows": []}
for dist, dname in ((0, "uniform(1,1000)"), (1, "lognormal(mu=4,sigma=1.5)")):
    e = aqe.Engine(0).generate(args.records, seed=7, dist=dist, columns=("amount",))
    tot = e.scan("amount")
    truth = {"sum": tot.sum, "avg": tot.sum / args.records}
    for agg in ("sum", "avg"):
        for eps in (0.1, 0.5, 1.0, 2.0, 5.0):
            hit = hit_ref = 0
            ns, us, relerr, status = [], [], [], []
            for seed in range(args.seeds):
                r = e.approx(agg, error_percent=eps, confidence_level=0.95, design=args.design, seed=seed)
                hit += r.ci_lower <= truth[agg] <= r.ci_upper
                # the reference CLI's formulas on the same sample moments (SRS: units are rows)
                st = aqe.Stats(n=r.n_units, mean=r.mean, m2=r.m2, sum=r.mean * r.n_units)
                _, lo, hi = aqe.estimate(st, args.records, agg, 1.96, legacy_ci=True)
                hit_ref += lo <= truth[agg] <= hi
                ns.append(r.n_samples); us.append(r.elapsed_us); status.append(r.status)
                relerr.append(abs(r.estimate - truth[agg]) / abs(truth[agg]) * 100)
            out["rows"].append({"distribution": dname, "agg": agg, "error_percent": eps, "coverage": hit / args.seeds,
                                "coverage_reference_cli_formula": hit_ref / args.seeds, "samples_median": statistics.median(ns),
                                "kernel_us_median": statistics.median(us), "kernel_us_max": max(us), "abs_error_percent_median": statistics.median(relerr),
                                "abs_error_percent_p99": sorted(relerr)[int(0.99 * len(relerr)) - 1], "stable_fraction": status.count(0) / len(status)})
            print(json.dumps(out["rows"][-1]), file=sys.stderr, flush=True)
    e.close()
out["binomial_sigma"] = math.sqrt(0.95 * 0.05 / args.seeds)
print(json.dumps(out, indent=1))
