#!/usr/bin/env python3
"""Secondary measurements on one B200 (everything next to the headline scan): other column types / predicates,
sampled aggregates over plans (K3/K5), persistent CLT kernel at 1 B rows (configs[3]), the legacy list[Record]
return path (K6) through the drop-in module, lock-step CLT sampler, ingest (K7).

    python tools/secondary_bench.py [--records 1000000000] > profiles/rN_secondary.json
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import approximatequeryengine_b200 as aqe

import faulthandler
faulthandler.enable()


def note(msg):
    print(msg, file=sys.stderr, flush=True)


ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=1_000_000_000)
ap.add_argument("--ingest-records", type=int, default=50_000_000)
ap.add_argument("--only-ingest", action="store_true")
args = ap.parse_args()
N = args.records
out = {"records": N}
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
slot = torch.zeros(8, dtype=torch.int64, device="cuda")


def wall(f, reps=7):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3, r


def dev_ms(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        f()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if args.only_ingest:
    n_in = args.ingest_records
    g = aqe.Engine(0).generate(n_in, seed=7)
    res = {}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        path = os.path.join(td, "t.aqe")
        g.save_file(path)
        for th in (1, 2, 4, 8, 16):
            os.environ["AQE_INGEST_THREADS"] = str(th)
            ts = []
            for _ in range(3):
                x = aqe.Engine(0); t1 = time.perf_counter(); x.load_file(path); ts.append(time.perf_counter() - t1)
                assert x.count == n_in and x.sum_int("id") == n_in * (n_in + 1) // 2
                x.close()
            res[f"threads_{th}"] = {"load_s_best": min(ts), "GBps": 32 * n_in / min(ts) / 1e9, "records_per_s": n_in / min(ts)}
            note(f"ingest threads={th}: {res[f'threads_{th}']}")
    print(json.dumps({"ingest_sweep": res, "records": n_in, "medium": "/dev/shm (page cache)"}, indent=1))
    sys.exit(0)

# ---- 1. exact scans over every column type / predicate combination (bytes = algorithmic bytes per record) ----
e = aqe.Engine(0).generate(N, seed=7)
t0 = 1700000000
scans = []
for agg, pred, lo, hi, bpr in [("amount", None, 0, 0, 8), ("amount", "amount", 100.0, 500.0, 8), ("amount", "timestamp", t0, t0 + N // 2, 16),
                               ("amount", "region", 2, 5, 12), ("id", None, 0, 0, 8), ("timestamp", "amount", 100.0, 500.0, 16),
                               ("region", None, 0, 0, 4), ("product_id", "region", 0, 3, 8), ("id", "id", 1, N // 2, 8)]:
    ms = dev_ms(lambda: e.scan_async(slot.data_ptr(), agg, pred, lo, hi, stream=stream.cuda_stream))
    note(f"scan {agg} {pred} {ms:.3f} ms")
    scans.append({"agg": agg, "pred": pred, "bytes_per_record": bpr, "ms": ms, "GBps": bpr * N / ms / 1e6, "records_per_s": N / ms * 1e3})
out["exact_scans"] = scans

# ---- 2. sampled aggregates over plans (K3/K5): device-side moments, no rows returned ----
plans = []
for m, p, kw in [("memory_stride", 1.0, {}), ("memory_stride", 10.0, {}), ("block", 1.0, {}), ("parallel_block", 1.0, {}), ("parallel_pointer", 1.0, {}),
                 ("index_based", 1.0, {}), ("optimized_clt", 20.0, {}), ("multithreaded_memory_stride", 4.0, {"seed": 1})]:
    prm = aqe.make_params(m, p, **kw)
    note(f"plan {m} {p}")
    ms_plan, pl = wall(lambda: e.plan(m, prm))
    ms, st = wall(lambda: e.stats(pl))
    plans.append({"method": m, "percent": p, "samples": st.n, "segments": pl.num_segments, "plan_ms": ms_plan, "stats_ms": ms, "samples_per_s": st.n / ms * 1e3,
                  "estimate_sum": st.sum * (N / st.n)})
out["plan_stats"] = plans

# ---- 3. persistent CLT kernel at full size (configs[3]: block sampling, APPROX SUM at 0.5 %) ----
approx = []
for design, agg, eps, kw in [("srs", "sum", 1.0, {}), ("srs", "sum", 0.5, {}), ("srs", "avg", 0.1, {}), ("block", "sum", 0.5, {}), ("block", "sum", 0.5, {"block_size": 128}),
                             ("block", "avg", 1.0, {}), ("srs", "sum", 0.5, {"where": (100.0, 500.0)}), ("srs", "count", 1.0, {"where": (100.0, 500.0)})]:
    note(f"approx {design} {agg} {eps} {kw}")
    lat, kus, ns, est = [], [], [], []
    for seed in range(30):
        t1 = time.perf_counter(); r = e.approx(agg, error_percent=eps, design=design, seed=seed, **kw); lat.append((time.perf_counter() - t1) * 1e6)
        kus.append(r.elapsed_us); ns.append(r.n_samples); est.append(r.estimate)
    approx.append({"design": design, "agg": agg, "error_percent": eps, **{k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()},
                   "call_us_p50": statistics.median(lat), "kernel_us_p50": statistics.median(kus), "rows_read_p50": statistics.median(ns), "estimate_p50": statistics.median(est)})
out["approx_full_size"] = approx
e.close()

# ---- 4. drop-in module, 10 M rows: the CLI's sampler calls (list[Record]) vs the array / device-stats forms ----
b = aqe.backend()
db = b.CustomBPlusDB(0)
db.generate_synthetic(10_000_000, 7)
legacy = []
for name, call in [("sum_amount()", lambda: db.sum_amount()), ("sum_amount_where(100,500)", lambda: db.sum_amount_where(100, 500)),
                   ("memory_stride_sample(1.0) -> list[Record]", lambda: len(db.memory_stride_sample(1.0, 0))),
                   ("sample_array('memory_stride', 1.0) -> numpy", lambda: len(db.sample_array("memory_stride", 1.0))),
                   ("sample_array('memory_stride', 1.0, stats=True) -> moments", lambda: db.sample_array("memory_stride", 1.0, stats=True)["n"]),
                   ("block_sample(1.0) -> list[Record]", lambda: len(db.block_sample(1.0))),
                   ("clt_validated_dual_pointer_sample(20,0.95,10,4,1.0) -> list[Record]", lambda: len(db.clt_validated_dual_pointer_sample(20, 0.95, 10, 4, 1.0))),
                   ("sample_array('clt_validated_dual_pointer', 20, max_error_percent=1.0)", lambda: len(db.sample_array("clt_validated_dual_pointer", 20.0, max_error_percent=1.0))),
                   ("random_pointer_sample(1.0) -> list[Record]", lambda: len(db.random_pointer_sample(1.0))),
                   ("approx_avg(1.0)", lambda: db.approx_avg(1.0).samples_used)]:
    note(f"dropin {name}")
    ms, r = wall(call, reps=5)
    legacy.append({"call": name, "ms": ms, "result": r})
out["dropin_10M"] = legacy
del db

# ---- 5. ingest: record file -> HBM columns (K7) ----
n_in = args.ingest_records
g = aqe.Engine(0).generate(n_in, seed=7)
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
    path = os.path.join(td, "t.aqe")
    note(f"ingest save {path}")
    t1 = time.perf_counter(); g.save_file(path); t_save = time.perf_counter() - t1
    note("ingest load")
    ts = []
    for _ in range(3):
        x = aqe.Engine(0); t1 = time.perf_counter(); x.load_file(path); ts.append(time.perf_counter() - t1); assert x.count == n_in; x.close()
    rows = g.read_rows(0, min(n_in, 20_000_000))
    t1 = time.perf_counter(); y = aqe.Engine(0).from_rows(rows); t_rows = time.perf_counter() - t1
    out["ingest"] = {"records": n_in, "file_bytes": 24 + 32 * n_in, "save_s": t_save, "load_s_best": min(ts), "load_GBps": 32 * n_in / min(ts) / 1e9,
                     "load_records_per_s": n_in / min(ts), "from_host_rows_GBps": 32 * len(rows) / t_rows / 1e9, "medium": os.path.dirname(path)}
# ---- 6. SURVEY 8d (i): the reference's own CPU implementation of the CLI's calls (oracle/_ref = the unmodified reference
#         compiled in place), 1 M rows in its B+ tree, median-free mean of `reps` calls inside C++, next to the same calls
#         through the drop-in module on the GPU at the same size ----
try:
    from oracle import Oracle, Ref
    if Ref.available():
        O = Oracle()
        n_ref = 1_000_000
        rows = O.synth(n_ref, seed=7)
        note("reference insert_batch 1M")
        t1 = time.perf_counter(); R = Ref(rows); t_load = time.perf_counter() - t1
        db = b.CustomBPlusDB(0)
        db.from_array(rows)
        ref = {"records": n_ref, "cores": os.cpu_count(), "reference_insert_batch_s": t_load, "calls": []}
        for name, what, a, bb, reps, ours in [("sum_amount()", 0, 0.0, 0.0, 5, lambda: db.sum_amount()),
                                              ("sum_amount_where(100,500)", 1, 100.0, 500.0, 5, lambda: db.sum_amount_where(100, 500)),
                                              ("memory_stride_sample(1.0)", 2, 1.0, 0.0, 20, lambda: db.sample_array("memory_stride", 1.0, stats=True)["n"]),
                                              ("clt_validated_dual_pointer_sample(20,0.95,10,4,1.0)", 3, 20.0, 1.0, 5, lambda: db.sample_array("clt_validated_dual_pointer", 20.0, max_error_percent=1.0, stats=True)["n"]),
                                              ("block_sample(1.0,1000)", 4, 1.0, 0.0, 5, lambda: db.sample_array("block", 1.0, stats=True)["n"])]:
            note(f"reference {name}")
            t, v = R.time(what, reps, a, bb)
            ms_ours, r_ours = wall(ours, reps=5)
            ref["calls"].append({"call": name, "reference_cpu_ms": t / reps * 1e3, "reference_result": v, "gpu_ms": ms_ours, "gpu_result": r_ours,
                                 "gpu_form": "device moments of the same sample (no list[Record])" if what >= 2 else "same call"})
        out["reference_cpu_1M"] = ref
except Exception as ex:  # noqa: BLE001
    out["reference_cpu_1M"] = {"unavailable": str(ex)}
# ---- 7. the reference's SQL-string path (executor.cpp over SQLite, compiled unmodified: oracle/_ref/libaqe_refsql.so) on a
#         1 M-row SQLite copy of the table, next to run_query* of the drop-in on the record file of the same rows ----
try:
    from oracle import Oracle, RefSql
    if RefSql.available():
        O = Oracle()
        n_sql = 1_000_000
        rows = O.synth(n_sql, seed=7)
        with tempfile.TemporaryDirectory() as td:
            sqlite_path, rec_path = os.path.join(td, "sales.db"), os.path.join(td, "sales.aqe")
            note("sqlite copy of 1M rows")
            t1 = time.perf_counter(); RefSql.make_sqlite(sqlite_path, rows); t_make = time.perf_counter() - t1
            O.save_file(rec_path, rows)
            R = RefSql(sqlite_path)
            res = {"records": n_sql, "cores": os.cpu_count(), "sqlite_build_s": t_make, "queries": []}
            for sql, p, mode, ours in [("SELECT SUM(amount) FROM sales", 0, "run_query", lambda q, pp: b.run_query(q, rec_path, pp)),
                                       ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10, "run_query_with_ci", lambda q, pp: b.run_query_with_ci(q, rec_path, pp).value),
                                       ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "run_query_groupby", lambda q, pp: len(b.run_query_groupby(q, rec_path, pp, 4))),
                                       ("SELECT AVG(amount) FROM sales GROUP BY region", 10, "run_query_groupby_with_ci", lambda q, pp: len(b.run_query_groupby_with_ci(q, rec_path, pp, 4))),
                                       ("SELECT COUNT(amount) FROM sales WHERE product_id < 100 GROUP BY product_id", 0, "run_query_groupby", lambda q, pp: len(b.run_query_groupby(q, rec_path, pp, 4)))]:
                note(f"reference sql {sql}")
                ts = []
                for _ in range(3):
                    t1 = time.perf_counter(); r_ref = R.run(sql, p, mode); ts.append(time.perf_counter() - t1)
                ours(sql, p)   # first call loads the record file into HBM (cached by path afterwards)
                ms_ours, r_ours = wall(lambda: ours(sql, p), reps=7)
                res["queries"].append({"sql": sql, "sample_percent": p, "entry_point": mode, "reference_cpu_ms": statistics.median(ts) * 1e3,
                                       "reference_groups_or_value": len(r_ref) if "groupby" in mode else r_ref[0][1], "gpu_ms": ms_ours, "gpu_groups_or_value": r_ours})
            b.close_cached_tables()
        out["reference_sql_1M"] = res
except Exception as ex:  # noqa: BLE001
    out["reference_sql_1M"] = {"unavailable": str(ex)}
print(json.dumps(out, indent=1))
