#!/usr/bin/env python3
"""Multi-GPU check of the fused exchange (run under torchrun on an N-GPU box, N >= 2; tests/test_multi_gpu.py runs it from
pytest -m gpu wherever two or more GPUs are visible and keeps its JSON line as the pass record):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Every rank scans its shard; the scan kernel's last block stores the 64-byte partial into every rank's mailbox
over NVLink and folds all of them (aqe_scan_exchange).  Checked bit-for-bit against the NCCL all-gather + host
merge path (sharded.ShardedTable.scan) and against closed forms; then both paths are timed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import approximatequeryengine_b200 as aqe
from approximatequeryengine_b200 import sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
sys.stdout.flush()
real_stdout = os.fdopen(os.dup(1), "w")   # keep stdout for the one JSON line: NCCL printf()s its version banner to fd 1
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(os.environ.get("AQE_CHECK_ROWS", 1_000_000_007))
t = sharded.ShardedTable.synthetic(N, rank, world, seed=7, device=local, columns=("id", "amount", "timestamp"))
t0 = 1700000000
cases = [("amount", None, 0, 0), ("amount", "amount", 100.0, 500.0), ("amount", "timestamp", t0 + 5, t0 + N // 3),
         ("id", None, 0, 0), ("timestamp", "amount", 1.0, 250.0), ("id", "id", 10, N - 10)]
ref = [t.scan(*c) for c in cases]                       # NCCL all-gather + aqe_merge_partials
assert t.enable_fused_exchange()
for it in range(3):
    for c, r in zip(cases, ref):
        f = t.scan(*c)                                  # fused: one kernel
        assert bytes(f) == bytes(r), (rank, it, c, (f.count, f.sum, f.comp, f.isum), (r.count, r.sum, r.comp, r.isum))
assert ref[3].isum == N * (N + 1) // 2 and ref[0].count == N
# stress: many back-to-back exchanges (slot reuse / sequence numbers)
for it in range(2000):
    f = t.scan("amount", "amount", 100.0, 500.0) if it % 500 == 0 else None
    if f is not None:
        assert bytes(f) == bytes(ref[1])
# ---- sampled estimates with ONE global stop rule (per-look moments exchanged inside k_approx) vs the CPU restatement ----
import numpy as np
from oracle import ApproxSpec, Oracle
O = Oracle()
n_small = 4_000_003
ts = sharded.ShardedTable.synthetic(n_small, rank, world, seed=11, device=local, columns=("amount", "timestamp"))
assert ts.enable_fused_exchange()
rows_all = O.synth(n_small, seed=11)
approx_cases = [dict(agg="sum"), dict(agg="avg", eps=0.25), dict(agg="sum", where=(100.0, 500.0)), dict(agg="count", where=(100.0, 500.0)),
                dict(agg="avg", where=(100.0, 500.0), eps=0.5), dict(agg="sum", design="block", min_samples=64, block=500), dict(agg="sum", eps=0.02, max_samples=300000),
                dict(agg="count")]
for kw in approx_cases:
    for seed in (0, 9):
        a = ts.approx(kw["agg"], error_percent=kw.get("eps", 1.0), seed=seed, design=kw.get("design", "srs"), where=kw.get("where"),
                      min_samples=kw.get("min_samples", 0), max_samples=kw.get("max_samples", 0), block_size=kw.get("block", 0))
        spec = ApproxSpec(agg=aqe.AGG[kw["agg"]], design=aqe.DESIGN[kw.get("design", "srs")], agg_col=1, pred_col=1 if kw.get("where") else -1,
                          lo=kw["where"][0] if kw.get("where") else 0.0, hi=kw["where"][1] if kw.get("where") else 0.0, error_percent=kw.get("eps", 1.0),
                          confidence_level=0.95, seed=seed, min_samples=kw.get("min_samples", 0), max_samples=kw.get("max_samples", 0), block_size=kw.get("block", 0))
        o = O.approx_sharded(rows_all, world, spec)
        assert (a.n_units, a.n_samples, a.rounds, a.status, a.population) == (o.n_units, o.n_samples, o.rounds, o.status, n_small), (rank, kw, seed, (a.n_units, a.n_samples, a.rounds, a.status), (o.n_units, o.n_samples, o.rounds, o.status))
        assert abs(a.estimate - o.estimate) <= 1e-10 * abs(o.estimate) and abs((a.ci_upper - a.ci_lower) - (o.ci_upper - o.ci_lower)) <= 1e-8 * abs(o.ci_upper - o.ci_lower + 1e-300), (kw, seed)
# ---- legacy sampler families across the ranks (configs[3]): every rank walks the plan, gathers inside its window ----
from oracle import make_params as orc_params
tsamp = sharded.ShardedTable.synthetic(n_small, rank, world, seed=11, device=local)
sampler_cases = [("parallel_pointer", 2.0, dict(num_threads=6)), ("parallel_block", 3.0, dict(block_size=500, num_threads=3)), ("block", 1.0, {}),
                 ("memory_stride", 1.0, {}), ("optimized_clt", 5.0, dict(num_threads=5)), ("sample_records", 2.5, dict(seed=11)),
                 ("address_arithmetic", 1.5, dict(seed=4)), ("random_pointer", 0.05, dict(seed=9)), ("multithreaded_memory_stride", 2.0, dict(num_threads=5, seed=5))]
for method, pct, kw in sampler_cases:
    prm = aqe.make_params(method, pct, **kw)
    st = tsamp.stats(method, prm)
    sw = tsamp.stats(method, prm, where=(100.0, 500.0))
    got_rows = tsamp.gather(method, prm)
    idx = O.indices(rows_all, method, orc_params(method, pct, **kw))
    o = O.stats(rows_all, idx)
    assert st.n == len(idx) == sw.n and got_rows.tobytes() == rows_all[idx].tobytes(), (rank, method)
    assert abs(st.sum - o.sum) <= 1e-12 * abs(o.sum) and abs(st.m2 - o.m2) <= 1e-9 * o.m2, (rank, method, st.sum, o.sum)
    xs = rows_all["amount"][idx]
    assert abs(sw.sum - float(np.sum(xs[(xs >= 100.0) & (xs <= 500.0)]))) <= 1e-9 * abs(sw.sum)
del tsamp
# ---- SQL-string path across the GPUs: facts / integer accumulators all-gathered over NCCL, merged exactly ----
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import SqlError
from sql_helpers import MODE_OF, REL, engine_rows, rows_close
tq = sharded.ShardedTable.synthetic(n_small, rank, world, seed=11, device=local)
single = aqe.Engine(local).from_rows(rows_all) if rank == 0 else None
for sql, p, mode in (("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10, "run_query_with_ci"),
                     ("SELECT AVG(amount) FROM sales GROUP BY region", 0, "run_query_groupby"),
                     ("SELECT SUM(amount) FROM sales WHERE region >= 2 GROUP BY region", 20, "run_query_groupby_with_ci"),
                     ("SELECT COUNT(amount) FROM sales WHERE id <= 400 GROUP BY product_id", 10, "run_query_groupby"),
                     ("SELECT SUM(amount) FROM sales GROUP BY product_id", 50, "run_query_groupby"),
                     ("SELECT SUM(timestamp) FROM sales", 7, "run_query")):
    try:
        got = engine_rows(tq.sql(sql, p, MODE_OF[mode]))
    except ValueError:
        got = "stod"
    try:
        want = O.sql(rows_all, sql, p, mode) if rank == 0 else None
    except SqlError as ex:
        assert ex.kind == "stod" and got == "stod", (sql, ex)
        continue
    if rank == 0:
        assert rows_close(got, want, REL) is None, (sql, rows_close(got, want, REL))
        one = single.sql(sql, p, MODE_OF[mode])          # the unsharded table on one GPU: identical bits
        assert got == [(r.key, r.value, r.ci_lower, r.ci_upper) for r in one], sql
# the same queries with the accumulators exchanged INSIDE the scan kernel (NVLink mailboxes): identical bits, every rank
plain = {}
SQLX = (("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", 10, "ci_reference"), ("SELECT AVG(amount) FROM sales GROUP BY region", 0, "value"),
        ("SELECT SUM(amount) FROM sales WHERE region >= 2 GROUP BY region", 20, "ci_reference"), ("SELECT SUM(amount) FROM sales GROUP BY product_id", 50, "value"),
        ("SELECT COUNT(*) FROM sales", 0, "value"), ("SELECT COUNT(amount) FROM sales WHERE amount < 0 GROUP BY region", 0, "value"),
        ("SELECT COUNT(amount) FROM sales WHERE id <= 400 GROUP BY product_id", 10, "value"), ("SELECT SUM(timestamp) FROM sales WHERE (region = 1 OR region = 5)", 7, "value"))
for sql, p, mode in SQLX:
    try:
        plain[sql] = [(r.key, r.count, r.value, r.ci_lower, r.ci_upper, r.is_null) for r in tq.sql(sql, p, mode)]
    except Exception as ex:  # noqa: BLE001
        plain[sql] = repr(ex)
assert tq.enable_fused_exchange()
for it in range(3):
    for sql, p, mode in SQLX:
        try:
            got = [(r.key, r.count, r.value, r.ci_lower, r.ci_upper, r.is_null) for r in tq.sql(sql, p, mode)]
        except Exception as ex:  # noqa: BLE001
            got = repr(ex)
        assert got == plain[sql] or (isinstance(got, list) and all(a[:2] == b[:2] and (a[2:] == b[2:] or a[5]) for a, b in zip(got, plain[sql]))), (rank, it, sql, got[:3], plain[sql][:3])
big_sql = {}
for sql in ("SELECT SUM(amount) FROM sales", "SELECT SUM(amount) FROM sales WHERE timestamp > 1700000005"):
    tq_big = t.sql(sql)            # the big table holds id, amount, timestamp
    big_sql[sql] = tq_big[0].value
assert big_sql["SELECT SUM(amount) FROM sales"] == ref[0].sum or abs(big_sql["SELECT SUM(amount) FROM sales"] - ref[0].sum) <= 4e-16 * ref[0].sum
# every rank holds the identical result; latency of the fused multi-GPU estimator on the big table (configs[3])
approx_lat = {}
import time
for design, eps in (("srs", 0.5), ("block", 0.5), ("srs", 0.1)):
    lat = []
    for seed in range(40):
        torch.cuda.synchronize(); dist.barrier()
        t1 = time.perf_counter(); r = t.approx("sum", error_percent=eps, design=design, seed=seed); lat.append((time.perf_counter() - t1) * 1e6)
    assert r.status == 0 and abs(r.estimate - ref[0].sum) / ref[0].sum < 3 * eps / 100
    approx_lat[f"{design}_sum_{eps}pct_us_p50"] = sorted(lat)[len(lat) // 2]
    approx_lat[f"{design}_sum_{eps}pct_kernel_us"] = r.elapsed_us
    approx_lat[f"{design}_sum_{eps}pct_rows_read"] = r.n_samples

# sampled aggregates over the big table through the sampler plans (configs[3] "block sampling + parallel fast/slow method")
sampler_lat = {}
for method, pct, kw in (("memory_stride", 1.0, {}), ("block", 1.0, {}), ("parallel_block", 1.0, {}), ("parallel_pointer", 1.0, {})):
    prm = aqe.make_params(method, pct, **kw)
    pl = t.plan(method, prm)
    lat = []
    for it in range(12):
        torch.cuda.synchronize(); dist.barrier()
        t1 = time.perf_counter(); st = t.stats(plan=pl); lat.append((time.perf_counter() - t1) * 1e3)
    est = st.sum * (N / st.n)
    assert abs(est - ref[0].sum) / ref[0].sum < 0.01, (method, est, ref[0].sum)
    sampler_lat[f"{method}_{pct}pct"] = {"samples": st.n, "ms_p50": sorted(lat)[len(lat) // 2], "estimate_rel_error": abs(est - ref[0].sum) / ref[0].sum}

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
merged = torch.zeros(8, dtype=torch.int64, device="cuda")
partial = torch.zeros(8, dtype=torch.int64, device="cuda"); gathered = torch.zeros(8 * world, dtype=torch.int64, device="cuda")
host = torch.zeros(8 * world, dtype=torch.int64).pin_memory()


def fused():
    t.engine.scan_exchange_async(merged.data_ptr(), "amount", "amount", 100.0, 500.0, stream=stream.cuda_stream)
    host[:8].copy_(merged, non_blocking=True)


def nccl():
    t.engine.scan_async(partial.data_ptr(), "amount", "amount", 100.0, 500.0, stream=stream.cuda_stream)
    dist.all_gather_into_tensor(gathered, partial)
    host.copy_(gathered, non_blocking=True)


def kernel_only():
    t.engine.scan_async(partial.data_ptr(), "amount", "amount", 100.0, 500.0, stream=stream.cuda_stream)


def timed(fn, k=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


res = {"world": world, "rows_total": N, "checks": "fused == nccl bytes, closed forms, approx == orc_approx_sharded, samplers == oracle, SQL == oracle / single GPU: all passed",
       "approx_fused": approx_lat, "samplers_sharded": sampler_lat, "ms_kernel_only": timed(kernel_only), "ms_nccl_allgather": timed(nccl), "ms_fused_exchange": timed(fused)}
torch.cuda.synchronize()
f = aqe.Partial.from_buffer_copy(host[:8].numpy().tobytes())     # async form: no moments, same count / sum / comp bits
assert (f.count, f.sum, f.comp) == (ref[1].count, ref[1].sum, ref[1].comp)
if rank == 0:
    real_stdout.write(json.dumps(res) + "\n")
    real_stdout.flush()
dist.barrier()
dist.destroy_process_group()
