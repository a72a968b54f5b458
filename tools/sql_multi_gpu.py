#!/usr/bin/env python3
"""SQL path across GPUs (run under torchrun, one rank per GPU): per-query latency of ShardedTable.sql over a table of
AQE_ROWS rows in total -- grouped scan on every shard + all-gather of the integer accumulators + exact merge + finish."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from approximatequeryengine_b200 import sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
sys.stdout.flush()
real_stdout = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(os.environ.get("AQE_ROWS", 1_000_000_000))
t = sharded.ShardedTable.synthetic(N, rank, world, seed=7, device=local, columns=("id", "amount", "region", "product_id"))
fused = os.environ.get("AQE_SQL_FUSED", "1") != "0" and world > 1 and t.enable_fused_exchange()
out = {"world": world, "rows_total": N, "exchange": "inside the scan kernel (NVLink mailboxes)" if fused else "NCCL all-gather + host merge", "queries": []}
for sql, p, mode in (("SELECT SUM(amount) FROM sales", 0, "value"), ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "value"),
                     ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", 10, "ci_reference"),
                     ("SELECT AVG(amount) FROM sales GROUP BY product_id", 0, "value")):
    t.sql(sql, p, mode)
    ts = []
    for _ in range(30):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); rows = t.sql(sql, p, mode); ts.append(time.perf_counter() - t0)
    local_only = []
    q = sharded.sql_parse(sql, p)
    from approximatequeryengine_b200 import SQL_MOMENTS, sql_layout
    layout = sql_layout(q, t._sql_facts[(q.group_col, q.agg_col)])
    for _ in range(30):
        t0 = time.perf_counter(); t.engine.sql_scan(q, layout, SQL_MOMENTS if mode != "value" else 0); local_only.append(time.perf_counter() - t0)
    ms = torch.tensor([statistics.median(ts) * 1e3, statistics.median(local_only) * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["queries"].append({"sql": sql, "sample_percent": p, "mode": mode, "groups": len(rows), "ms_sharded_query": float(ms[0]), "ms_local_scan_only": float(ms[1]),
                           "records_per_s": N / (float(ms[0]) * 1e-3)})
if rank == 0:
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()
dist.barrier()
dist.destroy_process_group()
