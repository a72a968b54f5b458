#!/usr/bin/env python3
"""One short pass over the kernels the round-2 ncu capture targets (run plain first, then under
`ncu --set full -k regex:'k_scan_ring|k_approx|k_plan_stats|k_plan_gather|k_aos_to_soa'`):

  k_scan_ring     the strong-scaling shard shape: 125 M rows, SUM+COUNT WHERE amount BETWEEN 100 AND 500
  k_approx        BASELINE configs[1]: 10 M rows, APPROX AVG at 1 %
  k_plan_stats    memory_stride 1 % (random sectors) and block 1 % (contiguous tiles) over the big table
  k_plan_gather   memory_stride 1 % -> 32-byte rows
  k_aos_to_soa    ingest of 1 M host rows (2 chunks)

    python tools/profile_targets.py [big_rows]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import approximatequeryengine_b200 as aqe

big = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000

e = aqe.Engine(0).generate(125_000_000, seed=7, columns=("amount",))
for _ in range(3):
    p = e.scan("amount", "amount", 100.0, 500.0)
print("scan 125M", p.count, p.sum, flush=True)
e.close()

e = aqe.Engine(0).generate(10_000_000, seed=7, columns=("amount",))
for s in range(3):
    r = e.approx("avg", error_percent=1.0, seed=s)
print("approx 10M", r.estimate, r.n_samples, r.rounds, r.elapsed_us, flush=True)
e.close()

e = aqe.Engine(0).generate(big, seed=7)
for method in ("memory_stride", "block"):
    pl = e.plan(method, aqe.make_params(method, 1.0))
    for _ in range(2):
        st = e.stats(pl)
    print("stats", method, st.n, st.mean, flush=True)
pl = e.plan("memory_stride", aqe.make_params("memory_stride", 0.5))
rows = e.gather(pl)
print("gather", len(rows), int(rows["id"][-1]), flush=True)
e.close()

host = aqe.synth_rows_host(1_000_000, seed=7)
e = aqe.Engine(0).from_rows(host)
print("ingest", e.count, e.sum_int("id"), flush=True)
e.close()
