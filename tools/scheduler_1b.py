import sys, time, json
sys.path.insert(0, '/root/repo')
import approximatequeryengine_b200 as aqe
b = aqe.backend()
s = b.CustomApproximateScheduler(0.05)
s.db.generate_synthetic(1_000_000_000, 7, 0, 0, 0b00010)   # amount only
out = {}
t0 = time.perf_counter(); ex = s.execute_exact_sum(); out["exact_sum_ms"] = (time.perf_counter() - t0) * 1e3
for p in (1.0, 10.0):
    t0 = time.perf_counter(); r = s.execute_sum_query("SELECT SUM(amount) FROM sales", p, 4); dt = (time.perf_counter() - t0) * 1e3
    out[f"sum_query_{p}pct"] = {"ms": dt, "rel_err_pct": abs(r.value - ex.value) / ex.value * 100, "samples_used": r.samples_used}
    t0 = time.perf_counter(); r = s.execute_sum_query("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500", p, 4); dt = (time.perf_counter() - t0) * 1e3
    out[f"sum_where_query_{p}pct_ms"] = dt
print(json.dumps(out))
