#!/usr/bin/env python3
"""A/B of programmatic dependent launch on back-to-back exact scans (ScanArgs::pdl_tail, AQE_SCAN_PDL=0/1): ms per query in a
short burst and in a 2 s sustained loop (with SM clock / power sampled), at the strong-scaling shard size (125 M rows) and at 1 B.
    python tools/pdl_ab.py > profiles/rN_pdl_ab.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import approximatequeryengine_b200 as aqe
from bench import ClockSampler

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
host_out = torch.zeros(8, dtype=torch.int64).pin_memory()


def loop(eng, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        eng.scan_async(host_out.data_ptr(), "amount", "amount", 100.0, 500.0, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (125_000_000, 1_000_000_000):
    eng = aqe.Engine(0).generate(n, seed=7, columns=("amount",))
    ref = eng.scan("amount", "amount", 100.0, 500.0)
    for pdl in (0, 1, 0, 1):
        os.environ["AQE_SCAN_PDL"] = str(pdl)
        loop(eng, 5)
        burst = min(loop(eng, 200 if n < 500_000_000 else 40) for _ in range(3))
        p = aqe.Partial.from_buffer_copy(host_out.numpy().tobytes())
        assert (p.count, p.sum) == (ref.count, ref.sum), (pdl, p.count, ref.count, p.sum, ref.sum)
        clk = ClockSampler(0)
        clk.start()
        sustained = loop(eng, int(2500 / burst))
        c = clk.stop()
        print(json.dumps({"rows": n, "pdl": pdl, "burst_ms_per_query": round(burst, 5), "burst_GBps": round(8 * n / burst / 1e6, 1),
                          "sustained_2s_ms_per_query": round(sustained, 5), "sustained_GBps": round(8 * n / sustained / 1e6, 1),
                          "sm_mhz_under_load": c.get("sm_mhz"), "power_w_max": c.get("power_w_max"), "reasons": c.get("reasons"),
                          "kernel": aqe.lib().aqe_last_scan_kernel().decode()}), flush=True)
    if n < 500_000_000:   # the skewed tile schedule (ScanArgs::even_rounds, AQE_SCAN_SKEW) at the strong-scaling shard size
        os.environ["AQE_SCAN_PDL"] = "1"
        for skew in (0, 2, 4, 6, 8, 12, 16, 24, 0, 6):
            os.environ["AQE_SCAN_SKEW"] = str(skew)
            loop(eng, 5)
            burst = min(loop(eng, 300) for _ in range(3))
            p = aqe.Partial.from_buffer_copy(host_out.numpy().tobytes())
            assert (p.count, p.sum) == (ref.count, ref.sum), (skew, p.count, ref.count, p.sum, ref.sum)
            print(json.dumps({"rows": n, "pdl": 1, "skew_tiles": skew, "burst_ms_per_query": round(burst, 5), "burst_GBps": round(8 * n / burst / 1e6, 1)}), flush=True)
        os.environ.pop("AQE_SCAN_SKEW", None)
    eng.close()
