#!/usr/bin/env python3
"""Sweep the scan-kernel variants / launch geometry on one B200 and print GB/s (8 algorithmic bytes per record).
Usage (GPU box): python tools/scan_sweep.py [records]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import approximatequeryengine_b200 as aqe

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
eng = aqe.Engine(0).generate(n, seed=7, columns=("amount",))
partial = torch.zeros(8, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)


def run(env, pred=True, reps=20):
    for k in ("AQE_SCAN_VARIANT", "AQE_SCAN_BPS", "AQE_SCAN_UNROLL", "AQE_SCAN_STAGES", "AQE_SCAN_CHUNK_KB"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    args = ("amount", "amount", 100.0, 500.0) if pred else ("amount", None, 0.0, 0.0)
    for _ in range(3):
        eng.scan_async(partial.data_ptr(), *args, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        eng.scan_async(partial.data_ptr(), *args, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    p = aqe.Partial.from_buffer_copy(partial.cpu().numpy().tobytes())
    return {"env": env, "pred": pred, "ms": round(ms, 4), "GBps": round(8 * n / ms / 1e6, 1), "count": p.count, "sum": p.sum}


configs = []
for bps in (0, 1, 2, 3, 4, 6, 8):
    configs.append({"AQE_SCAN_VARIANT": 0, "AQE_SCAN_BPS": bps})
for u in (1, 2, 8):
    for bps in (0, 4, 8):
        configs.append({"AQE_SCAN_VARIANT": 0, "AQE_SCAN_UNROLL": u, "AQE_SCAN_BPS": bps})
for u in (2, 4, 8):
    for bps in (0, 4):
        configs.append({"AQE_SCAN_VARIANT": 1, "AQE_SCAN_UNROLL": u, "AQE_SCAN_BPS": bps})
for st, ck in ((4, 16), (8, 16), (4, 32), (6, 32), (8, 8)):
    for bps in (0, 1, 2):
        configs.append({"AQE_SCAN_VARIANT": 2, "AQE_SCAN_STAGES": st, "AQE_SCAN_CHUNK_KB": ck, "AQE_SCAN_BPS": bps})
ref = None
for c in configs:
    for pred in (True, False):
        r = run(c, pred)
        key = (r["count"], r["sum"])
        print(json.dumps(r), flush=True)
