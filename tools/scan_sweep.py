#!/usr/bin/env python3
"""Sweep the scan-kernel variants / launch geometry on one B200 and print GB/s (8 algorithmic bytes per record).
Usage (GPU box): python tools/scan_sweep.py [records] [round]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import approximatequeryengine_b200 as aqe

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "2"
eng = aqe.Engine(0).generate(n, seed=7, columns=("amount",))
partial = torch.zeros(8, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
KEYS = ("AQE_SCAN_VARIANT", "AQE_SCAN_BPS", "AQE_SCAN_UNROLL", "AQE_SCAN_STAGES", "AQE_SCAN_CHUNK_KB", "AQE_SCAN_MINB")


def run(env, pred=True, reps=20):
    for k in KEYS:
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


configs = [{"AQE_SCAN_VARIANT": 0}]
for st in (2, 3, 4, 6):
    for bps in (1, 2, 3):
        configs.append({"AQE_SCAN_VARIANT": 2, "AQE_SCAN_STAGES": st, "AQE_SCAN_BPS": bps})
configs += [{"AQE_SCAN_VARIANT": 4}, {"AQE_SCAN_VARIANT": 4, "AQE_SCAN_UNROLL": 6}, {"AQE_SCAN_VARIANT": 4, "AQE_SCAN_MINB": 3}, {"AQE_SCAN_VARIANT": 1}]
configs.append({"AQE_SCAN_VARIANT": 0})
for c in configs:
    for pred in (True, False):
        print(json.dumps(run(c, pred)), flush=True)
