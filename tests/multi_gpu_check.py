#!/usr/bin/env python3
"""Multi-GPU check of the fused exchange (run under torchrun on an N-GPU box, N >= 2; not collected by pytest):

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


res = {"world": world, "rows_total": N, "ms_kernel_only": timed(kernel_only), "ms_nccl_allgather": timed(nccl), "ms_fused_exchange": timed(fused)}
torch.cuda.synchronize()
f = aqe.Partial.from_buffer_copy(host[:8].numpy().tobytes())     # async form: no moments, same count / sum / comp bits
assert (f.count, f.sum, f.comp) == (ref[1].count, ref[1].sum, ref[1].comp)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
