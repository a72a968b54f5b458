"""Range-sharded tables: one process per GPU, contiguous record ranges, tiny partials merged in rank order.

The reference has no multi-process story; its own region split is ``[N*t/T, N*(t+1)/T)`` per thread
(custom_bplus_db.cpp:925-926, 1904-1918) and that is the split used here per rank.  The only exchange is
one all-gather of a 64-byte ``aqe_partial`` (exact scans) or a 96-byte ``aqe_approx_result`` (sampled
estimates) per query -- ``torch.distributed`` carries it (NCCL over NVLink on GPUs, gloo in the CPU
tests); the merge itself is the fixed-rank-order host code of the C-ABI (``aqe_merge_partials`` /
``aqe_approx_merge``), so every rank computes bit-identical results.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import (AGG, SQL_MOMENTS, SQL_UNSAMPLED, ApproxResult, Engine, Partial, Plan, SampleParams, SqlFacts, Stats, StatsPartial,
               build_plan, check, lib, merge_stats, sql_finish, sql_layout, sql_merge, sql_parse)


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [a, b) owned by `rank` (custom_bplus_db.cpp:925-926 with t=rank, T=world)."""
    return n * rank // world, n * (rank + 1) // world


def _struct_to_i64(obj, words: int) -> np.ndarray:
    return np.frombuffer(bytes(obj), dtype=np.int64, count=words).copy()


def _i64_to_struct(arr: np.ndarray, cls):
    return cls.from_buffer_copy(np.ascontiguousarray(arr, dtype=np.int64).tobytes())


def allgather_struct(obj, cls, group=None, device=None):
    """All-gather one POD struct (as raw 8-byte words so no bit is altered) -> list in rank order."""
    import torch
    import torch.distributed as dist

    words = C.sizeof(cls) // 8
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    mine = torch.from_numpy(_struct_to_i64(obj, words)).to(dev)
    out = torch.empty(world * words, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = out.cpu().numpy().reshape(world, words)
    return [_i64_to_struct(flat[r], cls) for r in range(world)]


def allgather_words(arr: np.ndarray, group=None, device=None) -> np.ndarray:
    """All-gather a fixed-length uint64 array (as int64 words) -> [world, len] in rank order."""
    import torch
    import torch.distributed as dist

    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arr.reshape(1, -1)
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    mine = torch.from_numpy(arr.view(np.int64).copy()).to(dev)
    out = torch.empty(world * len(arr), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy().view(np.uint64).reshape(world, len(arr))


def merge_partials(parts, is_integer: bool = False) -> Partial:
    arr = (Partial * len(parts))(*parts)
    out = Partial()
    check(lib().aqe_merge_partials(arr, len(parts), int(is_integer), C.byref(out)))
    return out


def merge_approx(parts, agg: str, confidence_level: float) -> ApproxResult:
    arr = (ApproxResult * len(parts))(*parts)
    out = ApproxResult()
    check(lib().aqe_approx_merge(arr, len(parts), AGG[agg], confidence_level, C.byref(out)))
    return out


class ShardedTable:
    """This rank's shard + the collective merge.  Every query is called by all ranks (SPMD)."""

    def __init__(self, engine: Engine, total_rows: int, first_row: int, group=None):
        self.engine = engine
        self.total_rows = total_rows
        self.first_row = first_row
        self.group = group

    @classmethod
    def synthetic(cls, total_rows: int, rank: int, world: int, seed: int = 7, device: int | None = None,
                  columns=("id", "amount", "region", "product_id", "timestamp"), dist: int = 0, group=None):
        a, b = shard_range(total_rows, rank, world)
        e = Engine(device).generate(b - a, seed=seed, first_row=a, dist=dist, columns=columns)
        return cls(e, total_rows, a, group)

    @classmethod
    def from_file(cls, path: str, rank: int, world: int, device: int | None = None, group=None):
        import struct
        with open(path, "rb") as f:
            total = struct.unpack("<QQQ", f.read(24))[2]
        a, b = shard_range(total, rank, world)
        e = Engine(device).load_file(path, first_row=a, n_rows=b - a)
        return cls(e, total, a, group)

    def enable_fused_exchange(self) -> bool:
        """Map every rank's mailbox over CUDA IPC so scans merge inside the kernel (NVLink peer stores) instead of
        through an NCCL all-gather.  Needs one process per GPU on one box; returns False (and keeps the
        all-gather path) for a single rank."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return False
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        mine = self.engine.exchange_init(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=self.group)
        self.engine.exchange_connect(handles)
        self.engine.exchange_set_total_rows(self.total_rows)
        dist.barrier(self.group)
        self.fused = True
        return True

    def scan(self, agg_col="amount", pred_col=None, lo=0.0, hi=0.0) -> Partial:
        if getattr(self, "fused", False):
            return self.engine.scan_exchange(agg_col, pred_col, lo, hi)
        local = self.engine.scan(agg_col, pred_col, lo, hi) if self.engine.count else Partial(minv=float("inf"), maxv=float("-inf"))
        parts = allgather_struct(local, Partial, self.group)
        return merge_partials(parts, is_integer=agg_col != "amount")

    def sum_amount(self) -> float:
        return self.scan("amount").sum

    def sum_amount_where(self, lo: float, hi: float):
        p = self.scan("amount", "amount", lo, hi)
        return p.sum, p.count

    def count(self) -> int:
        return self.total_rows

    def sql(self, query: str, sample_percent: int = 0, mode: str = "value"):
        """run_query* over the sharded table (executor.cpp:28-338): every rank scans its shard into the order-independent
        128-bit integer accumulators of the grouped-scan kernel, the accumulators are all-gathered and added with carry
        (exact: the result does not depend on the shard count), and every rank applies the reference's arithmetic.
        Exchanges: the 32-byte column facts (so that all shards use one key range and one fixed-point scale; once per
        column pair, cached), the accumulators, and -- only when a group has no sampled row anywhere -- the unsampled
        group counts."""
        q = sql_parse(query, sample_percent)
        # column facts depend on (GROUP BY column, aggregate column) and the table only: exchanged once, then cached
        cache = self.__dict__.setdefault("_sql_facts", {})
        key = (q.group_col, q.agg_col)
        if key not in cache:
            cache[key] = allgather_struct(self.engine.sql_facts(q), SqlFacts, self.group)
        layout = sql_layout(q, cache[key])
        grouped = q.group_col >= 0
        step = 0 if (sample_percent <= 0 or sample_percent >= 100) else max(1, 100 // sample_percent)
        if mode == "value":
            moments = False
        elif grouped:
            moments = mode == "ci_reference" or q.agg != AGG["count"]
        else:
            moments = q.agg != AGG["count"] and step > 0
        def table_level(flags):
            if getattr(self, "fused", False):   # merged inside the scan kernel over the NVLink mailboxes
                return self.engine.sql_scan(q, layout, flags, exchange=True)
            total = np.zeros(layout.n_groups * 5, dtype=np.uint64)
            for part in allgather_words(self.engine.sql_scan(q, layout, flags), self.group):
                sql_merge(total, part)
            return total
        acc = table_level(SQL_MOMENTS if moments else 0)
        exists = None
        if grouped and step > 1 and (acc[0::5] == 0).any():   # identical on every rank: the pass below is collective
            exists = table_level(SQL_UNSAMPLED)
        return sql_finish(q, layout, acc, mode, exists)

    # ---- the legacy sampler families across the shards (BASELINE configs[3]: "block sampling + parallel fast/slow method") ----
    def plan(self, method: str, params: SampleParams) -> Plan:
        """The sampler's position list over the WHOLE table (closed-form segments; identical on every rank, nothing is
        exchanged).  Samplers that read the table to place their samples (adaptive_block, clt_validated_dual_pointer) need the
        whole table behind one handle: Engine(devices=...) / CustomBPlusDB(devices)."""
        return build_plan(self.total_rows, method, params)

    def stats(self, method: str | None = None, params: SampleParams | None = None, col="amount", where=None, where_col="amount",
              plan: Plan | None = None) -> Stats:
        """Moments of `col` over the sampler's rows.  Every rank walks the whole plan and gathers the positions inside its own
        row range (k_plan_stats with a window), the 64-byte partial sums are all-gathered and merged in rank order
        (aqe_stats_merge: double-double sum, Chan's update for M2) -- n and the sum are those of the one-GPU plan, bit for bit /
        to the last ulp.  Replaces the per-thread region loops of parallel_pointer_sample (custom_bplus_db.cpp:814-854),
        parallel_block_sample (:1218-1271), clt/optimized_clt regions (:925-926, :1046-1147), memory_stride_sample (:1526)."""
        plan = plan or self.plan(method, params)
        part = self.engine.stats_window(plan, self.first_row, col, where, where_col)
        return merge_stats(allgather_struct(part, StatsPartial, self.group))

    def gather(self, method: str | None = None, params: SampleParams | None = None, plan: Plan | None = None) -> np.ndarray:
        """The sampled rows in plan order, on every rank (the list[Record] return path).  Each rank fills the slots of the
        positions it owns and leaves the others zero; one all-reduce (sum of the 64-bit words: x + 0 + ... + 0) assembles them."""
        import torch
        import torch.distributed as dist
        plan = plan or self.plan(method, params)
        rows, _ = self.engine.gather_window(plan, self.first_row)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return rows
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        words = torch.from_numpy(rows.view(np.int64).copy()).to(dev)
        dist.all_reduce(words, op=dist.ReduceOp.SUM, group=self.group)
        return words.cpu().numpy().view(rows.dtype)

    def approx(self, agg="sum", error_percent=1.0, confidence_level=0.95, seed=0, **kw) -> ApproxResult:
        """Shards are strata: each rank runs its persistent CLT kernel to the same relative target with
        an independent Philox key (seed, rank); totals and variances add (aqe_approx_merge)."""
        import torch.distributed as dist
        if getattr(self, "fused", False):
            # one global stop rule: per-look moments exchanged inside the persistent kernel (NVLink mailboxes)
            return self.engine.approx(agg, error_percent, confidence_level, seed=seed, exchange=True, **kw)
        rank = dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0
        local = self.engine.approx(agg, error_percent, confidence_level, seed=(seed << 8) + rank, **kw)
        parts = allgather_struct(local, ApproxResult, self.group)
        return merge_approx(parts, agg, confidence_level)
