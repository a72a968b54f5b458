"""The N>1 host path on CPU: world_size-2 (and 3) gloo process groups.  Each rank computes the partial of its
contiguous shard with the ORACLE (no GPU here), the product's host layer (approximatequeryengine_b200.sharded)
all-gathers the 64-byte partials and folds them in rank order with the C-ABI's aqe_merge_partials /
aqe_approx_merge, and every rank must end up with the oracle's whole-table answer."""
import math
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    import approximatequeryengine_b200 as aqe
    from approximatequeryengine_b200 import sharded
    from oracle import ApproxSpec, Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = Oracle()
        a, b = sharded.shard_range(n, rank, world)
        rows = O.synth(b - a, seed=7, first_row=a)          # this rank's shard of the synthetic table
        out = {}
        for name, (agg, pred, lo, hi) in {"sum": ("amount", None, 0, 0), "where": ("amount", "amount", 100.0, 500.0),
                                          "ts": ("timestamp", "region", 1, 6), "id": ("id", None, 0, 0)}.items():
            o = O.scan(rows, agg, pred, lo, hi)
            local = aqe.Partial(count=o.count, sum=o.sum, comp=0.0, isum_lo=o.isum_lo, isum_hi=o.isum_hi, sumsq=o.sumsq, minv=o.minv, maxv=o.maxv)
            parts = sharded.allgather_struct(local, aqe.Partial)
            assert len(parts) == world and parts[rank].count == o.count and parts[rank].sum == o.sum
            m = sharded.merge_partials(parts, is_integer=agg != "amount")
            out[name] = (m.count, m.sum, m.comp, m.isum, m.minv, m.maxv)
        # sampled estimates: shards are strata
        r = O.approx(rows, ApproxSpec(agg=0, design=0, agg_col=1, pred_col=-1, error_percent=1.0, confidence_level=0.95, seed=(5 << 8) + rank))
        local = aqe.ApproxResult.from_buffer_copy(bytes(r))
        parts = sharded.allgather_struct(local, aqe.ApproxResult)
        m = sharded.merge_approx(parts, "sum", 0.95)
        out["approx"] = (m.estimate, m.ci_lower, m.ci_upper, m.n_samples, m.population, m.status)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_over_gloo(oracle, world):
    n = 300_007
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = oracle.synth(n, seed=7)
    assert all(got[r] == got[0] for r in range(world))       # every rank computed identical bits
    g = got[0]
    whole = oracle.scan(rows, "amount")
    assert g["sum"][0] == n and abs(g["sum"][1] - math.fsum(rows["amount"])) <= 1e-12 * whole.sum
    assert (g["sum"][4], g["sum"][5]) == (whole.minv, whole.maxv)
    w = oracle.scan(rows, "amount", "amount", 100.0, 500.0)
    assert g["where"][0] == w.count and abs(g["where"][1] - w.sum) <= 1e-12 * w.sum
    assert g["ts"][3] == oracle.scan(rows, "timestamp", "region", 1, 6).isum and g["ts"][0] == oracle.scan(rows, "timestamp", "region", 1, 6).count
    assert g["id"][3] == n * (n + 1) // 2
    est, lo, hi, ns, pop, status = g["approx"]
    assert pop == n and status == 0 and lo < est < hi
    assert abs(est - whole.sum) / whole.sum < 0.02 and (hi - lo) / 2 / est <= 0.011


def test_shard_ranges_tile_the_table():
    from approximatequeryengine_b200.sharded import shard_range
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1
