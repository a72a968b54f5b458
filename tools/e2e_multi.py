#!/usr/bin/env python3
"""End-to-end scan of ONE pinned host column through every GPU of the box from one process (aqe_scan_host_column_multi: chunks handed
out from one counter) checked against the resident scan and against the same column cut into equal shards.
No torch: ctypes on the C-ABI only.  python tools/e2e_multi.py [rows] [steps] [chunk MiB ...] > out.json"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LO, HI, SEED = 100.0, 500.0, 7


def child(n, steps):
    import approximatequeryengine_b200 as aqe
    L = aqe.lib()
    c = C.c_int(0)
    L.aqe_device_count(C.byref(c))
    G = c.value
    hp = C.c_void_p()
    aqe.check(L.aqe_host_alloc(n * 8, C.byref(hp)))
    bounds = [n * g // G for g in range(G + 1)]
    whole = []
    for g in range(G):
        e = aqe.Engine(g).generate(bounds[g + 1] - bounds[g], seed=SEED, first_row=bounds[g], columns=("amount",))
        e.read_column("amount", out_ptr=hp.value + bounds[g] * 8)
        whole.append(e.scan("amount", "amount", LO, HI))
        e.close()
    want = aqe.merge_partials(whole)
    devs = list(range(G))

    def multi():
        return aqe.host_scan_column(None, LO, HI, use_pred=True, device=devs, ptr=hp.value, n=n, kind=0)

    def equal_shards():   # the same column as equal shards, one device after the other (bench.py times that form with one process per GPU)
        return aqe.merge_partials([aqe.host_scan_column(None, LO, HI, use_pred=True, device=g, ptr=hp.value + bounds[g] * 8, n=bounds[g + 1] - bounds[g], kind=0)
                                   for g in range(G)])

    res = {"gpus": G, "rows": n, "chunk_mb": int(os.environ.get("AQE_E2E_CHUNK_MB", 64))}
    for _ in range(3):
        p = multi()
    res["equal_to_resident_scan"] = bool(p.count == want.count and abs(p.sum - want.sum) <= 1e-12 * abs(want.sum))
    q = equal_shards()
    res["equal_shards_same_count"] = bool(q.count == p.count)
    sums = set()
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        p = multi()
        ts.append(time.perf_counter() - t)
        sums.add((p.count, p.sum))
    ts.sort()
    res.update({"same_bits_every_step": len(sums) == 1, "ms_p50": ts[len(ts) // 2] * 1e3, "ms_min": ts[0] * 1e3,
                "h2d_GBps_p50": n * 8 / ts[len(ts) // 2] / 1e9, "records_per_s_p50": n / ts[len(ts) // 2]})
    print(json.dumps(res))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    if os.environ.get("AQE_E2E_MULTI_CHILD"):
        return child(n, steps)
    chunks = [int(x) for x in sys.argv[3:]] or [64]
    out = []
    for mb in chunks:
        env = dict(os.environ, AQE_E2E_CHUNK_MB=str(mb), AQE_E2E_MULTI_CHILD="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(steps)], env=env, capture_output=True, text=True)
        out.append(json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"chunk_mb": mb, "error": r.stderr[-1500:]})
        print(out[-1], file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
