#!/usr/bin/env python3
"""First-load (cold: fresh process, staging buffers not allocated yet) and second-load (warm) time of a record file through
the drop-in module's open_database, for ingest chunk sizes AQE_INGEST_CHUNK_MB in {2, 4, 8, 16} -- what the reference's CLI pays per
invocation (every run is a new process).   python tools/ingest_cold.py [rows] > profiles/rN_ingest_cold.json"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
CHILD = r'''
import sys, time, json
sys.path.insert(0, %r)
import approximatequeryengine_b200 as aqe
b = aqe.backend()
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")       # the CUDA context exists (a caller's process has one): time the load, not cuInit
out = {}
for name in ("cold", "warm", "warm2"):
    db = b.CustomBPlusDB(0)
    t0 = time.perf_counter(); ok = db.open_database(%r); out[name + "_s"] = time.perf_counter() - t0
    assert ok and db.get_total_records() == %d
    t0 = time.perf_counter(); v = db.sum_amount_where(100.0, 500.0); out[name + "_first_query_ms"] = (time.perf_counter() - t0) * 1e3
    del db
print(json.dumps(out))
'''
tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
res = {"rows": n, "file_GB": 32 * n / 1e9, "medium": tmpdir or "tmp", "runs": []}
with tempfile.TemporaryDirectory(dir=tmpdir) as td:
    path = os.path.join(td, "sales.aqe")
    import approximatequeryengine_b200 as aqe
    aqe.Engine(0).generate(n, seed=7).save_file(path)
    for mb in (16, 8, 4, 2, 16):
        env = dict(os.environ, AQE_INGEST_CHUNK_MB=str(mb), AQE_DEVICE="0")
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, path, n)], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = {"error": r.stderr[-400:]}
        d["chunk_MB"] = mb
        for k in ("cold", "warm", "warm2"):
            if k + "_s" in d:
                d[k + "_GBps"] = 32 * n / d[k + "_s"] / 1e9
        res["runs"].append(d)
        print(json.dumps(d), file=sys.stderr, flush=True)
print(json.dumps(res, indent=1))
