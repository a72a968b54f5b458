"""Multi-GPU parity as a pytest module (`-m gpu`; skipped below two GPUs -- the 1-GPU box runs the colocated-shard checks of
tests/test_gpu_sharded.py instead).  Runs tests/multi_gpu_check.py under torchrun, one rank per GPU: fused in-kernel exchange
== NCCL all-gather + host merge bit for bit, closed forms at 1 B rows, the fused multi-GPU estimator == its CPU restatement
(orc_approx_sharded), the sampler families across ranks == the oracle, the SQL path == the oracle and == one GPU.  The JSON
line the check prints is the pass record (copied to gpurun_out/ when that directory exists)."""
import json
import os
import subprocess
import sys

import pytest

import approximatequeryengine_b200 as aqe

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_check_under_torchrun():
    c = aqe.C.c_int()
    aqe.lib().aqe_device_count(aqe.C.byref(c))
    if c.value < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU)")
    world = min(c.value, 8)
    env = dict(os.environ, AQE_CHECK_ROWS=os.environ.get("AQE_CHECK_ROWS", "1000000007"))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "AQE_DEVICE", "AQE_DEVICES", "AQE_MIN_SHARD_ROWS"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["world"] == world and "all passed" in rec["checks"]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"multi_gpu_check_n{world}.json"), "w") as f:
            json.dump(rec, f, indent=1)
