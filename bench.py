#!/usr/bin/env python3
"""bench.py -- the reference's headline benchmark on B200: exact SUM/COUNT(amount) with a range predicate
over the synthetic sales table (BASELINE.json configs[2]), plus the CLT-terminated APPROX latency
(configs[1]) as a secondary figure.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling weak|strong] [--impl reference]

N > 1 is launched by the driver under torchrun (one rank per GPU, NCCL).  A "step" is one pass of the hot
path over the rank's shard: the fused predicate scan kernel (k_scan, aqe_kernels.cuh) over the HBM-resident
amount column, one all-gather of the 64-byte partials, and an async copy of the gathered partials to
pinned host memory.  Rank 0 prints ONE JSON line.

  value     whole-job records/s with the table resident in HBM (CUDA events on the launching stream,
            barrier + synchronize on both sides, max over ranks)
  e2e       the same metric through the C-ABI with HOST buffers: aqe_scan_host_column() streams each rank's
            pinned host column through the device every step (H2D inside the timed region)
  roofline  the scan kernel alone: 8 algorithmic bytes per record / its CUDA-event duration vs the
            measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own CPU path (oracle/_ref = the unmodified reference compiled in place) on a
            bounded 1 M-record sample, same query, timed on this box's host cores (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LO, HI = 100.0, 500.0           # `amount BETWEEN 100 AND 500` (custom_scheduler.cpp:279), selectivity ~40 %
SEED = 7
METRIC = "exact_sum_count_where_records_per_sec"
UNIT = "records/s"
CPU_SAMPLE_ROWS = 1_000_000


def ncu_traffic_per_record():
    """DRAM bytes per record of the scan kernel from the committed ncu --set full capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)["k_scan"]
        return (t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]) / t["records_per_launch"]
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples in the upper half of the observed power range
        if sm:
            cut = (max(power) + min(power)) / 2 if power else 0
            load = [c for c, p in zip(sm, power) if p >= cut] or sm
            return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                    "power_w_max": max(power) if power else None}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own implementation of the query
# ---------------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, port_too: bool = True) -> dict:
    """sum_amount_where(100, 500) of the UNMODIFIED reference (oracle/_ref) over a 1 M-record sample.  The
    reference's exact scan is single-threaded by construction (custom_bplus_db.cpp:263-274), so cores = 1 is
    every thread it can use.  Loading goes through insert_batch (untimed; open_database deadlocks)."""
    import numpy as np
    from oracle import Oracle, Ref
    O = Oracle()
    rows = O.synth(CPU_SAMPLE_ROWS, seed=SEED)
    out = {"unit": UNIT}
    if Ref.available():
        t0 = time.time()
        R = Ref(rows)
        load_s = time.time() - t0
        R.time(1, max(1, warmup), LO, HI)
        per = []
        for _ in range(max(steps, 3)):
            t, v = R.time(1, 1, LO, HI)
            per.append(t)
        t_step = statistics.median(per)
        want, _ = O.sum_amount_where(rows, LO, HI)
        assert v == want, "reference and oracle disagree"
        out.update({"value": CPU_SAMPLE_ROWS / t_step, "cores": 1, "kind": "reference", "ms_per_step": t_step * 1e3,
                    "sample": f"{CPU_SAMPLE_ROWS} records (32-byte rows in the reference B+ tree, loaded by insert_batch in {load_s:.1f} s, untimed); "
                              f"CustomBPlusDB::sum_amount_where(100,500), median of {len(per)} calls, single-threaded as shipped"})
        del R
    else:
        per = []
        for _ in range(max(steps, 3)):
            t0 = time.perf_counter(); O.sum_amount_where(rows, LO, HI); per.append(time.perf_counter() - t0)
        t_step = statistics.median(per)
        out.update({"value": CPU_SAMPLE_ROWS / t_step, "cores": 1, "kind": "port", "ms_per_step": t_step * 1e3,
                    "sample": f"{CPU_SAMPLE_ROWS} records, oracle/aqe_oracle.c orc_sum_amount_where (oracle/_ref not built)"})
    if port_too:
        # the restated multithreaded scan, all host cores, for an honest upper bound of the CPU side
        cores = os.cpu_count() or 1
        n = 100_000_000
        col = np.empty(n, dtype=np.float64)
        blk = O.synth(1_000_000, seed=SEED)["amount"]
        for i in range(0, n, len(blk)):
            col[i:i + len(blk)] = blk
        best = min(_timeit(lambda: O.scan_mt(col, aos=False, threads=cores, pred=(LO, HI))) for _ in range(3))
        out["port_mt"] = {"value": n / best, "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": f"{n} records, 8-byte amount column (SoA), orc_scan_mt: contiguous regions x {cores} threads"}
        # the same scan over 32-byte rows (AoS) -- the bytes the reference's own loops touch per record
        from oracle import RECORD_DTYPE
        n2 = 50_000_000
        aos = np.zeros(n2, dtype=RECORD_DTYPE)
        aos["amount"] = col[:n2]
        best2 = min(_timeit(lambda: O.scan_mt(aos, aos=True, threads=cores, pred=(LO, HI))) for _ in range(3))
        out["port_mt_aos"] = {"value": n2 / best2, "unit": UNIT, "cores": cores, "kind": "port",
                              "sample": f"{n2} records, 32-byte rows (AoS), orc_scan_mt: contiguous regions x {cores} threads"}
    return out


def bind_to_gpu_numa_node(index: int):
    """Restrict this process to the CPUs NVML reports as local to GPU `index`, so that host buffers pinned afterwards sit
    on that GPU's NUMA node (one rank per GPU; on a two-socket box half the GPUs otherwise pull their column across the
    socket interconnect).  Returns a note for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no CPU affinity reported for the GPU"
        os.sched_setaffinity(0, cpus)
        return f"process bound to the {len(cpus)} CPUs local to GPU {index}"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__}: {e})"


def _timeit(f):
    t0 = time.perf_counter(); f(); return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup, port_too=False)
    cfg = workload_config(args, per_gpu_rows(args), cpu=True)
    # what this arm actually times: the reference cannot load more than a few million rows (O(N^2) load)
    cfg["records_timed_per_step"] = CPU_SAMPLE_ROWS
    cfg["note"] = (f"the reference arm times {CPU_SAMPLE_ROWS} records per step (a bounded sample of the workload: the reference's load is O(N^2)); "
                   "records_per_gpu / total_records describe the workload the B200 arm runs")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def scaling_of(args) -> str:
    """BASELINE.json configs[2] is ONE 1 B-record table sharded over 1/2/4/8 GPUs: strong scaling is the default for N > 1
    (`--scaling weak` keeps 1 B records per GPU as a secondary experiment)."""
    if args.scaling != "auto":
        return args.scaling
    return "strong" if args.gpus > 1 else "weak"


def shard_rows(args, rank: int):
    """(first row, rows) of this rank: [N r / G, N (r+1) / G) of the table (custom_bplus_db.cpp:925-926)."""
    if scaling_of(args) == "weak":
        return rank * args.records, args.records
    a, b = args.records * rank // args.gpus, args.records * (rank + 1) // args.gpus
    return a, b - a


def per_gpu_rows(args) -> int:
    return args.records if scaling_of(args) == "weak" else args.records // args.gpus


def total_rows(args) -> int:
    return args.records * args.gpus if scaling_of(args) == "weak" else args.records


def workload_config(args, rows_per_gpu, cpu=False):
    tot = total_rows(args)
    return {"workload": f"BASELINE.json configs[2]: {tot / 1e9:g}B-record synthetic sales table, exact SUM+COUNT(amount) WHERE amount BETWEEN 100 AND 500, " +
                        ("the whole table resident on one GPU" if args.gpus == 1 else
                         f"range-sharded across {args.gpus} GPU(s) ({scaling_of(args)} scaling: " +
                         ("the table grows with the GPU count, 1B records per GPU)" if scaling_of(args) == "weak" else "ONE table split over the GPUs)")),
            "records_per_gpu": rows_per_gpu, "total_records": tot, "predicate": f"amount BETWEEN {LO:g} AND {HI:g}",
            "data_seed": SEED, "distribution": "amount ~ U(1,1000) fp64 (Philox4x32-10)",
            "l2": "inputs larger than L2: 8 bytes x records_per_gpu per pass vs 126 MB L2, no flush needed",
            "parallelism": (f"range-shard x{args.gpus}; 64-byte shard partials exchanged " +
                            ("by an NCCL all-gather" if getattr(args, "no_fused", False) else "inside the scan kernel (NVLink peer stores, CUDA IPC mailboxes)"))
                           if args.gpus > 1 else "one shard, nothing to exchange: the kernel stores its 64-byte result straight into pinned host memory"}


def expected_answers(total: int):
    """Exact COUNT / SUM of the benchmark query over the first `total` rows (tests/golden/bench_expected.json, minted by
    tests/golden/make_bench_expected.py from the oracle's generator and long-double sums)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_expected.json")) as f:
            t = json.load(f)["tables"].get(str(total))
        return (t["count_where"], float.fromhex(t["sum_where"])) if t else None
    except Exception:
        return None


def static_profile(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def h2d_peak_gbs(torch, stream, gib=1):
    """Pinned host -> device copy bandwidth of this rank's GPU (plain copies, 1 GiB, best of 3): the e2e path's own roofline."""
    host = torch.empty(gib << 30, dtype=torch.uint8).pin_memory()
    dev = torch.empty(gib << 30, dtype=torch.uint8, device="cuda")
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); dev.copy_(host, non_blocking=True); e1.record(stream); torch.cuda.synchronize()
        best = max(best, (gib << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del host, dev
    return best


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto = strong for N > 1 (ONE 1 B-record table split over the GPUs, BASELINE configs[2]); weak = 1 B records per GPU")
    ap.add_argument("--records", type=int, default=int(os.environ.get("AQE_BENCH_RECORDS", 1_000_000_000)),
                    help="records in total (strong) or per GPU (weak)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-fused", action="store_true", help="merge shards with an NCCL all-gather instead of the in-kernel NVLink exchange")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained back-to-back loop reported next to the headline")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-approx", action="store_true")
    ap.add_argument("--skip-sql", action="store_true")
    ap.add_argument("--skip-single-process", action="store_true")
    ap.add_argument("--skip-weak", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import approximatequeryengine_b200 as aqe
    from approximatequeryengine_b200 import sharded

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched under torchrun (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line (rank 0): libraries that printf to fd 1 (NCCL's version banner does) are sent
    # to stderr for the life of the process, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = aqe.lib()
    first, rows = shard_rows(args, rank)
    total = total_rows(args)
    eng = aqe.Engine(local).generate(rows, seed=SEED, first_row=first, columns=("amount",))
    # a dedicated stream: aqe_scan_async(stream=0) means "the handle's own stream", so never hand it the
    # default stream's 0 handle -- kernels, collectives, copies and the timing events all go on `stream`
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    partial = torch.zeros(8, dtype=torch.int64, device="cuda")               # one 64-byte aqe_partial
    gathered = torch.zeros(8 * world, dtype=torch.int64, device="cuda")
    host_out = torch.zeros(8 * world, dtype=torch.int64).pin_memory()        # pinned + UVA: kernels can store into it
    host_nccl = torch.zeros(8 * world, dtype=torch.int64).pin_memory()
    fused = world > 1 and not args.no_fused
    fused_note = None
    if fused:
        # one process per GPU on one NVSwitch box: map every rank's mailbox (CUDA IPC) so that the scan kernel's last
        # block publishes the shard partial to all peers over NVLink and folds all of them itself.  If peer mapping is
        # not possible in this environment every rank switches (together) to the NCCL all-gather path -- still the GPU path.
        ok = 1
        try:
            mine = eng.exchange_init(rank, world)
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            eng.exchange_connect(handles)
            eng.exchange_set_total_rows(total)
        except Exception as e:  # noqa: BLE001
            ok, fused_note = 0, f"CUDA IPC peer mapping unavailable ({e}); NCCL all-gather used"
        okt = torch.tensor([ok], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:
            fused = False
            fused_note = fused_note or "CUDA IPC peer mapping unavailable on another rank; NCCL all-gather used"
            args.no_fused = True
        dist.barrier()

    def step_nccl(out=host_nccl):
        eng.scan_async(partial.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream)
        dist.all_gather_into_tensor(gathered, partial)
        out.copy_(gathered, non_blocking=True)

    if fused:
        def step():       # scan + exchange + rank-order merge in ONE kernel; the table-level result lands in pinned memory
            eng.scan_exchange_async(host_out.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream)
    elif world > 1:
        def step():
            step_nccl(host_out)
    else:
        def step():       # single shard: the kernel stores its 64-byte result straight into pinned host memory
            eng.scan_async(host_out.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step()
    scan_kernel_name = L.aqe_last_scan_kernel().decode()      # what the step's launch path actually launched
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = L.aqe_launch_count()
    total_ms = timed(step, args.steps)
    launches = L.aqe_launch_count() - launches0
    # the scan kernel alone, same stream, for the roofline
    kern_ms = timed(lambda: eng.scan_async(partial.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream), args.steps)

    # ---- parity of the timed result: exact COUNT, SUM to 1e-12 of the exactly rounded sum (tests/golden/bench_expected.json) ----
    sync_all()
    if fused:
        eng.exchange_check()
        merged = aqe.Partial.from_buffer_copy(host_out[:8].numpy().tobytes())     # already the table-level partial
        local_part = eng.scan("amount", "amount", LO, HI)
    else:
        parts = [aqe.Partial.from_buffer_copy(host_out[8 * r:8 * r + 8].numpy().tobytes()) for r in range(world)]
        merged = sharded.merge_partials(parts)
        local_part = parts[rank]
    want = expected_answers(total)
    parity = {"count": merged.count, "sum": merged.sum}
    if want:
        parity.update({"expected_count": want[0], "expected_sum": want[1], "count_exact": merged.count == want[0],
                       "sum_rel_err": abs(merged.sum - want[1]) / want[1], "sum_tolerance": 1e-12,
                       "source": "tests/golden/bench_expected.json (oracle generator + long-double Neumaier sum over the same rows)"})
        assert merged.count == want[0], ("COUNT differs from the oracle", merged.count, want[0])
        assert abs(merged.sum - want[1]) <= 1e-12 * want[1], ("SUM differs from the oracle", merged.sum, want[1])
    else:
        sel = merged.count / total
        parity.update({"expected_count": None, "note": "no minted answer for this table size: plausibility window only"})
        assert 0.39 < sel < 0.41 and 290.0 < merged.sum / merged.count < 310.0, (merged.count, merged.sum)
    if world > 1 and fused:
        # once, untimed: the fused in-kernel exchange and the NCCL all-gather + host merge must leave the same bytes
        step_nccl(); sync_all()
        via_nccl = sharded.merge_partials([aqe.Partial.from_buffer_copy(host_nccl[8 * r:8 * r + 8].numpy().tobytes()) for r in range(world)])
        same = (via_nccl.count, via_nccl.sum, via_nccl.comp) == (merged.count, merged.sum, merged.comp)
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity["fused_equals_nccl_bytes_on_all_ranks"] = bool(int(flag.item()))
        assert parity["fused_equals_nccl_bytes_on_all_ranks"], "fused exchange and NCCL path disagree"

    lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(lt)
    ms_per_step = total_ms / args.steps
    value = total / (ms_per_step * 1e-3)
    peak, peak_src = measured_peak()
    # When the step IS one launch of the scan kernel (single shard, or the fused-exchange kernel) the timed region's own
    # events give the kernel's average launch duration; otherwise the kernel-only loop on the same stream does.
    step_is_one_kernel = world == 1 or fused
    kms = ms_per_step if step_is_one_kernel else kern_ms / args.steps
    achieved = 8.0 * rows / (kms * 1e-3) / 1e9

    # ---- sustained: the same step back to back for >= 2 s (the 50-step region is ~55 ms: the device never warms up in it) ----
    sustained = None
    if args.sustained_seconds > 0:
        k = max(args.steps, int(args.sustained_seconds * 1e3 / ms_per_step) + 1)
        sms = timed(step, k)
        sustained = {"seconds": sms / 1e3, "steps": k, "ms_per_step": sms / k, "records_per_s": total / (sms / k * 1e-3),
                     "GBps_per_gpu": 8.0 * rows / (sms / k * 1e-3) / 1e9, "frac_of_measured_hbm_peak": 8.0 * rows / (sms / k * 1e-3) / 1e9 / peak}

    # ---- e2e: host buffers in, scalar out, through the C-ABI ------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 10)
        all_cpus = os.sched_getaffinity(0)
        numa_note = bind_to_gpu_numa_node(local)   # the pinned column is allocated (and first touched) next to this rank's GPU
        h2d_alone = h2d_peak_gbs(torch, stream)                # this GPU's link with the others idle
        if world > 1:
            dist.barrier()
        h2d_together = h2d_peak_gbs(torch, stream)             # every rank copying at once: what the host side sustains
        hptr = C.c_void_p()
        aqe.check(L.aqe_host_alloc(rows * 8, C.byref(hptr)))
        eng.read_column("amount", out_ptr=hptr.value)          # device -> pinned host (setup, untimed)
        def e2e_step():
            return aqe.host_scan_column(None, LO, HI, use_pred=True, device=local, ptr=hptr.value, n=rows, kind=0)
        for _ in range(3):
            p = e2e_step()
        assert p.count == local_part.count and abs(p.sum - local_part.sum) <= 1e-12 * abs(p.sum)
        l0 = L.aqe_launch_count()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        agg = torch.tensor([h2d_alone, h2d_together], dtype=torch.float64, device="cuda")
        slowest = torch.tensor([h2d_together], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            dist.all_reduce(slowest, op=dist.ReduceOp.MIN)
        chunk_mb = int(os.environ.get("AQE_E2E_CHUNK_MB", 64))
        nchunks = -(-rows * 8 // (chunk_mb << 20))
        e2e_val = total / (float(dt.item()) / e2e_steps)
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": rows * 8 * world,
               "d2h_bytes_per_step": 64 * nchunks * world, "steps": e2e_steps, "ms_per_step": float(dt.item()) / e2e_steps * 1e3,
               "api": "aqe_scan_host_column (C-ABI): pinned host amount column -> chunked H2D overlapped with k_scan -> merged partial",
               "launches_per_step": (L.aqe_launch_count() - l0) // e2e_steps, "host_numa": numa_note,
               "h2d_GBps": e2e_val * 8 / 1e9,
               "h2d_peak": {"measured": "plain pinned cudaMemcpyAsync, 1 GiB, best of 4, CUDA events", "sum_over_gpus_each_alone_GBps": float(agg[0].item()),
                            "sum_over_gpus_all_at_once_GBps": float(agg[1].item())},
               "frac_of_h2d_peak": e2e_val * 8 / 1e9 / float(agg[1].item()),
               "frac_of_h2d_peak_note": "against every GPU copying at once (what the host's memory system and PCIe root ports sustain together); "
                                        "sum_over_gpus_each_alone is N x one link",
               # equal shards finish when the GPU with the slowest link does: the step cannot beat N x that GPU's concurrent rate
               "slowest_gpu_all_at_once_GBps": float(slowest.item()), "frac_of_n_times_slowest_gpu": e2e_val * 8 / 1e9 / (world * float(slowest.item())),
               "host_limit": "the pod's cpuset confines host memory to one NUMA node (Mems_allowed 0; tools/microbench h2dnuma: mbind to other nodes is not "
                             "possible, interleave = default), so at 8 GPUs every link pulls from one socket: 236-261 GB/s in all, 24-35 GB/s per GPU "
                             "(profiles/r2_mb_h2dnuma_n8.jsonl)" if world >= 4 else None}
        L.aqe_host_free(hptr)
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline below uses every core

    clk = clocks.stop() if rank == 0 else None   # sampled over the timed region, the kernel-only loop, the sustained loop and the e2e phase

    # ---- secondary: CLT-terminated APPROX AVG at 1 % on 10 M records (configs[1]) through the drop-in module ----
    approx = None
    if rank == 0 and not args.skip_approx:
        b = aqe.backend()
        db = b.CustomBPlusDB(local)
        db.generate_synthetic(10_000_000, SEED)
        exact = db.sum_amount() / 10_000_000
        lat, kus, ns, errs = [], [], [], []
        for s in range(220):
            t0 = time.perf_counter(); r = db.approx_avg(error_percent=1.0, confidence_level=0.95, seed=s); t1 = time.perf_counter()
            if s >= 20:
                lat.append((t1 - t0) * 1e6); kus.append(r.kernel_us); ns.append(r.samples_used); errs.append(abs(r.value - exact) / exact * 100)
        ceil = static_profile("r2_gather_ceiling.json") or {}
        approx = {"workload": "BASELINE.json configs[1]: 10M records, APPROX AVG(amount), CLT early termination at 1% error, 95% confidence",
                  "latency_us_p50": statistics.median(lat), "latency_us_p99": sorted(lat)[int(len(lat) * 0.99) - 1], "kernel_us_p50": statistics.median(kus),
                  "samples_p50": statistics.median(ns), "abs_error_percent_p50": statistics.median(errs), "abs_error_percent_max": max(errs),
                  "api": "aqe_backend.CustomBPlusDB.approx_avg (pybind11 -> aqe_approx -> k_approx, one cooperative launch)",
                  "roofline": {"bound": "latency (one look of 16384 random 8-byte gathers + one grid barrier); DRAM lines touched / kernel time vs the measured random-gather ceiling",
                               "lines_per_launch": statistics.median(ns), "bytes_per_line": 128,
                               "achieved_Gsamples_per_s": statistics.median(ns) / (statistics.median(kus) * 1e-6) / 1e9,
                               "ceiling_Gsamples_per_s": ceil.get("random_Gsamples_per_s"), "ceiling_source": "static, profiles/r2_gather_ceiling.json (tools/microbench.cu)"}}
        # the sampler-plan path at the headline size: device moments over memory_stride / block plans of this rank's table
        gather = {}
        for method in ("memory_stride", "block"):
            pl = eng.plan(method, aqe.make_params(method, 1.0))
            eng.stats(pl)
            ts = []
            for _ in range(9):
                t0 = time.perf_counter(); st = eng.stats(pl); ts.append(time.perf_counter() - t0)
            ms = statistics.median(ts) * 1e3
            gather[method] = {"samples": st.n, "ms": ms, "Gsamples_per_s": st.n / (ms * 1e-3) / 1e9}
        g = gather["memory_stride"]
        gather["roofline"] = {"bound": "hbm (random 8-byte gathers: one 128-byte line of DRAM traffic per sample on B200, ncu: profiles/r2_mb_gather_ncu.csv)",
                              "kernel": "aqe::k_plan_stats, memory_stride 1% of this rank's table", "achieved_Gsamples_per_s": g["Gsamples_per_s"],
                              "achieved_line_GBps": g["Gsamples_per_s"] * 128, "ceiling_Gsamples_per_s": ceil.get("stride100_Gsamples_per_s"),
                              "frac": g["Gsamples_per_s"] / ceil["stride100_Gsamples_per_s"] if ceil.get("stride100_Gsamples_per_s") else None,
                              "ceiling_source": "static, profiles/r2_gather_ceiling.json (tools/microbench.cu, 16 loads in flight)", "includes": "host wall clock of the synchronous call (launch + sync)"}
        approx["gather"] = gather
        # the legacy return path the unmodified CLI uses (enhanced_aqe_cli.py:179-200): list[Record] by value, then a Python loop
        ts, ts_arr, ts_loop = [], [], []
        for _ in range(7):
            t0 = time.perf_counter(); recs = db.memory_stride_sample(1.0, 0); t1 = time.perf_counter()
            tot = sum(r.amount for r in recs); t2 = time.perf_counter()
            arr = db.sample_array("memory_stride", 1.0); t3 = time.perf_counter()
            ts.append(t1 - t0); ts_loop.append(t2 - t1); ts_arr.append(t3 - t2)
        approx["list_record_path"] = {"records": len(recs), "memory_stride_sample_ms": statistics.median(ts) * 1e3, "python_sum_loop_ms": statistics.median(ts_loop) * 1e3,
                                      "sample_array_ms": statistics.median(ts_arr) * 1e3, "ns_per_record_object": statistics.median(ts) / max(len(recs), 1) * 1e9,
                                      "note": "Record is a plain CPython type (csrc/aqe_pybind.cpp); round 1 (pybind11 class): 23.7 ms for the same 100 k records"}
        del db

    # ---- secondary, N > 1: BASELINE.json configs[3] -- block sampling, APPROX SUM at 0.5 % over the sharded table, one global
    # stop rule evaluated inside the persistent kernels (per-look moments exchanged through the NVLink mailboxes) ----
    approx_multi = None
    if fused and not args.skip_approx:
        res = {}
        truth = expected_answers(total)
        for design in ("block", "srs"):
            lat, kus, nrows, relerr = [], [], [], []
            for sd in range(60):
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter(); r = eng.approx("sum", error_percent=0.5, confidence_level=0.95, design=design, seed=sd, exchange=True); t1 = time.perf_counter()
                if sd >= 10:
                    lat.append((t1 - t0) * 1e6); kus.append(r.elapsed_us); nrows.append(r.n_samples); relerr.append(abs(r.estimate - 500.5 * total) / (500.5 * total) * 100)
            res[design] = {"latency_us_p50": statistics.median(lat), "kernel_us_p50": statistics.median(kus), "rows_read_p50": statistics.median(nrows),
                           "abs_error_percent_p50_vs_expectation": statistics.median(relerr)}
        # the legacy sampler families across the ranks: every rank walks the plan and gathers inside its window (ShardedTable.stats)
        tbl = sharded.ShardedTable(eng, total, first)
        samp = {}
        for method in ("parallel_block", "parallel_pointer", "memory_stride"):
            pl = tbl.plan(method, aqe.make_params(method, 1.0))
            ts = []
            for it in range(8):
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter(); st = tbl.stats(plan=pl); ts.append(time.perf_counter() - t0)
            samp[method] = {"samples": st.n, "ms_p50": statistics.median(ts) * 1e3, "estimate_rel_error_percent": abs(st.sum * (total / st.n) - 500.5 * total) / (500.5 * total) * 100}
        approx_multi = {"workload": f"BASELINE.json configs[3]: {total} records sharded over {world} GPUs, APPROX SUM(amount) at 0.5% error, 95% confidence, "
                                    "block (1000-row tiles) and SRS designs, global CLT stop rule", "api": "aqe_approx_exchange (C-ABI), all ranks", **res,
                        "sampler_plans_1pct": {"api": "sharded.ShardedTable.stats: aqe_stats_window per rank + all-gather of 64-byte partials + aqe_stats_merge", **samp}}

    # ---- secondary, N > 1: the whole box behind ONE CustomBPlusDB of ONE process (rank 0; the other ranks wait on the CPU) ----
    single = None
    if world > 1 and not args.skip_single_process:
        try:
            store = dist.distributed_c10d._get_default_store()
        except Exception:  # noqa: BLE001
            store = None
        sync_all()
        if store is None:
            single = {"skipped": "no c10d store to park the other ranks on"}
        elif rank == 0:
            try:
                b = aqe.backend()
                sp = b.CustomBPlusDB(list(range(world)))
                sp.generate_synthetic(total, SEED, 0, 0, 1 << 1)                   # amount only
                v = sp.sum_amount_where(LO, HI)
                for _ in range(5):
                    sp.sum_amount_where(LO, HI)
                ts = []
                for _ in range(200):
                    t0 = time.perf_counter(); v2 = sp.sum_amount_where(LO, HI); ts.append(time.perf_counter() - t0)
                ms = statistics.median(ts) * 1e3
                single = {"api": "aqe_backend.CustomBPlusDB(devices).sum_amount_where(100, 500): one process, one host thread per GPU, shard partials exchanged inside the kernels (peer-mapped mailboxes)",
                          "gpus": sp.shard_count, "fused": sp.shards_fused, "records": total, "ms_per_query_p50": ms, "ms_per_query_min": min(ts) * 1e3,
                          "records_per_s": total / (ms * 1e-3), "sum": v, "bits_equal_to_process_per_gpu_result": v == merged.sum and v2 == v,
                          "timing": "host wall clock of the synchronous call, the other ranks' processes idle"}
                lat = []
                for sd in range(30):
                    t0 = time.perf_counter(); r = sp.approx_sum(error_percent=0.5, seed=sd, design="block"); lat.append((time.perf_counter() - t0) * 1e6)
                single["approx_sum_0.5pct_block"] = {"status": str(r.status), "latency_us_p50": statistics.median(lat[5:]), "kernel_us": r.kernel_us,
                                                     "rel_err_percent": abs(r.value - 500.5 * total) / (500.5 * total) * 100}
                st = sp.sample_array("parallel_block", 1.0, stats=True)
                ts2 = []
                for _ in range(8):
                    t0 = time.perf_counter(); st = sp.sample_array("parallel_block", 1.0, stats=True); ts2.append(time.perf_counter() - t0)
                single["parallel_block_1pct_stats"] = {"samples": st["n"], "ms_p50": statistics.median(ts2) * 1e3,
                                                       "estimate_rel_error_percent": abs(st["sum"] * (total / st["n"]) - 500.5 * total) / (500.5 * total) * 100}
                del sp
                # the e2e step of ONE process: the table's whole amount column in pinned host memory -> every GPU, chunks handed out from one
                # counter (aqe_scan_host_column_multi), so that a GPU behind a slower link takes fewer of them
                if not args.skip_e2e and scaling_of(args) == "strong":
                    hp = C.c_void_p()
                    aqe.check(L.aqe_host_alloc(total * 8, C.byref(hp)))
                    for d in range(world):
                        f, r = shard_rows(args, d)
                        ed = aqe.Engine(d).generate(r, seed=SEED, first_row=f, columns=("amount",))
                        ed.read_column("amount", out_ptr=hp.value + f * 8)        # device -> pinned host (setup, untimed)
                        ed.close()
                    devs = list(range(world))
                    for _ in range(3):
                        pm = aqe.host_scan_column(None, LO, HI, use_pred=True, device=devs, ptr=hp.value, n=total, kind=0)
                    ok = pm.count == merged.count and abs(pm.sum - merged.sum) <= 1e-12 * abs(merged.sum)
                    k_e2e = args.e2e_steps or min(args.steps, 10)
                    l0 = L.aqe_launch_count()
                    t0 = time.perf_counter()
                    for _ in range(k_e2e):
                        aqe.host_scan_column(None, LO, HI, use_pred=True, device=devs, ptr=hp.value, n=total, kind=0)
                    dt1 = (time.perf_counter() - t0) / k_e2e
                    chunk_mb = int(os.environ.get("AQE_E2E_CHUNK_MB", 64))
                    single["e2e"] = {"value": total / dt1, "unit": UNIT, "ms_per_step": dt1 * 1e3, "steps": k_e2e, "h2d_bytes_per_step": total * 8,
                                     "d2h_bytes_per_step": 64 * -(-total * 8 // (chunk_mb << 20)), "h2d_GBps": total * 8 / dt1 / 1e9,
                                     "launches_per_step": (L.aqe_launch_count() - l0) // k_e2e, "count_and_sum_equal_the_table_level_result": ok,
                                     "api": "aqe_scan_host_column_multi (C-ABI): ONE pinned host amount column of the whole table -> chunks of "
                                            f"{chunk_mb} MiB handed out to the {world} GPUs from one counter, H2D overlapped with k_scan -> partials merged in chunk order",
                                     "timing": "host wall clock of the synchronous calls, the other ranks' processes idle"}
                    assert ok, ("single-process e2e differs from the table-level result", pm.count, pm.sum, merged.count, merged.sum)
                    L.aqe_host_free(hp)
            except Exception as ex:  # noqa: BLE001
                single = {**(single or {}), "error": repr(ex)}
            store.set("aqe_single_process_done", "1")
        else:
            store.wait(["aqe_single_process_done"])      # on the CPU: no kernel of this rank runs while rank 0 uses every GPU
        sync_all()

    # ---- secondary, N > 1: the weak-scaling figure (1 B records per GPU), kept next to the strong headline ----
    weak = None
    if world > 1 and scaling_of(args) == "strong" and not args.skip_weak:
        del eng
        torch.cuda.empty_cache()
        engw = aqe.Engine(local).generate(args.records, seed=SEED, first_row=rank * args.records, columns=("amount",))
        if fused:
            handles = [None] * world
            dist.all_gather_object(handles, engw.exchange_init(rank, world))
            engw.exchange_connect(handles)
            engw.exchange_set_total_rows(args.records * world)
            dist.barrier()
            def wstep():
                engw.scan_exchange_async(host_out.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream)
        else:
            def wstep():
                engw.scan_async(partial.data_ptr(), "amount", "amount", LO, HI, stream=stream.cuda_stream)
                dist.all_gather_into_tensor(gathered, partial)
                host_out.copy_(gathered, non_blocking=True)
        for _ in range(args.warmup):
            wstep()
        wms = timed(wstep, args.steps) / args.steps
        sync_all()
        wmerged = aqe.Partial.from_buffer_copy(host_out[:8].numpy().tobytes()) if fused else sharded.merge_partials(
            [aqe.Partial.from_buffer_copy(host_out[8 * r:8 * r + 8].numpy().tobytes()) for r in range(world)])
        wwant = expected_answers(args.records * world)
        weak = {"scaling": "weak", "records_per_gpu": args.records, "total_records": args.records * world, "ms_per_step": wms,
                "records_per_s": args.records * world / (wms * 1e-3), "GBps_per_gpu": 8.0 * args.records / (wms * 1e-3) / 1e9,
                "count_exact": (wmerged.count == wwant[0]) if wwant else None, "sum_rel_err": (abs(wmerged.sum - wwant[1]) / wwant[1]) if wwant else None}
        del engw

    # ---- secondary: the SQL-string path (run_query*) -- grouped scans over the whole table (BASELINE configs[2]: 1 B rows) ----
    sql = None
    if rank == 0 and world == 1 and not args.skip_sql:
        n_sql = min(rows, 1_000_000_000)
        es = aqe.Engine(local).generate(n_sql, seed=SEED, columns=("id", "amount", "region", "product_id"))
        sql = {"rows": n_sql, "api": "aqe_sql_run (C-ABI, caller-owned row buffer) -> k_sql_ring, one launch per query; ms = median host wall clock of the synchronous call",
               "queries": []}
        buf = (aqe.SqlRow * aqe.SQL_MAX_GROUPS)()
        ngot = C.c_uint32()
        # what holds each of these kernels back, from the committed ncu captures (profiles/r2_sql_ncu_details.txt)
        for q, pct, mode, width, limiter in (
                ("SELECT SUM(amount) FROM sales GROUP BY region", 0, "value", 12, "hbm"),
                ("SELECT SUM(amount) FROM sales WHERE amount BETWEEN 100 AND 500 GROUP BY region", 0, "ci_reference", 12,
                 "instruction issue: ~43 warp instructions per row (56 before the branch-free row add; two f64 -> 62-bit conversions, 3-word packing of rows / "
                 "value / square, two shared-memory read-modify-writes) on 18 warps per SM; no pipe above 68 % (DESIGN 8)"),
                ("SELECT AVG(amount) FROM sales GROUP BY product_id", 0, "value", 12,
                 "shared-memory data pipe at 77 % of its wavefront peak: 3 ATOMS per row on 1000 CTA-shared bins, ~4.2 wavefronts each (bank conflicts of random keys)"),
                ("SELECT SUM(amount) FROM sales WHERE amount > 900 GROUP BY product_id", 0, "value", 12, "hbm + shared-memory atomics for the passing rows"),
                ("SELECT SUM(amount) FROM sales GROUP BY region", 10, "ci_reference", 12, None)):
            def call():
                aqe.check(es.L.aqe_sql_run(es.h, q.encode(), pct, aqe.SQL_MODE[mode], buf, aqe.SQL_MAX_GROUPS, C.byref(ngot)))
            call()
            ts = []
            for _ in range(9):
                t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
            ms = statistics.median(ts) * 1e3
            gbps = width * n_sql / 1e9 / (ms * 1e-3) if pct == 0 else None
            sql["queries"].append({"sql": q, "sample_percent": pct, "mode": mode, "ms": ms, "records_per_s": n_sql / (ms * 1e-3), "groups": int(ngot.value),
                                   "algorithmic_GBps_full_scan": gbps, "frac_of_measured_hbm_peak": gbps / measured_peak()[0] if gbps else None,
                                   "limiter": limiter})
        del es

    # ---- secondary: the drop-in call sequence of enhanced_aqe_cli.py:327-346 end to end -- open_database(file) once, then queries ----
    dropin = None
    if rank == 0 and world == 1 and not args.skip_e2e:
        import tempfile
        n_file = min(rows, 100_000_000)
        tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        with tempfile.TemporaryDirectory(dir=tmpdir) as td:
            path = os.path.join(td, "sales.aqe")
            aqe.Engine(local).generate(n_file, seed=SEED).save_file(path)
            b = aqe.backend()
            db = b.CustomBPlusDB(local)
            t0 = time.perf_counter(); ok = db.open_database(path); t_open = time.perf_counter() - t0
            assert ok and db.get_total_records() == n_file
            v = db.sum_amount_where(LO, HI)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter(); db.sum_amount_where(LO, HI); ts.append(time.perf_counter() - t0)
            ms = statistics.median(ts) * 1e3
            dropin = {"api": "aqe_backend: CustomBPlusDB().open_database(file) once, then sum_amount_where(100, 500) per query (enhanced_aqe_cli.py:327-346)",
                      "records": n_file, "file_bytes": 24 + 32 * n_file, "medium": tmpdir or "tmp", "open_database_s": t_open, "ingest_GBps": 32 * n_file / t_open / 1e9,
                      "query_ms_p50": ms, "query_records_per_s": n_file / (ms * 1e-3), "first_query_records_per_s_including_open": n_file / (t_open + ms * 1e-3), "sum": v}
            del db

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_reference(steps=10, warmup=2)

    if rank == 0:
        # N > 1: the e2e step exists in two deployments, both measured above on the same table -- one process per GPU (equal shards, each rank
        # its own pinned slice) and ONE process feeding every GPU from one pinned column with chunks handed out dynamically.  The line's `e2e`
        # is the faster one (named in `api`); the other stays beside it.
        sp_e2e = (single or {}).get("e2e") if isinstance(single, dict) else None
        if e2e and sp_e2e and sp_e2e.get("count_and_sum_equal_the_table_level_result") and sp_e2e["value"] > e2e["value"]:
            per_rank = {k: e2e[k] for k in ("value", "ms_per_step", "h2d_GBps", "frac_of_h2d_peak", "frac_of_n_times_slowest_gpu", "api", "launches_per_step", "d2h_bytes_per_step")}
            e2e = {**e2e, **{k: sp_e2e[k] for k in ("value", "ms_per_step", "h2d_GBps", "api", "launches_per_step", "d2h_bytes_per_step", "steps", "timing")},
                   "frac_of_h2d_peak": sp_e2e["h2d_GBps"] / e2e["h2d_peak"]["sum_over_gpus_all_at_once_GBps"],
                   "frac_of_n_times_slowest_gpu": None, "deployment": "one process, every GPU of the box (rank 0; the other ranks idle)",
                   "process_per_gpu": per_rank}
        traffic = ncu_traffic_per_record()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, rows),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (traffic * rows) if traffic else None,
                             "traffic_source": "static: DRAM bytes per record of this kernel from the committed ncu --set full capture (profiles/roofline_traffic.json) x records per launch",
                             "kernel": scan_kernel_name, "kernel_source": "reported by the library for the launch the timed step made (aqe_last_scan_kernel)", "kernel_ms": kms,
                             "kernel_ms_source": "timed region (a step is exactly one launch of this kernel)" if step_is_one_kernel else "kernel-only loop, same stream",
                             "kernel_ms_isolated_loop": kern_ms / args.steps, "algorithmic_bytes_per_launch": 8 * rows, "peak_source": peak_src},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_dropin": dropin, "gpu_launches": int(lt.item()), "clocks": clk, "sustained": sustained, "parity_check": parity,
                "approx": approx, "approx_multi_gpu": approx_multi, "single_process": single, "weak_scaling": weak, "sql": sql,
                "exchange": ("fused in-kernel NVLink mailbox" if fused else ("nccl all_gather" if world > 1 else None)), "exchange_note": fused_note,
                "result": {"count": merged.count, "sum": merged.sum}}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
