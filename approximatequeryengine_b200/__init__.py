"""approximatequeryengine_b200 -- B200-native aggregation engine behind ApproximateQueryEngine's
``aqe_backend`` API.

Two native artefacts live in ``_lib/`` (built in-tree by ``python -m approximatequeryengine_b200.build``):

* ``libaqe_b200.so``  -- the C-ABI engine (include/aqe_b200.h): hand-written sm_100a kernels + host code;
* ``aqe_backend*.so`` -- the pybind11 module with the reference's Python surface
  (reference: src/aqe_backend/bindings/bindings.cpp).

This package is the host-side mirror: ``backend()`` returns the drop-in module, ``Engine`` is a ctypes
view of the C-ABI for tests/bench, ``sharded`` merges per-rank partials (one process per GPU).
There is no CPU fallback: without the built libraries or without a CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import sys

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libaqe_b200.so")

RECORD_DTYPE = np.dtype(
    [("id", "<i8"), ("amount", "<f8"), ("region", "<i4"), ("product_id", "<i4"), ("timestamp", "<i8")]
)

COLS = {"id": 0, "amount": 1, "region": 2, "product_id": 3, "timestamp": 4, None: -1}
COL_KIND = {"amount": 0, "id": 1, "timestamp": 1, "region": 2, "product_id": 2}  # 0 f64, 1 i64, 2 i32
AGG = {"sum": 0, "avg": 1, "count": 2}
DESIGN = {"srs": 0, "block": 1}
METHODS = {
    "slow_pointer": 0, "fast_pointer": 1, "dual_pointer": 2, "parallel_pointer": 3, "random_pointer": 4,
    "memory_stride": 5, "optimized_address_arithmetic": 6, "index_based": 7, "byte_offset": 8,
    "optimized_clt": 9, "block": 10, "page": 11, "parallel_block": 12, "node_skip": 13, "balanced_tree": 14,
    "direct_access": 15, "adaptive_block": 16, "stratified_block": 17, "sample_records": 18,
    "optimized_sequential": 19, "random_start_nth": 20, "address_arithmetic": 21,
    "random_start_memory_stride": 22, "multithreaded_memory_stride": 23, "clt_validated_dual_pointer": 24,
    "signal_based_clt": 25,
}
STATUS = {0: "STABLE", 1: "DRIFTING", 2: "INSUFFICIENT_DATA", 3: "ERROR"}
CI_MODE = {"default": 0, "plain": 1, "stein": 2, "stein_guarded": 3}


class AqeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[aqe status {code}] {msg}")
        self.code = code


class SampleParams(C.Structure):
    _fields_ = [
        ("sample_percent", C.c_double), ("step_size", C.c_int64), ("num_threads", C.c_int64),
        ("block_size", C.c_int64), ("block_size_max", C.c_int64), ("check_interval", C.c_int64),
        ("confidence_level", C.c_double), ("max_error_percent", C.c_double), ("seed", C.c_uint64),
    ]


class ScanSpec(C.Structure):
    _fields_ = [("agg_col", C.c_int32), ("pred_col", C.c_int32), ("lo", C.c_double), ("hi", C.c_double)]


class Partial(C.Structure):
    _fields_ = [
        ("count", C.c_uint64), ("sum", C.c_double), ("comp", C.c_double), ("isum_lo", C.c_uint64),
        ("isum_hi", C.c_int64), ("sumsq", C.c_double), ("minv", C.c_double), ("maxv", C.c_double),
    ]

    @property
    def isum(self) -> int:
        return (int(self.isum_hi) << 64) + int(self.isum_lo)


class Stats(C.Structure):
    _fields_ = [("n", C.c_uint64), ("mean", C.c_double), ("m2", C.c_double), ("sum", C.c_double)]


class StatsPartial(C.Structure):
    """One shard's mergeable sums over a sample plan (aqe_stats_partial, 64 bytes)."""
    _fields_ = [("n", C.c_uint64), ("sum", C.c_double), ("sum_c", C.c_double), ("shift", C.c_double), ("sd", C.c_double),
                ("sd_c", C.c_double), ("sdd", C.c_double), ("sdd_c", C.c_double)]


class Segment(C.Structure):
    _fields_ = [("base", C.c_int64), ("outer_step", C.c_int64), ("inner_len", C.c_int64), ("count", C.c_int64),
                ("scale", C.c_double), ("kind", C.c_int32), ("_pad", C.c_int32)]


class ApproxSpec(C.Structure):
    _fields_ = [
        ("agg", C.c_int32), ("design", C.c_int32), ("agg_col", C.c_int32), ("pred_col", C.c_int32),
        ("lo", C.c_double), ("hi", C.c_double), ("error_percent", C.c_double), ("confidence_level", C.c_double),
        ("seed", C.c_uint64), ("min_samples", C.c_uint64), ("max_samples", C.c_uint64),
        ("block_size", C.c_uint32), ("ci_mode", C.c_uint32),
    ]


class ApproxResult(C.Structure):
    _fields_ = [
        ("estimate", C.c_double), ("ci_lower", C.c_double), ("ci_upper", C.c_double), ("error_margin", C.c_double),
        ("confidence_level", C.c_double), ("n_samples", C.c_uint64), ("n_units", C.c_uint64),
        ("population", C.c_uint64), ("mean", C.c_double), ("m2", C.c_double), ("rounds", C.c_uint32),
        ("status", C.c_int32), ("elapsed_us", C.c_double), ("pass_fraction", C.c_double),
    ]


class SqlTerm(C.Structure):
    _fields_ = [("col", C.c_int32), ("has_ne", C.c_int32), ("lo", C.c_double), ("hi", C.c_double),
                ("ilo", C.c_int64), ("ihi", C.c_int64), ("ne", C.c_double), ("ine", C.c_int64)]


class SqlQuery(C.Structure):
    """Parsed + compiled query: the reference's ``struct Query`` (parser.h:17-24) with names resolved."""
    _fields_ = [("agg", C.c_int32), ("agg_col", C.c_int32), ("group_col", C.c_int32), ("sample_percent", C.c_int32),
                ("n_alt", C.c_int32), ("always_false", C.c_int32), ("top_level_or", C.c_int32), ("_pad", C.c_int32), ("n_terms", C.c_int32 * 8), ("terms", (SqlTerm * 5) * 8),
                ("agg_text", C.c_char * 32), ("column", C.c_char * 64), ("table", C.c_char * 64),
                ("group_by", C.c_char * 64), ("where", C.c_char * 512)]

    def branches(self):
        """The WHERE clause as a list of conjunctions (OR-ed), each a list of SqlTerm."""
        return [[self.terms[a][t] for t in range(self.n_terms[a])] for a in range(self.n_alt)]


class SqlRow(C.Structure):
    _fields_ = [("key", C.c_int64), ("value", C.c_double), ("ci_lower", C.c_double), ("ci_upper", C.c_double),
                ("count", C.c_uint64), ("sum", C.c_double), ("sumsq", C.c_double), ("isum_lo", C.c_uint64),
                ("isum_hi", C.c_int64), ("is_null", C.c_int32), ("_pad", C.c_int32)]

    @property
    def isum(self) -> int:
        return (int(self.isum_hi) << 64) + int(self.isum_lo)


class SqlFacts(C.Structure):
    _fields_ = [("key_min", C.c_int64), ("key_max", C.c_int64), ("agg_absmax", C.c_double),
                ("agg_is_integer", C.c_int32), ("_pad", C.c_int32)]


class SqlLayout(C.Structure):
    _fields_ = [("key_min", C.c_int64), ("n_groups", C.c_uint32), ("sum_shift", C.c_int32), ("sq_shift", C.c_int32),
                ("is_integer", C.c_int32)]


SQL_MODE = {"value": 0, "ci_reference": 1, "ci_correct": 2}
SQL_MOMENTS, SQL_UNSAMPLED = 1, 2
SQL_MAX_GROUPS = 4096
SQL_MAX_ALT = 8

_lib = None


def lib() -> C.CDLL:
    """The C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m approximatequeryengine_b200.build` (needs nvcc)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
    sig = {
        "aqe_abi_version": (i32, []),
        "aqe_last_error": (C.c_char_p, []),
        "aqe_last_scan_kernel": (C.c_char_p, []),
        "aqe_device_count": (i32, [C.POINTER(C.c_int)]),
        "aqe_launch_count": (u64, []),
        "aqe_host_alloc": (i32, [C.c_size_t, C.POINTER(vp)]),
        "aqe_host_free": (i32, [vp]),
        "aqe_create": (i32, [i32, C.POINTER(vp)]),
        "aqe_open": (i32, [C.c_char_p, i32, C.POINTER(vp)]),
        "aqe_load_file": (i32, [vp, C.c_char_p, u64, u64]),
        "aqe_save_file": (i32, [vp, C.c_char_p]),
        "aqe_append_records": (i32, [vp, vp, C.c_size_t]),
        "aqe_from_host_records": (i32, [vp, vp, C.c_size_t]),
        "aqe_attach_device_columns": (i32, [vp, vp, vp, vp, vp, vp, u64]),
        "aqe_generate_synthetic": (i32, [vp, u64, u64, u64, i32, C.c_uint32]),
        "aqe_synth_rows_host": (i32, [u64, u64, u64, i32, vp]),
        "aqe_close": (i32, [vp]),
        "aqe_create_sharded": (i32, [C.POINTER(C.c_int), i32, C.POINTER(vp)]),
        "aqe_open_sharded": (i32, [C.c_char_p, i32, C.POINTER(vp)]),
        "aqe_shard_count": (i32, [vp]),
        "aqe_shard": (vp, [vp, i32]),
        "aqe_shard_first_row": (u64, [vp, i32]),
        "aqe_shards_fused": (i32, [vp]),
        "aqe_reference_order": (i32, [vp, u64, vp, vp, C.c_size_t, vp]),
        "aqe_count": (u64, [vp]),
        "aqe_node_count": (u64, [vp]),
        "aqe_tree_height": (u64, [vp]),
        "aqe_device": (i32, [vp]),
        "aqe_column_device_ptr": (vp, [vp, i32]),
        "aqe_read_records": (i32, [vp, u64, u64, vp]),
        "aqe_read_column": (i32, [vp, i32, u64, u64, vp]),
        "aqe_scan": (i32, [vp, C.POINTER(ScanSpec), C.POINTER(Partial)]),
        "aqe_scan_async": (i32, [vp, C.POINTER(ScanSpec), vp, vp]),
        "aqe_scan_host_column": (i32, [i32, vp, i32, u64, dbl, dbl, i32, C.POINTER(Partial)]),
        "aqe_scan_host_column_multi": (i32, [C.POINTER(C.c_int), i32, vp, i32, u64, dbl, dbl, i32, C.POINTER(Partial)]),
        "aqe_merge_partials": (i32, [C.POINTER(Partial), i32, i32, C.POINTER(Partial)]),
        "aqe_exchange_init": (i32, [vp, i32, i32, vp]),
        "aqe_exchange_connect": (i32, [vp, vp]),
        "aqe_exchange_check": (i32, [vp]),
        "aqe_scan_exchange": (i32, [vp, C.POINTER(ScanSpec), C.POINTER(Partial)]),
        "aqe_scan_exchange_async": (i32, [vp, C.POINTER(ScanSpec), vp, vp]),
        "aqe_sum_f64": (i32, [vp, i32, C.POINTER(dbl)]),
        "aqe_sum_where_f64": (i32, [vp, i32, dbl, dbl, C.POINTER(dbl), C.POINTER(u64)]),
        "aqe_sum_i128": (i32, [vp, i32, C.POINTER(u64), C.POINTER(C.c_int64)]),
        "aqe_sample_params_default": (None, [C.POINTER(SampleParams), i32]),
        "aqe_plan_build": (i32, [vp, u64, i32, C.POINTER(SampleParams), C.POINTER(vp)]),
        "aqe_plan_from_indices": (i32, [vp, u64, C.POINTER(vp)]),
        "aqe_plan_count": (u64, [vp]),
        "aqe_plan_num_segments": (C.c_uint32, [vp]),
        "aqe_plan_segments": (i32, [vp, C.POINTER(Segment), C.c_uint32]),
        "aqe_plan_indices": (i32, [vp, vp, u64]),
        "aqe_plan_sorted_by_amount": (i32, [vp]),
        "aqe_plan_free": (None, [vp]),
        "aqe_stats_from_plan": (i32, [vp, vp, i32, C.POINTER(Stats)]),
        "aqe_stats_from_plan_where": (i32, [vp, vp, i32, i32, dbl, dbl, C.POINTER(Stats)]),
        "aqe_stats_from_indices": (i32, [vp, vp, u64, i32, C.POINTER(Stats)]),
        "aqe_stats_window": (i32, [vp, vp, i32, i32, dbl, dbl, u64, C.POINTER(StatsPartial)]),
        "aqe_stats_merge": (i32, [C.POINTER(StatsPartial), i32, C.POINTER(Stats)]),
        "aqe_gather_window": (i32, [vp, vp, u64, u64, u64, vp, C.POINTER(u64)]),
        "aqe_plan_table_rows": (u64, [vp]),
        "aqe_gather_plan": (i32, [vp, vp, vp, u64]),
        "aqe_gather_records": (i32, [vp, vp, u64, vp]),
        "aqe_fast_aggregated_sum": (i32, [vp, C.POINTER(SampleParams), C.POINTER(dbl), C.POINTER(u64)]),
        "aqe_estimate": (i32, [C.POINTER(Stats), u64, i32, dbl, i32, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]),
        "aqe_approx": (i32, [vp, C.POINTER(ApproxSpec), C.POINTER(ApproxResult)]),
        "aqe_exchange_set_total_rows": (i32, [vp, u64]),
        "aqe_approx_exchange": (i32, [vp, C.POINTER(ApproxSpec), C.POINTER(ApproxResult)]),
        "aqe_approx_merge": (i32, [C.POINTER(ApproxResult), i32, i32, dbl, C.POINTER(ApproxResult)]),
        "aqe_z_score": (dbl, [dbl, i32]),
        "aqe_sql_parse": (i32, [C.c_char_p, i32, C.POINTER(SqlQuery)]),
        "aqe_sql_execute": (i32, [vp, C.POINTER(SqlQuery), i32, C.POINTER(SqlRow), C.c_uint32, C.POINTER(C.c_uint32)]),
        "aqe_sql_run": (i32, [vp, C.c_char_p, i32, i32, C.POINTER(SqlRow), C.c_uint32, C.POINTER(C.c_uint32)]),
        "aqe_sql_facts_of": (i32, [vp, C.POINTER(SqlQuery), C.POINTER(SqlFacts)]),
        "aqe_sql_layout_of": (i32, [C.POINTER(SqlQuery), C.POINTER(SqlFacts), i32, C.POINTER(SqlLayout)]),
        "aqe_sql_scan": (i32, [vp, C.POINTER(SqlQuery), C.POINTER(SqlLayout), i32, vp]),
        "aqe_sql_scan_exchange": (i32, [vp, C.POINTER(SqlQuery), C.POINTER(SqlLayout), i32, vp]),
        "aqe_sql_merge": (i32, [vp, vp, C.c_uint32]),
        "aqe_sql_finish": (i32, [C.POINTER(SqlQuery), i32, C.POINTER(SqlLayout), vp, vp, C.POINTER(SqlRow), C.c_uint32, C.POINTER(C.c_uint32)]),
        "aqe_sql_shifts": (i32, [dbl, i32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def backend():
    """The drop-in ``aqe_backend`` module (same surface as the reference's pybind11 module)."""
    if LIB_DIR not in sys.path:
        sys.path.insert(0, LIB_DIR)
    try:
        return importlib.import_module("aqe_backend")
    except ImportError as e:  # loud, never a fallback
        raise ImportError(f"aqe_backend extension not built in {LIB_DIR}: run `python -m approximatequeryengine_b200.build`") from e


def check(rc: int) -> None:
    if rc != 0:
        raise AqeError(rc, lib().aqe_last_error().decode(errors="replace"))


def make_params(method: str, sample_percent: float, **kw) -> SampleParams:
    p = SampleParams()
    lib().aqe_sample_params_default(C.byref(p), METHODS[method])
    p.sample_percent = sample_percent
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Plan:
    """A sample position list (affine segments or explicit indices)."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            try:
                _lib.aqe_plan_free(self.h)
            except Exception:  # interpreter shutdown
                pass
            self.h = None

    @property
    def count(self) -> int:
        return lib().aqe_plan_count(self.h)

    @property
    def num_segments(self) -> int:
        return lib().aqe_plan_num_segments(self.h)

    @property
    def by_amount_order(self) -> bool:
        return bool(lib().aqe_plan_sorted_by_amount(self.h))

    def segments(self):
        n = self.num_segments
        arr = (Segment * max(n, 1))()
        check(lib().aqe_plan_segments(self.h, arr, n))
        return list(arr[:n])

    def indices(self) -> np.ndarray:
        out = np.empty(self.count, dtype=np.int64)
        check(lib().aqe_plan_indices(self.h, _ptr(out), len(out)))
        return out


def build_plan(n_rows: int, method: str, params: SampleParams, engine: "Engine | None" = None) -> Plan:
    h = C.c_void_p()
    check(lib().aqe_plan_build(engine.h if engine else None, n_rows, METHODS[method], C.byref(params), C.byref(h)))
    return Plan(h)


def host_scan_column(col: np.ndarray, lo: float = 0.0, hi: float = 0.0, use_pred: bool = False, device: "int | list[int]" = 0,
                     ptr: int | None = None, n: int | None = None, kind: int | None = None) -> Partial:
    """End-to-end form: a host-resident column in, a scalar partial out (chunked H2D overlapped with the scan)."""
    out = Partial()
    if ptr is None:
        kind = {np.dtype("float64"): 0, np.dtype("int64"): 1, np.dtype("int32"): 2}[col.dtype]
        ptr, n = col.ctypes.data, len(col)
    if isinstance(device, (list, tuple)):   # several GPUs of this process: chunks handed out from one counter (aqe_scan_host_column_multi)
        devs = (C.c_int * len(device))(*device)
        check(lib().aqe_scan_host_column_multi(devs, len(device), C.c_void_p(ptr), kind, n, lo, hi, int(use_pred), C.byref(out)))
        return out
    check(lib().aqe_scan_host_column(device, C.c_void_p(ptr), kind, n, lo, hi, int(use_pred), C.byref(out)))
    return out


def merge_partials(parts, is_integer: bool = False) -> Partial:
    arr = (Partial * len(parts))(*parts)
    out = Partial()
    check(lib().aqe_merge_partials(arr, len(parts), int(is_integer), C.byref(out)))
    return out


class Engine:
    """ctypes view of one ``aqe_db`` handle (one shard on one GPU)."""

    def __init__(self, device: int | None = None, devices=None):
        """One shard on one GPU (`device`), or -- `devices` = a list of device ids, or "all" -- the whole table range-sharded
        over several GPUs of this process (aqe_create_sharded): every method below then answers for the whole table."""
        self.L = lib()
        self.h = C.c_void_p()
        if devices is not None:
            if devices == "all":
                check(self.L.aqe_create_sharded(None, 0, C.byref(self.h)))
            else:
                arr = (C.c_int * len(devices))(*devices)
                check(self.L.aqe_create_sharded(arr, len(devices), C.byref(self.h)))
            return
        if device is None:
            device = int(os.environ.get("AQE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        check(self.L.aqe_create(device, C.byref(self.h)))

    @property
    def shard_count(self) -> int:
        return self.L.aqe_shard_count(self.h)

    @property
    def fused(self) -> bool:
        return bool(self.L.aqe_shards_fused(self.h))

    def shard_first_row(self, g: int) -> int:
        return self.L.aqe_shard_first_row(self.h, g)

    def close(self):
        if getattr(self, "h", None):
            self.L.aqe_close(self.h)
            self.h = None

    __del__ = close

    # ---- ingest ----
    def load_file(self, path: str, first_row: int = 0, n_rows: int | None = None):
        check(self.L.aqe_load_file(self.h, path.encode(), first_row, (1 << 64) - 1 if n_rows is None else n_rows))
        return self

    def save_file(self, path: str):
        check(self.L.aqe_save_file(self.h, path.encode()))

    def from_rows(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=RECORD_DTYPE)
        check(self.L.aqe_from_host_records(self.h, _ptr(rows), len(rows)))
        return self

    def append(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=RECORD_DTYPE)
        check(self.L.aqe_append_records(self.h, _ptr(rows), len(rows)))
        return self

    def generate(self, n_rows: int, seed: int = 7, first_row: int = 0, dist: int = 0, columns=("id", "amount", "region", "product_id", "timestamp")):
        mask = 0
        for c in columns:
            mask |= 1 << COLS[c]
        check(self.L.aqe_generate_synthetic(self.h, seed, first_row, n_rows, dist, mask))
        return self

    def attach(self, n: int, id=0, amount=0, region=0, product_id=0, timestamp=0):
        check(self.L.aqe_attach_device_columns(self.h, id or None, amount or None, region or None, product_id or None, timestamp or None, n))
        return self

    def read_rows(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.count - first if n is None else n
        out = np.empty(n, dtype=RECORD_DTYPE)
        check(self.L.aqe_read_records(self.h, first, n, _ptr(out)))
        return out

    def read_column(self, col: str, first: int = 0, n: int | None = None, out_ptr: int | None = None):
        """Column slice -> host.  With out_ptr the copy lands in caller memory (e.g. a pinned buffer)."""
        n = self.count - first if n is None else n
        if out_ptr is not None:
            check(self.L.aqe_read_column(self.h, COLS[col], first, n, C.c_void_p(out_ptr)))
            return None
        out = np.empty(n, dtype={0: np.float64, 1: np.int64, 2: np.int32}[COL_KIND[col]])
        check(self.L.aqe_read_column(self.h, COLS[col], first, n, _ptr(out)))
        return out

    @property
    def count(self) -> int:
        return self.L.aqe_count(self.h)

    def column_ptr(self, col: str) -> int:
        return self.L.aqe_column_device_ptr(self.h, COLS[col]) or 0

    # ---- exact ----
    def scan(self, agg_col="amount", pred_col=None, lo=0.0, hi=0.0) -> Partial:
        sp = ScanSpec(COLS[agg_col], COLS[pred_col], lo, hi)
        out = Partial()
        check(self.L.aqe_scan(self.h, C.byref(sp), C.byref(out)))
        return out

    def scan_async(self, partial_dev_ptr: int, agg_col="amount", pred_col=None, lo=0.0, hi=0.0, stream: int = 0):
        sp = ScanSpec(COLS[agg_col], COLS[pred_col], lo, hi)
        check(self.L.aqe_scan_async(self.h, C.byref(sp), C.c_void_p(partial_dev_ptr), C.c_void_p(stream)))

    # ---- fused cross-GPU exchange (see include/aqe_b200.h) ----
    def exchange_init(self, rank: int, world: int) -> bytes:
        buf = C.create_string_buffer(64)
        check(self.L.aqe_exchange_init(self.h, rank, world, buf))
        return buf.raw

    def exchange_connect(self, handles) -> None:
        blob = b"".join(handles)
        check(self.L.aqe_exchange_connect(self.h, blob))

    def scan_exchange(self, agg_col="amount", pred_col=None, lo=0.0, hi=0.0) -> Partial:
        sp = ScanSpec(COLS[agg_col], COLS[pred_col], lo, hi)
        out = Partial()
        check(self.L.aqe_scan_exchange(self.h, C.byref(sp), C.byref(out)))
        return out

    def scan_exchange_async(self, merged_dev_ptr: int, agg_col="amount", pred_col=None, lo=0.0, hi=0.0, stream: int = 0):
        sp = ScanSpec(COLS[agg_col], COLS[pred_col], lo, hi)
        check(self.L.aqe_scan_exchange_async(self.h, C.byref(sp), C.c_void_p(merged_dev_ptr), C.c_void_p(stream)))

    def exchange_check(self) -> None:
        check(self.L.aqe_exchange_check(self.h))

    def sum_amount(self) -> float:
        v = C.c_double()
        check(self.L.aqe_sum_f64(self.h, COLS["amount"], C.byref(v)))
        return v.value

    def sum_amount_where(self, lo: float, hi: float):
        v, c = C.c_double(), C.c_uint64()
        check(self.L.aqe_sum_where_f64(self.h, COLS["amount"], lo, hi, C.byref(v), C.byref(c)))
        return v.value, c.value

    def sum_int(self, col: str) -> int:
        lo, hi = C.c_uint64(), C.c_int64()
        check(self.L.aqe_sum_i128(self.h, COLS[col], C.byref(lo), C.byref(hi)))
        return (hi.value << 64) + lo.value

    # ---- sampled ----
    def plan(self, method: str, params: SampleParams) -> Plan:
        return build_plan(self.count, method, params, self)

    def plan_from_indices(self, idx) -> Plan:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        h = C.c_void_p()
        check(self.L.aqe_plan_from_indices(_ptr(idx), len(idx), C.byref(h)))
        return Plan(h)

    def stats(self, plan: Plan, col="amount", where=None, where_col="amount") -> Stats:
        s = Stats()
        if where is None:
            check(self.L.aqe_stats_from_plan(self.h, plan.h, COLS[col], C.byref(s)))
        else:
            check(self.L.aqe_stats_from_plan_where(self.h, plan.h, COLS[col], COLS[where_col], where[0], where[1], C.byref(s)))
        return s

    def stats_window(self, plan: Plan, window_first: int, col="amount", where=None, where_col="amount") -> StatsPartial:
        """This handle holds rows [window_first, window_first + count) of the plan's table: its mergeable sums."""
        out = StatsPartial()
        check(self.L.aqe_stats_window(self.h, plan.h, COLS[col], COLS[where_col] if where else -1, where[0] if where else 0.0,
                                      where[1] if where else 0.0, window_first, C.byref(out)))
        return out

    def gather_window(self, plan: Plan, window_first: int, k_first: int = 0, k_count: int | None = None):
        """Rows of plan positions [k_first, k_first + k_count) inside this handle's window (other slots zero) and their number."""
        k_count = plan.count - k_first if k_count is None else k_count
        out = np.zeros(k_count, dtype=RECORD_DTYPE)
        n = C.c_uint64()
        check(self.L.aqe_gather_window(self.h, plan.h, window_first, k_first, k_count, _ptr(out), C.byref(n)))
        return out, n.value

    def stats_from_indices(self, idx, col="amount") -> Stats:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        s = Stats()
        check(self.L.aqe_stats_from_indices(self.h, _ptr(idx), len(idx), COLS[col], C.byref(s)))
        return s

    def gather(self, plan: Plan) -> np.ndarray:
        out = np.empty(plan.count, dtype=RECORD_DTYPE)
        check(self.L.aqe_gather_plan(self.h, plan.h, _ptr(out), len(out)))
        return out

    def gather_indices(self, idx) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty(len(idx), dtype=RECORD_DTYPE)
        check(self.L.aqe_gather_records(self.h, _ptr(idx), len(idx), _ptr(out)))
        return out

    def fast_aggregated(self, params: SampleParams):
        s, n = C.c_double(), C.c_uint64()
        check(self.L.aqe_fast_aggregated_sum(self.h, C.byref(params), C.byref(s), C.byref(n)))
        return s.value, n.value

    def approx(self, agg="sum", error_percent=1.0, confidence_level=0.95, design="srs", seed=0, where=None,
               where_col="amount", agg_col="amount", block_size=0, min_samples=0, max_samples=0, exchange=False, ci_mode="default") -> ApproxResult:
        sp = ApproxSpec(AGG[agg], DESIGN[design], COLS[agg_col], COLS[where_col] if where else -1,
                        where[0] if where else 0.0, where[1] if where else 0.0, error_percent, confidence_level, seed,
                        min_samples, max_samples, block_size, CI_MODE[ci_mode])
        out = ApproxResult()
        check((self.L.aqe_approx_exchange if exchange else self.L.aqe_approx)(self.h, C.byref(sp), C.byref(out)))
        return out

    def exchange_set_total_rows(self, total_rows: int) -> None:
        check(self.L.aqe_exchange_set_total_rows(self.h, total_rows))

    # ---- SQL-string path (run_query*, bindings.cpp:126-136) ----
    def sql(self, query: str, sample_percent: int = 0, mode: str = "value"):
        """List of SqlRow (one without GROUP BY), ascending key."""
        rows = self.__dict__.get("_sql_rows")
        if rows is None:
            rows = self._sql_rows = (SqlRow * SQL_MAX_GROUPS)()   # reused: allocating 320 KiB per call dwarfs a small query
        n = C.c_uint32()
        check(self.L.aqe_sql_run(self.h, query.encode(), sample_percent, SQL_MODE[mode], rows, SQL_MAX_GROUPS, C.byref(n)))
        return list((SqlRow * n.value).from_buffer_copy(rows))   # one copy out of the reused buffer

    def sql_facts(self, q: SqlQuery) -> SqlFacts:
        f = SqlFacts()
        check(self.L.aqe_sql_facts_of(self.h, C.byref(q), C.byref(f)))
        return f

    def sql_scan(self, q: SqlQuery, layout: SqlLayout, flags: int = 0, exchange: bool = False) -> np.ndarray:
        """This shard's accumulators; with exchange=True (after exchange_connect) the table-level ones, merged inside the kernel."""
        acc = np.zeros(layout.n_groups * 5, dtype=np.uint64)
        check((self.L.aqe_sql_scan_exchange if exchange else self.L.aqe_sql_scan)(self.h, C.byref(q), C.byref(layout), flags, _ptr(acc)))
        return acc


def reference_order(ids, ops=None) -> np.ndarray:
    """Host only: the order in which the reference's B+ tree holds rows that arrived with these ids (aqe_reference_order; matters only
    where ids repeat).  ops: [(rows, kind)] -- the calls that inserted them, kind 0 = insert_batch / load / insert_record, 1 = rows already
    in table order; None = one insert_batch.  Returns perm with table[k] = arrival[perm[k]]."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    perm = np.empty(len(ids), dtype=np.uint64)
    ops = list(ops or [])
    op_rows = np.asarray([o[0] for o in ops], dtype=np.uint64)
    op_kinds = np.asarray([o[1] for o in ops], dtype=np.int32)
    check(lib().aqe_reference_order(_ptr(ids), len(ids), _ptr(op_rows) if ops else None, _ptr(op_kinds) if ops else None, len(ops), _ptr(perm)))
    return perm


def merge_stats(parts) -> Stats:
    """Fold the shards' sums of a sample plan in rank order (aqe_stats_merge, host code)."""
    arr = (StatsPartial * len(parts))(*parts)
    out = Stats()
    check(lib().aqe_stats_merge(arr, len(parts), C.byref(out)))
    return out


def estimate(stats: Stats, population: int, agg: str, z: float = 1.96, legacy_ci: bool = False):
    e, lo, hi = C.c_double(), C.c_double(), C.c_double()
    check(lib().aqe_estimate(C.byref(stats), population, AGG[agg], z, int(legacy_ci), C.byref(e), C.byref(lo), C.byref(hi)))
    return e.value, lo.value, hi.value


def sql_parse(query: str, sample_percent: int = 0) -> SqlQuery:
    """Host-only: the reference's parser (parser.cpp:20-75) + WHERE compilation."""
    q = SqlQuery()
    check(lib().aqe_sql_parse(query.encode(), sample_percent, C.byref(q)))
    return q


def sql_layout(q: SqlQuery, facts) -> SqlLayout:
    arr = (SqlFacts * len(facts))(*facts)
    out = SqlLayout()
    check(lib().aqe_sql_layout_of(C.byref(q), arr, len(facts), C.byref(out)))
    return out


def sql_merge(acc: np.ndarray, other: np.ndarray) -> np.ndarray:
    check(lib().aqe_sql_merge(_ptr(acc), _ptr(np.ascontiguousarray(other, dtype=np.uint64)), len(acc) // 5))
    return acc


def sql_finish(q: SqlQuery, layout: SqlLayout, acc: np.ndarray, mode: str = "value", exists: np.ndarray | None = None):
    rows = (SqlRow * max(layout.n_groups, 1))()
    n = C.c_uint32()
    check(lib().aqe_sql_finish(C.byref(q), SQL_MODE[mode], C.byref(layout), _ptr(acc), _ptr(exists) if exists is not None else None,
                               rows, layout.n_groups, C.byref(n)))
    return list(rows[: n.value])


def synth_rows_host(n: int, seed: int = 7, first_row: int = 0, dist: int = 0) -> np.ndarray:
    rows = np.empty(n, dtype=RECORD_DTYPE)
    check(lib().aqe_synth_rows_host(seed, first_row, n, dist, _ptr(rows)))
    return rows
