"""oracle -- TEST INFRASTRUCTURE (the checker), never the product.

ctypes front-ends for

* ``Oracle``  : oracle/libaqe_oracle.so, the plain-C restatement of the reference hot path
                (oracle/aqe_oracle.c; every function cites the reference file:line it follows);
* ``Ref``     : oracle/_ref/libaqe_ref.so, the UNMODIFIED reference core compiled in place from
                /root/reference by oracle/Makefile and driven through oracle/ref_harness.cpp.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (approximatequeryengine_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libaqe_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libaqe_ref.so")
REFSQL_SO = os.path.join(HERE, "_ref", "libaqe_refsql.so")
REFCLI_PYC = os.path.join(HERE, "_ref", "enhanced_aqe_cli.bytecode")   # the reference's command line as bytecode (make refcli)

RECORD_DTYPE = np.dtype(
    [("id", "<i8"), ("amount", "<f8"), ("region", "<i4"), ("product_id", "<i4"), ("timestamp", "<i8")]
)
assert RECORD_DTYPE.itemsize == 32

# aqe_method ids (include/aqe_b200.h)
METHODS = {
    "slow_pointer": 0, "fast_pointer": 1, "dual_pointer": 2, "parallel_pointer": 3, "random_pointer": 4,
    "memory_stride": 5, "optimized_address_arithmetic": 6, "index_based": 7, "byte_offset": 8,
    "optimized_clt": 9, "block": 10, "page": 11, "parallel_block": 12, "node_skip": 13, "balanced_tree": 14,
    "direct_access": 15, "adaptive_block": 16, "stratified_block": 17, "sample_records": 18,
    "optimized_sequential": 19, "random_start_nth": 20, "address_arithmetic": 21,
    "random_start_memory_stride": 22, "multithreaded_memory_stride": 23, "clt_validated_dual_pointer": 24,
    "signal_based_clt": 25,
}
COLS = {"id": 0, "amount": 1, "region": 2, "product_id": 3, "timestamp": 4, None: -1}
AGG = {"sum": 0, "avg": 1, "count": 2}


class SampleParams(C.Structure):
    _fields_ = [
        ("sample_percent", C.c_double), ("step_size", C.c_int64), ("num_threads", C.c_int64),
        ("block_size", C.c_int64), ("block_size_max", C.c_int64), ("check_interval", C.c_int64),
        ("confidence_level", C.c_double), ("max_error_percent", C.c_double), ("seed", C.c_uint64),
    ]


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


def make_params(method: str, sample_percent: float, **kw) -> SampleParams:
    """pybind defaults of the reference (bindings.cpp:56-101), overridable by keyword."""
    m = METHODS[method]
    p = SampleParams(
        sample_percent=sample_percent,
        step_size=10 if method == "random_start_nth" else 2,
        num_threads=4,
        block_size={"page": 4096, "adaptive_block": 500, "memory_stride": 0, "random_start_memory_stride": 0}.get(method, 1000),
        block_size_max=4 if method == "stratified_block" else 2000,
        check_interval=20 if method == "optimized_clt" else 10,
        confidence_level=0.95, max_error_percent=2.0, seed=42,
    )
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the checker libraries (oracle always; oracle/_ref only where /root/reference exists)."""
    targets = ["oracle"]
    if ref and os.path.isdir(os.environ.get("AQE_REFERENCE", "/root/reference")):
        targets.append("ref")
        if any(os.path.exists(p) for p in ("/usr/lib/x86_64-linux-gnu/libsqlite3.so.0", "/usr/lib64/libsqlite3.so.0", "/usr/lib/libsqlite3.so.0")):
            targets.append("refsql")
        targets.append("refcli")
    cmd = ["make", "-C", HERE, "-s"] + targets
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)


def _rows(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=RECORD_DTYPE)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class SqlParsed(C.Structure):
    _fields_ = [("agg", C.c_char * 32), ("column", C.c_char * 64), ("table", C.c_char * 64), ("where", C.c_char * 512),
                ("group_by", C.c_char * 64)]


class SqlRow(C.Structure):
    _fields_ = [("key", C.c_int64), ("value", C.c_double), ("ci_lower", C.c_double), ("ci_upper", C.c_double)]


class SqlError(Exception):
    """kind: 'runtime_error' (std::runtime_error in the reference), 'stod' (std::stod on "NULL": ValueError through
    pybind11), 'unsupported' (valid SQL outside the restated grammar)."""

    def __init__(self, kind: str, msg: str):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.msg = msg


SQL_MODES = {"run_query": 0, "run_query_with_ci": 1, "run_query_groupby": 2, "run_query_groupby_with_ci": 3}
_SQL_ERR = {1: "runtime_error", 2: "stod", 3: "unsupported"}


class Oracle:
    """The C restatement."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        self.L = L
        L.orc_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_save_file.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        L.orc_file_count.argtypes = [C.c_char_p]; L.orc_file_count.restype = C.c_int64
        L.orc_load_file.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_sum_amount.argtypes = [C.c_void_p, C.c_uint64]; L.orc_sum_amount.restype = C.c_double
        L.orc_sum_amount_ld.argtypes = [C.c_void_p, C.c_uint64]; L.orc_sum_amount_ld.restype = C.c_double
        L.orc_avg_amount.argtypes = [C.c_void_p, C.c_uint64]; L.orc_avg_amount.restype = C.c_double
        L.orc_sum_amount_where.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_uint64)]
        L.orc_sum_amount_where.restype = C.c_double
        L.orc_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Partial)]
        L.orc_indices.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(SampleParams), C.c_void_p, C.c_uint64]
        L.orc_indices.restype = C.c_int64
        L.orc_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Stats)]
        L.orc_estimate.argtypes = [C.POINTER(Stats), C.c_uint64, C.c_int, C.c_double, C.c_int,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_fast_aggregated.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SampleParams), C.POINTER(C.c_uint64)]
        L.orc_fast_aggregated.restype = C.c_double
        L.orc_z_score.argtypes = [C.c_double, C.c_int]; L.orc_z_score.restype = C.c_double
        L.orc_approx.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(ApproxSpec), C.POINTER(ApproxResult)]
        L.orc_approx_sharded.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(ApproxSpec), C.POINTER(ApproxResult)]
        L.orc_draw_position.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64]
        L.orc_draw_position.restype = C.c_uint64
        L.orc_scan_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(C.c_uint64)]
        L.orc_scan_mt.restype = C.c_double
        L.orc_synth_amount_mt.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.orc_sum_col_serial.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint64)]
        L.orc_sum_col_serial.restype = C.c_double
        L.orc_sum_col_exact_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_sum_col_exact_mt.restype = C.c_double
        for f in ("orc_tree_height", "orc_leaf_count", "orc_node_count"):
            getattr(L, f).argtypes = [C.c_uint64]; getattr(L, f).restype = C.c_uint64
        L.orc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32 * 4)]
        L.orc_sql_parse.argtypes = [C.c_char_p, C.POINTER(SqlParsed), C.c_char_p, C.c_size_t]
        L.orc_sql_run.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_int, C.c_int, C.POINTER(SqlRow), C.c_uint32,
                                  C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]

    # -- data -------------------------------------------------------------------------------------
    def synth(self, n: int, seed: int = 7, first_row: int = 0, dist: int = 0) -> np.ndarray:
        rows = np.empty(n, dtype=RECORD_DTYPE)
        self.L.orc_synth_rows(seed, first_row, n, dist, _ptr(rows))
        return rows

    def synth_amount(self, n: int, seed: int = 7, first_row: int = 0, dist: int = 0, threads: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """The amount column of rows [first_row, first_row + n) of the synthetic table, generated on `threads` host threads."""
        col = np.empty(n, dtype=np.float64) if out is None else out
        self.L.orc_synth_amount_mt(seed, first_row, n, dist, _ptr(col), threads or (os.cpu_count() or 1))
        return col

    def sum_col_serial(self, col: np.ndarray, pred=None):
        """(sum, count) the way the reference's loops add: one double, strictly left to right (cbd:242-251, :263-274)."""
        c = C.c_uint64()
        s = self.L.orc_sum_col_serial(_ptr(col), len(col), int(pred is not None), pred[0] if pred else 0.0, pred[1] if pred else 0.0, C.byref(c))
        return s, c.value

    def sum_col_exact(self, col: np.ndarray, pred=None, threads: int | None = None):
        """(sum, count) with long double Neumaier accumulation: the exactly rounded sum for all practical purposes."""
        c = C.c_uint64()
        s = self.L.orc_sum_col_exact_mt(_ptr(col), len(col), int(pred is not None), pred[0] if pred else 0.0, pred[1] if pred else 0.0,
                                        threads or (os.cpu_count() or 1), C.byref(c))
        return s, c.value

    def save_file(self, path: str, rows) -> None:
        rows = _rows(rows)
        if self.L.orc_save_file(path.encode(), _ptr(rows), len(rows)) != 0:
            raise IOError(path)

    def load_file(self, path: str) -> np.ndarray:
        n = self.L.orc_file_count(path.encode())
        if n < 0:
            raise IOError(path)
        rows = np.empty(n, dtype=RECORD_DTYPE)
        got = C.c_uint64()
        if self.L.orc_load_file(path.encode(), _ptr(rows), n, C.byref(got)) != 0:
            raise IOError(path)
        return rows

    # -- exact ------------------------------------------------------------------------------------
    def sum_amount(self, rows) -> float:
        rows = _rows(rows); return self.L.orc_sum_amount(_ptr(rows), len(rows))

    def sum_amount_exactly_rounded(self, rows) -> float:
        rows = _rows(rows); return self.L.orc_sum_amount_ld(_ptr(rows), len(rows))

    def avg_amount(self, rows) -> float:
        rows = _rows(rows); return self.L.orc_avg_amount(_ptr(rows), len(rows))

    def sum_amount_where(self, rows, lo: float, hi: float):
        rows = _rows(rows); c = C.c_uint64()
        s = self.L.orc_sum_amount_where(_ptr(rows), len(rows), lo, hi, C.byref(c))
        return s, c.value

    def scan(self, rows, agg_col="amount", pred_col=None, lo=0.0, hi=0.0) -> Partial:
        rows = _rows(rows); out = Partial()
        self.L.orc_scan(_ptr(rows), len(rows), COLS[agg_col], COLS[pred_col], lo, hi, C.byref(out))
        return out

    def scan_mt(self, base: np.ndarray, aos: bool, threads: int, pred=None):
        c = C.c_uint64()
        lo, hi = pred if pred else (0.0, 0.0)
        s = self.L.orc_scan_mt(_ptr(base), len(base), int(aos), int(pred is not None), lo, hi, threads, C.byref(c))
        return s, c.value

    # -- samplers ---------------------------------------------------------------------------------
    def indices(self, rows, method: str, params: SampleParams, n_rows: int | None = None) -> np.ndarray:
        """Positions (into rows-in-ascending-id) chosen by `method`.  rows may be None for methods that
        do not read data (then n_rows is required)."""
        if rows is not None:
            rows = _rows(rows); n = len(rows); ptr = _ptr(rows)
        else:
            n = int(n_rows); ptr = None
        cap = max(16, int(n * min(max(params.sample_percent, 0.0), 100.0) / 100.0 * 2.2) + 64)
        while True:
            out = np.empty(cap, dtype=np.int64)
            got = self.L.orc_indices(ptr, n, METHODS[method], C.byref(params), _ptr(out), cap)
            if got < 0:
                raise ValueError(f"oracle status {-got} for {method}")
            if got <= cap:
                return out[:got].copy()
            cap = got

    def stats(self, rows, idx, col="amount") -> Stats:
        rows = _rows(rows); idx = np.ascontiguousarray(idx, dtype=np.int64); s = Stats()
        self.L.orc_stats(_ptr(rows), _ptr(idx), len(idx), COLS[col], C.byref(s))
        return s

    def estimate(self, s: Stats, population: int, agg: str, z: float = 1.96, legacy_ci: bool = False):
        e, lo, hi = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_estimate(C.byref(s), population, AGG[agg], z, int(legacy_ci), C.byref(e), C.byref(lo), C.byref(hi))
        return e.value, lo.value, hi.value

    def fast_aggregated(self, rows, params: SampleParams):
        rows = _rows(rows); n = C.c_uint64()
        return self.L.orc_fast_aggregated(_ptr(rows), len(rows), C.byref(params), C.byref(n)), n.value

    def z_score(self, conf: float, exact: bool = True) -> float:
        return self.L.orc_z_score(conf, int(exact))

    def approx(self, rows, spec: ApproxSpec) -> ApproxResult:
        rows = _rows(rows); out = ApproxResult()
        self.L.orc_approx(_ptr(rows), len(rows), C.byref(spec), C.byref(out))
        return out

    def approx_sharded(self, rows, world: int, spec: ApproxSpec) -> ApproxResult:
        rows = _rows(rows); out = ApproxResult()
        if self.L.orc_approx_sharded(_ptr(rows), len(rows), world, C.byref(spec), C.byref(out)) != 0:
            raise ValueError("orc_approx_sharded")
        return out

    def draw_position(self, seed: int, design: int, j: int, units: int) -> int:
        return self.L.orc_draw_position(seed, design, j, units)

    def philox(self, key: int, ctr_lo: int, ctr_hi: int = 0):
        out = (C.c_uint32 * 4)()
        self.L.orc_philox(key, ctr_lo, ctr_hi, C.byref(out))
        return list(out)

    # -- SQL-string path (executor.cpp / parser.cpp) ----------------------------------------------
    def sql_parse(self, sql: str) -> dict:
        q = SqlParsed(); err = C.create_string_buffer(256)
        rc = self.L.orc_sql_parse(sql.encode(), C.byref(q), err, 256)
        if rc:
            raise SqlError(_SQL_ERR[rc], err.value.decode())
        return {k: getattr(q, k).decode() for k in ("agg", "column", "table", "where", "group_by")}

    def sql(self, rows, sql: str, sample_percent: int = 0, mode: str = "run_query"):
        """[(key, value, ci_lower, ci_upper)] in ascending numeric key order (one entry, key 0, without GROUP BY)."""
        rows = _rows(rows)
        cap = 8192
        out = (SqlRow * cap)(); n = C.c_uint32(); err = C.create_string_buffer(256)
        rc = self.L.orc_sql_run(_ptr(rows), len(rows), sql.encode(), sample_percent, SQL_MODES[mode], out, cap, C.byref(n), err, 256)
        if rc:
            raise SqlError(_SQL_ERR[rc], err.value.decode())
        return [(out[i].key, out[i].value, out[i].ci_lower, out[i].ci_upper) for i in range(min(n.value, cap))]

    def tree_height(self, n): return self.L.orc_tree_height(n)
    def leaf_count(self, n): return self.L.orc_leaf_count(n)
    def node_count(self, n): return self.L.orc_node_count(n)


class SchedResult(C.Structure):
    _fields_ = [("value", C.c_double), ("status", C.c_int), ("confidence_level", C.c_double),
                ("error_margin", C.c_double), ("samples_used", C.c_int), ("ms", C.c_double)]


class Ref:
    """The unmodified reference CustomBPlusDB, loaded through insert_batch (open_database deadlocks)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.ref_new.restype = C.c_void_p
            L.ref_free.argtypes = [C.c_void_p]
            L.ref_insert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            L.ref_insert_record.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_save.argtypes = [C.c_void_p, C.c_char_p]
            for f in ("ref_total", "ref_node_count", "ref_tree_height"):
                getattr(L, f).argtypes = [C.c_void_p]; getattr(L, f).restype = C.c_uint64
            for f in ("ref_sum_amount", "ref_avg_amount"):
                getattr(L, f).argtypes = [C.c_void_p]; getattr(L, f).restype = C.c_double
            L.ref_sum_amount_where.argtypes = [C.c_void_p, C.c_double, C.c_double]; L.ref_sum_amount_where.restype = C.c_double
            L.ref_fast_aggregated.argtypes = [C.c_void_p, C.c_double, C.c_int]; L.ref_fast_aggregated.restype = C.c_double
            L.ref_parallel_sum_sample.argtypes = [C.c_void_p, C.c_double, C.c_int]; L.ref_parallel_sum_sample.restype = C.c_double
            L.ref_sample.argtypes = [C.c_void_p, C.c_int, C.POINTER(SampleParams), C.c_void_p, C.c_uint64]
            L.ref_sample.restype = C.c_int64
            L.ref_time.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]
            L.ref_time.restype = C.c_double
            L.ref_sched_new.argtypes = [C.c_double]; L.ref_sched_new.restype = C.c_void_p
            L.ref_sched_free.argtypes = [C.c_void_p]
            L.ref_sched_insert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            L.ref_sched_exec.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double, C.c_int, C.POINTER(SchedResult)]
            L.ref_sched_size_mb.argtypes = [C.c_void_p]; L.ref_sched_size_mb.restype = C.c_double
            cls._lib = L
        return cls._lib

    def __init__(self, rows=None):
        self.L = self.lib()
        self.h = C.c_void_p(self.L.ref_new())
        if rows is not None:
            self.insert_batch(rows)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_free(self.h); self.h = None

    def insert_batch(self, rows):
        rows = _rows(rows)
        if self.L.ref_insert_batch(self.h, _ptr(rows), len(rows)) != 0:
            raise RuntimeError("insert_batch failed")

    def insert_record(self, row):
        row = _rows(row)
        assert len(row) == 1
        if self.L.ref_insert_record(self.h, _ptr(row)) != 0:
            raise RuntimeError("insert_record failed")

    def save(self, path: str):
        if self.L.ref_save(self.h, path.encode()) != 0:
            raise IOError(path)

    def total(self): return self.L.ref_total(self.h)
    def node_count(self): return self.L.ref_node_count(self.h)
    def tree_height(self): return self.L.ref_tree_height(self.h)
    def sum_amount(self): return self.L.ref_sum_amount(self.h)
    def avg_amount(self): return self.L.ref_avg_amount(self.h)
    def sum_amount_where(self, lo, hi): return self.L.ref_sum_amount_where(self.h, lo, hi)
    def fast_aggregated(self, p, threads=4): return self.L.ref_fast_aggregated(self.h, p, threads)
    def parallel_sum_sample(self, p, threads=4): return self.L.ref_parallel_sum_sample(self.h, p, threads)

    def sample(self, method: str, params: SampleParams) -> np.ndarray:
        cap = 1 << 16
        while True:
            out = np.empty(cap, dtype=RECORD_DTYPE)
            got = self.L.ref_sample(self.h, METHODS[method], C.byref(params), _ptr(out), cap)
            if got < 0:
                raise ValueError(method)
            if got <= cap:
                return out[:got].copy()
            cap = int(got)

    def time(self, what: int, reps: int, a: float = 0.0, b: float = 0.0):
        v = C.c_double()
        t = self.L.ref_time(self.h, what, reps, a, b, C.byref(v))
        return t, v.value


class RefScheduler:
    def __init__(self, rows, error_threshold=0.05):
        self.L = Ref.lib()
        self.h = C.c_void_p(self.L.ref_sched_new(error_threshold))
        rows = _rows(rows)
        self.L.ref_sched_insert_batch(self.h, _ptr(rows), len(rows))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_sched_free(self.h); self.h = None

    def run(self, what: int, query: str = "", p: float = 10.0, threads: int = 4) -> SchedResult:
        r = SchedResult()
        self.L.ref_sched_exec(self.h, what, query.encode(), p, threads, C.byref(r))
        return r

    def size_mb(self): return self.L.ref_sched_size_mb(self.h)


class RefSqlRow(C.Structure):
    _fields_ = [("key", C.c_char * 64), ("value", C.c_double), ("ci_lower", C.c_double), ("ci_upper", C.c_double)]


class RefSql:
    """The unmodified reference SQL-string path (executor.cpp, parser.cpp, core/db.cpp) over a SQLite file holding the
    record table as ``sales(id INTEGER PRIMARY KEY, amount REAL, region INTEGER, product_id INTEGER, timestamp INTEGER)``."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REFSQL_SO)

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REFSQL_SO)
            L.ref_sql_error.restype = C.c_char_p
            L.ref_sql_parse.argtypes = [C.c_char_p, C.c_int] + [C.c_char_p] * 5 + [C.c_size_t]
            L.ref_run_query.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
            L.ref_run_query_with_ci.argtypes = [C.c_char_p, C.c_char_p, C.c_int] + [C.POINTER(C.c_double)] * 3
            for f in ("ref_run_query_groupby", "ref_run_query_groupby_with_ci"):
                getattr(L, f).argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(RefSqlRow), C.c_size_t, C.POINTER(C.c_size_t)]
            cls._lib = L
        return cls._lib

    @staticmethod
    def make_sqlite(path: str, rows, table: str = "sales") -> None:
        import sqlite3
        if os.path.exists(path):
            os.remove(path)
        con = sqlite3.connect(path)
        con.execute(f"CREATE TABLE {table} (id INTEGER PRIMARY KEY, amount REAL, region INTEGER, product_id INTEGER, timestamp INTEGER)")
        rows = _rows(rows)
        con.executemany(f"INSERT INTO {table} VALUES (?,?,?,?,?)",
                        zip(rows["id"].tolist(), rows["amount"].tolist(), rows["region"].tolist(), rows["product_id"].tolist(), rows["timestamp"].tolist()))
        con.commit()
        con.close()

    def __init__(self, sqlite_path: str):
        self.L = self.lib()
        self.path = sqlite_path.encode()

    def _err(self):
        msg = self.L.ref_sql_error().decode()
        return SqlError("stod" if msg == "stod" else "runtime_error", msg)

    def parse(self, sql: str, p: int = 0) -> dict:
        bufs = [C.create_string_buffer(512) for _ in range(5)]
        if self.L.ref_sql_parse(sql.encode(), p, *bufs, 512):
            raise self._err()
        return dict(zip(("agg", "column", "table", "where", "group_by"), (b.value.decode() for b in bufs)))

    def run(self, sql: str, p: int = 0, mode: str = "run_query", threads: int = 4):
        """Same shape as Oracle.sql (keys as int, ascending numeric order)."""
        if mode == "run_query":
            v = C.c_double()
            if self.L.ref_run_query(sql.encode(), self.path, p, C.byref(v)):
                raise self._err()
            return [(0, v.value, v.value, v.value)]
        if mode == "run_query_with_ci":
            v, lo, hi = C.c_double(), C.c_double(), C.c_double()
            if self.L.ref_run_query_with_ci(sql.encode(), self.path, p, C.byref(v), C.byref(lo), C.byref(hi)):
                raise self._err()
            return [(0, v.value, lo.value, hi.value)]
        cap = 8192
        out = (RefSqlRow * cap)(); n = C.c_size_t()
        f = self.L.ref_run_query_groupby if mode == "run_query_groupby" else self.L.ref_run_query_groupby_with_ci
        if f(sql.encode(), self.path, p, threads, out, cap, C.byref(n)):
            raise self._err()
        return sorted((int(out[i].key.decode()), out[i].value, out[i].ci_lower, out[i].ci_upper) for i in range(min(n.value, cap)))
