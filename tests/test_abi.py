"""The C-ABI library loads on a CPU-only box and exports every symbol include/aqe_b200.h declares; the
drop-in module imports and carries the reference's surface (bindings.cpp:10-137).  No compute calls."""
import os
import re

import pytest

import approximatequeryengine_b200 as aqe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aqe_b200.h")).read()
    return sorted(set(re.findall(r"^AQE_API\s+[\w \*]+?\b(aqe_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert len(syms) >= 45
    for must in ("aqe_open", "aqe_scan", "aqe_sum_f64", "aqe_sum_where_f64", "aqe_sum_i128", "aqe_plan_build",
                 "aqe_stats_from_indices", "aqe_gather_records", "aqe_approx", "aqe_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    L = aqe.lib()
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/aqe_b200.h but not exported"
    assert set(L._signatures) == set(declared_symbols()), "ctypes table out of sync with the header"
    assert L.aqe_abi_version() == 5


def test_host_only_entry_points_validate_their_arguments():
    """What runs before any CUDA call: merging no partials is the empty result; the several-device host scan refuses a device named
    twice, an empty device list and a NULL column without touching a GPU."""
    import ctypes as C
    L = aqe.lib()
    out = aqe.Partial()
    assert L.aqe_merge_partials(None, 0, 0, C.byref(out)) == 0 and out.count == 0 and out.sum == 0.0
    assert L.aqe_merge_partials(None, 2, 0, C.byref(out)) != 0
    two = (aqe.Partial * 2)()
    two[0].count, two[0].sum, two[1].count, two[1].sum = 3, 1.5, 4, 2.25
    assert L.aqe_merge_partials(two, 2, 0, C.byref(out)) == 0 and out.count == 7 and out.sum == 3.75
    col = (C.c_double * 4)(1.0, 2.0, 3.0, 4.0)
    for devs, n in (((0, 0), 2), ((0,), 0), ((0,) * 17, 17)):
        arr = (C.c_int * max(len(devs), 1))(*devs)
        assert L.aqe_scan_host_column_multi(arr, n, col, 0, 4, 0.0, 0.0, 0, C.byref(out)) != 0, devs
        assert L.aqe_last_error()
    one = (C.c_int * 1)(0)
    assert L.aqe_scan_host_column_multi(one, 1, None, 0, 4, 0.0, 0.0, 0, C.byref(out)) != 0
    assert L.aqe_scan_host_column_multi(one, 1, col, 7, 4, 0.0, 0.0, 0, C.byref(out)) != 0
    assert L.aqe_scan_host_column_multi(None, 1, col, 0, 4, 0.0, 0.0, 0, C.byref(out)) != 0


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(aqe.Partial) == 64
    assert C.sizeof(aqe.Stats) == 32
    assert C.sizeof(aqe.Segment) == 48
    assert C.sizeof(aqe.SampleParams) == 72
    assert C.sizeof(aqe.ApproxSpec) == 80
    assert C.sizeof(aqe.ApproxResult) == 104
    assert C.sizeof(aqe.StatsPartial) == 64
    assert C.sizeof(aqe.SqlTerm) == 56
    assert C.sizeof(aqe.SqlQuery) == 32 + 8 * 4 + 8 * 5 * 56 + 32 + 64 + 64 + 64 + 512
    assert C.sizeof(aqe.SqlRow) == 80
    assert C.sizeof(aqe.SqlFacts) == 32
    assert C.sizeof(aqe.SqlLayout) == 24
    assert aqe.RECORD_DTYPE.itemsize == 32


REFERENCE_DB_METHODS = {  # bindings.cpp:42-101, name -> keyword defaults
    "create_database": {}, "open_database": {}, "close_database": {}, "insert_record": {}, "sum_amount": {},
    "sum_amount_where": {}, "sample_records": {}, "optimized_sequential_sample": {}, "get_total_records": {},
    "get_node_count": {}, "save_to_file": {}, "load_from_file": {},
    "fast_pointer_sample": {"step_size": 2}, "slow_pointer_sample": {}, "dual_pointer_sample": {},
    "parallel_pointer_sample": {"num_threads": 4}, "random_pointer_sample": {"seed": 42},
    "clt_validated_dual_pointer_sample": {"confidence_level": 0.95, "check_interval": 10, "num_threads": 4, "max_error_percent": 2.0},
    "optimized_clt_sample": {"confidence_level": 0.95, "check_interval": 20, "num_threads": 4, "max_error_percent": 2.0},
    "block_sample": {"block_size": 1000}, "page_sample": {"page_size": 4096},
    "parallel_block_sample": {"block_size": 1000, "num_threads": 4},
    "adaptive_block_sample": {"min_block_size": 500, "max_block_size": 2000},
    "stratified_block_sample": {"block_size": 1000, "strata_count": 4}, "index_based_sample": {},
    "node_skip_sample": {"skip_factor": 2}, "balanced_tree_sample": {}, "direct_access_sample": {},
    "byte_offset_sample": {}, "random_start_nth_sample": {"nth": 10}, "memory_stride_sample": {"stride_bytes": 0},
    "address_arithmetic_sample": {}, "optimized_address_arithmetic_sample": {},
    "random_start_memory_stride_sample": {"stride_bytes": 0}, "multithreaded_memory_stride_sample": {"num_threads": 4},
    "fast_aggregated_memory_stride_sum": {"num_threads": 4}, "signal_based_clt_sample": {"check_interval": 10},
}


def test_dropin_module_surface():
    b = aqe.backend()
    assert len(REFERENCE_DB_METHODS) == 37
    for name, defaults in REFERENCE_DB_METHODS.items():
        f = getattr(b.CustomBPlusDB, name)
        doc = f.__doc__
        for k, v in defaults.items():
            assert re.search(rf"\b{k}: [^,)]+ = {re.escape(str(v))}(?![\w.])", doc), (name, k, doc)
    for name in ("create_database", "open_database", "close_database", "insert_record", "insert_batch", "execute_sum_query",
                 "execute_avg_query", "execute_count_query", "execute_exact_sum", "execute_exact_avg", "execute_exact_count",
                 "benchmark_query", "get_total_records", "get_tree_height", "get_database_size_mb"):
        assert hasattr(b.CustomApproximateScheduler, name)
    doc = b.CustomApproximateScheduler.execute_sum_query.__doc__
    assert re.search(r"sample_percent: [^,)]+ = 10.0, num_threads: [^,)]+ = 4\)", doc), doc
    for name in ("run_query", "run_query_groupby", "run_query_with_ci", "run_query_groupby_with_ci"):
        assert callable(getattr(b, name))
    assert [s for s in ("STABLE", "DRIFTING", "INSUFFICIENT_DATA", "ERROR") if hasattr(b.CustomApproximationStatus, s)] == ["STABLE", "DRIFTING", "INSUFFICIENT_DATA", "ERROR"]
    r = b.Record()
    assert (r.id, r.amount, r.region, r.product_id, r.timestamp) == (0, 0.0, 0, 0, 0)
    for f in ("value", "status", "confidence_level", "error_margin", "samples_used", "computation_time"):
        assert hasattr(b.CustomValidationResult, f)
    for f in ("value", "ci_lower", "ci_upper"):
        assert hasattr(b.QueryResult, f)
    assert not hasattr(b, "ApproximationStatus")  # the CLI probes this name and falls back (enhanced_aqe_cli.py:142-155)


def test_run_query_rejects_what_it_cannot_open(tmp_path):
    """run_query* take a record file; a missing file or a SQLite file (the reference's storage for this path) fail loudly."""
    b = aqe.backend()
    with pytest.raises(RuntimeError, match="Cannot open database"):
        b.run_query("SELECT SUM(amount) FROM sales", str(tmp_path / "missing.aqe"), 0)
    import sqlite3
    p = str(tmp_path / "x.db")
    con = sqlite3.connect(p); con.execute("CREATE TABLE sales (id INTEGER PRIMARY KEY, amount REAL)"); con.commit(); con.close()
    with pytest.raises(RuntimeError, match="SQLite file"):
        b.run_query("SELECT SUM(amount) FROM sales", p, 0)
    for name, defaults in (("run_query", "sample_percent: [^,)]+ = 0\\)"), ("run_query_groupby", "sample_percent: [^,)]+ = 0, num_threads: [^,)]+ = 4\\)"),
                           ("run_query_with_ci", "sample_percent: [^,)]+ = 0\\)"), ("run_query_groupby_with_ci", "sample_percent: [^,)]+ = 0, num_threads: [^,)]+ = 4\\)")):
        doc = getattr(b, name).__doc__
        assert re.search(r"sql_query: [^,)]+, db_path: [^,)]+, " + defaults, doc), doc   # bindings.cpp:126-136


def test_where_clause_forms():
    w = aqe.backend().CustomApproximateScheduler._where_conditions  # custom_scheduler.cpp:277-294
    assert w("SELECT SUM(amount) FROM s WHERE amount BETWEEN 100 AND 500") == (100.0, 500.0)
    assert w("... WHERE amount  BETWEEN 1.5   AND 2.25") == (1.5, 2.25)
    assert w("WHERE amount >= 3 AND amount <= 7") == (3.0, 7.0)
    assert w("WHERE amount>9") == (9.0, 99999.99)
    assert w("WHERE amount between 1 and 2") == (-1.0, -1.0)  # keywords are case-sensitive in the reference
    assert w("SELECT SUM(amount) FROM s") == (-1.0, -1.0)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    b = aqe.backend()
    if b.device_count() > 0:
        pytest.skip("a GPU is present")
    db = b.CustomBPlusDB()
    r = b.Record(); r.id = 1; r.amount = 2.0
    assert db.insert_record(r) and db.get_total_records() == 1
    with pytest.raises(RuntimeError, match="cuda|CUDA"):
        db.sum_amount()
    with pytest.raises(RuntimeError, match="cuda|CUDA"):
        db.query("SELECT SUM(amount) FROM sales")           # the SQL path has no host evaluator either
    with pytest.raises(RuntimeError, match="cuda|CUDA"):
        db.query_groupby("SELECT COUNT(amount) FROM sales GROUP BY region")
    e = aqe.Engine(0)
    with pytest.raises(aqe.AqeError):
        e.generate(100)


def test_reference_frontend_package_runs_on_top_of_the_module(tmp_path):
    """The reference's own Python package (src/aqe_frontend, unmodified, imported from /root/reference where that exists)
    finds `aqe_backend`, and its run_query / run_query_groupby forward to this engine: parser errors come back as the
    reference's RuntimeError, and without a GPU the call stops at the device boundary (no CPU evaluation)."""
    import importlib
    import struct
    import sys
    ref_root = os.environ.get("AQE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "src", "aqe_frontend")):
        pytest.skip("reference checkout not present")
    b = aqe.backend()                       # puts _lib/ on sys.path: `import aqe_backend` inside runner.py resolves to it
    sys.path.insert(0, ref_root)
    try:
        fe = importlib.import_module("src.aqe_frontend")
        runner = importlib.import_module("src.aqe_frontend.runner")
    finally:
        sys.path.remove(ref_root)
    assert runner.aqe_backend is b
    path = str(tmp_path / "sales.aqe")
    rows = aqe.synth_rows_host(1000, seed=7)
    with open(path, "wb") as f:
        f.write(struct.pack("<QQQ", len(rows), 1, len(rows)) + rows.tobytes())
    with pytest.raises(RuntimeError, match="Unsupported aggregation function"):
        fe.run_query("SELECT MAX(amount) FROM sales", path, 0)
    if b.device_count() == 0:
        with pytest.raises(RuntimeError, match="cuda|CUDA"):
            fe.run_query("SELECT SUM(amount) FROM sales", path, 10)
    else:
        assert fe.run_query("SELECT COUNT(amount) FROM sales", path, 0) == 1000.0
        assert set(runner.run_query_groupby("SELECT COUNT(amount) FROM sales GROUP BY region", path, 0, 4)) == {str(k) for k in range(8)}
    assert fe.parse_query("SELECT SUM(amount) FROM sales;") == {"agg_func": "SUM", "column": "AMOUNT", "table": "SALES"}


def test_bench_refuses_to_run_without_a_gpu():
    """bench.py's own arm has no CPU path: on a box without a device it exits non-zero with a message instead of timing anything."""
    import subprocess
    import sys
    if aqe.backend().device_count() > 0:
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout) and not r.stdout.strip().startswith("{")
