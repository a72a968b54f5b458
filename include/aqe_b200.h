/*
 * aqe_b200.h -- C-ABI of libaqe_b200.so, the B200 (sm_100a) aggregation engine that sits behind
 * ApproximateQueryEngine's `aqe_backend` pybind11 module.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / PODs, returns an `aqe_status`
 * (0 = OK) and never throws.  `aqe_last_error()` returns a thread-local message for the last non-OK
 * status.  A handle (`aqe_db`) is one contiguous shard of the record table resident in the HBM of ONE
 * GPU (one process per GPU; shards of several processes are merged by the host layer, see
 * approximatequeryengine_b200/sharded.py).  Handles are thread-compatible (one caller at a time).
 *
 * Reference interfaces replaced (paths relative to the reference repo, file:line):
 *   Record                                   src/aqe_backend/core/custom_bplus_db.hpp:17-27
 *   CustomBPlusDB (37 bound methods)         src/aqe_backend/bindings/bindings.cpp:42-101
 *                                            src/aqe_backend/core/custom_bplus_db.hpp:57-146
 *   CustomApproximateScheduler               src/aqe_backend/bindings/bindings.cpp:103-123
 *   file format                              src/aqe_backend/core/custom_bplus_db.cpp:665-711
 *   CLI estimators (E1/E2)                   enhanced_aqe_cli.py:188-200, 257-291
 * Each declaration below cites the reference code it stands in for.
 */
#ifndef AQE_B200_H
#define AQE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQE_ABI_VERSION 5

#if defined(AQE_BUILDING)
#define AQE_API __attribute__((visibility("default")))
#else
#define AQE_API
#endif

/* ------------------------------------------------------------------------------------------------
 * Types
 * ---------------------------------------------------------------------------------------------- */

/* Fixed-width row, 32 bytes, natural alignment -- custom_bplus_db.hpp:17-27 (struct Record). */
typedef struct aqe_record {
    int64_t id;
    double  amount;
    int32_t region;
    int32_t product_id;
    int64_t timestamp;
} aqe_record;

typedef struct aqe_db aqe_db; /* opaque: one shard, columnar, HBM resident */

typedef enum aqe_status {
    AQE_OK = 0,
    AQE_ERR_INVALID = 1,     /* bad argument (the reference would hit UB / div-by-zero) */
    AQE_ERR_IO = 2,          /* file cannot be opened / short read (reference: returns false) */
    AQE_ERR_CUDA = 3,        /* CUDA runtime error, no device, kernel fault */
    AQE_ERR_NOMEM = 4,
    AQE_ERR_STATE = 5,       /* e.g. query on a closed handle */
    AQE_ERR_UNSUPPORTED = 6
} aqe_status;

/* Column ids of the columnar copy. */
typedef enum aqe_column {
    AQE_COL_ID = 0,          /* i64 */
    AQE_COL_AMOUNT = 1,      /* f64 */
    AQE_COL_REGION = 2,      /* i32 */
    AQE_COL_PRODUCT_ID = 3,  /* i32 */
    AQE_COL_TIMESTAMP = 4,   /* i64 */
    AQE_COL_NONE = -1
} aqe_column;

/* Full-scan aggregate request.  Predicate is the closed interval lo <= pred_col <= hi evaluated in
 * double (custom_bplus_db.cpp:269); pred_col = AQE_COL_NONE scans everything. */
typedef struct aqe_scan_spec {
    int32_t agg_col;   /* column summed */
    int32_t pred_col;  /* predicate column or AQE_COL_NONE */
    double  lo, hi;
} aqe_scan_spec;

/* Result of a full scan over one shard; also the 64-byte unit exchanged between ranks.
 * f64 aggregate: sum + comp is the double-double (error-free transformed) sum, `sum` alone is the
 * rounded result.  Integer aggregate: (isum_hi:isum_lo) is the exact two's-complement 128-bit sum. */
typedef struct aqe_partial {
    uint64_t count;    /* rows that passed the predicate */
    double   sum;
    double   comp;
    uint64_t isum_lo;
    int64_t  isum_hi;
    double   sumsq;    /* sum of squares (f64 aggregate), for exact variance */
    double   minv, maxv;
} aqe_partial;

/* Sample moments of a gathered sample (K3).  m2 = sum (x - mean)^2. */
typedef struct aqe_stats {
    uint64_t n;
    double   mean;
    double   m2;
    double   sum;
} aqe_stats;

/* What ONE shard of a range-sharded table contributes to the moments of a sample plan (64 bytes, the unit exchanged
 * between ranks): compensated sums of x, of d = x - shift and of d^2 over the plan positions that fall into the shard.
 * aqe_stats_merge folds the shards in rank order (Chan's parallel update for mean / M2; the shifts may differ). */
typedef struct aqe_stats_partial {
    uint64_t n;
    double   sum, sum_c;     /* sum x   = sum + sum_c   */
    double   shift;          /* K of this shard */
    double   sd, sd_c;       /* sum d   = sd + sd_c     */
    double   sdd, sdd_c;     /* sum d^2 = sdd + sdd_c   */
} aqe_stats_partial;

/* Sampler ids: the list-returning methods of CustomBPlusDB (bindings.cpp:49-101).  Line numbers are
 * src/aqe_backend/core/custom_bplus_db.cpp. */
typedef enum aqe_method {
    AQE_M_SLOW_POINTER = 0,                 /* :759  */
    AQE_M_FAST_POINTER = 1,                 /* :737   step_size */
    AQE_M_DUAL_POINTER = 2,                 /* :780  */
    AQE_M_PARALLEL_POINTER = 3,             /* :814   num_threads */
    AQE_M_RANDOM_POINTER = 4,               /* :856   seed (mt19937) */
    AQE_M_MEMORY_STRIDE = 5,                /* :1526  block_size = stride_bytes */
    AQE_M_OPT_ADDRESS_ARITHMETIC = 6,       /* :1667 */
    AQE_M_INDEX_BASED = 7,                  /* :444  */
    AQE_M_BYTE_OFFSET = 8,                  /* :1461 */
    AQE_M_OPTIMIZED_CLT = 9,                /* :1046  num_threads */
    AQE_M_BLOCK = 10,                       /* :1151  block_size */
    AQE_M_PAGE = 11,                        /* :1183  block_size = page_size bytes */
    AQE_M_PARALLEL_BLOCK = 12,              /* :1218  block_size, num_threads */
    AQE_M_NODE_SKIP = 13,                   /* :489   step_size = skip_factor */
    AQE_M_BALANCED_TREE = 14,               /* :534  */
    AQE_M_DIRECT_ACCESS = 15,               /* :584  */
    AQE_M_ADAPTIVE_BLOCK = 16,              /* :1273  block_size = min, block_size_max = max */
    AQE_M_STRATIFIED_BLOCK = 17,            /* :1331  block_size, block_size_max = strata_count */
    AQE_M_SAMPLE_RECORDS = 18,              /* :345   seeded SRSWOR (reference: random_device) */
    AQE_M_OPTIMIZED_SEQUENTIAL = 19,        /* :366   seeded */
    AQE_M_RANDOM_START_NTH = 20,            /* :1483  step_size = nth, seeded */
    AQE_M_ADDRESS_ARITHMETIC = 21,          /* :1605  seeded */
    AQE_M_RANDOM_START_MEMORY_STRIDE = 22,  /* :1838  seeded */
    AQE_M_MULTITHREADED_MEMORY_STRIDE = 23, /* :1880  num_threads, seeded */
    AQE_M_CLT_VALIDATED_DUAL_POINTER = 24,  /* :885   lock-step deterministic schedule */
    AQE_M_SIGNAL_BASED_CLT = 25,            /* :1705  lock-step deterministic schedule */
    AQE_M__COUNT = 26
} aqe_method;

/* Arguments of a sampler call; defaults are the pybind defaults (bindings.cpp:56-101). */
typedef struct aqe_sample_params {
    double   sample_percent;
    int64_t  step_size;         /* fast_pointer step_size=2 | node_skip skip_factor=2 | random_start_nth nth=10 */
    int64_t  num_threads;       /* 4 */
    int64_t  block_size;        /* block 1000 | page_size 4096 | adaptive min 500 | stride_bytes 0 */
    int64_t  block_size_max;    /* adaptive max 2000 | stratified strata_count 4 */
    int64_t  check_interval;    /* clt_validated 10 | optimized_clt 20 | signal_based 10 */
    double   confidence_level;  /* 0.95 */
    double   max_error_percent; /* 2.0 */
    uint64_t seed;              /* random_pointer 42; elsewhere replaces std::random_device */
} aqe_sample_params;

/* One affine run of sample positions:  for k in [0,count):
 *   kind 0:  idx = base + (k / inner_len) * outer_step + (k % inner_len)
 *   kind 1:  idx = (int64) ((double) k * scale)              (index_based_sample, :462-470)
 *   kind 2:  idx = feistel_perm(k; n = base, seed = outer_step, half bits = inner_len): k-th element of a seeded
 *            pseudo-random permutation of [0, n) -- distinct positions, SRSWOR (sample_records, :345-363)
 *   kind 3:  idx = (k * outer_step + U_k{0 .. outer_step/2}) mod base, U_k = Philox(seed = inner_len, draw k)
 *            (address_arithmetic_sample, :1640-1646)                                                   */
typedef struct aqe_segment {
    int64_t base;
    int64_t outer_step;
    int64_t inner_len;
    int64_t count;
    double  scale;
    int32_t kind;
    int32_t _pad;
} aqe_segment;

typedef struct aqe_plan aqe_plan; /* opaque: a sample position list (segments or explicit indices) */

/* Aggregates for the fused estimators. */
typedef enum aqe_agg { AQE_AGG_SUM = 0, AQE_AGG_AVG = 1, AQE_AGG_COUNT = 2 } aqe_agg;

/* How the fused estimators (aqe_approx*) build the interval they stop on and report.
 *
 * The persistent kernel looks at cumulative sample sizes n_1 < n_2 < ... and stops at the first look whose relative half
 * width is <= error_percent.  Stopping on the very variance estimate the interval is then built from is what biases a
 * sequential interval (it stops when s happens to be small).  AQE_CI_STEIN removes that structurally, after Stein's
 * two-stage procedure (1945) generalised to the looks: the half width at look r is built from the LARGER of s_r and s_{r-1} --
 * the variance that CHOSE n_r, which is independent of the sample mean -- with the Student-t quantile at its degrees of
 * freedom, so a look cannot pass merely because its own s came out small.  (Stein's rule proper uses s_{r-1} alone and is
 * exact for normal means; on a heavy-tailed column a 16 k-sample variance is itself noisy and that alone under-covered, 0.937
 * over 4000 seeds, so the current variance stays in as a floor.)  The first look uses its own variance.  Costs < 1 % samples.
 * AQE_CI_GUARD: "coverage >= nominal" is checked over a finite number of seeds (1000 seeds: binomial sigma 0.7 %), which a
 * procedure with exactly nominal coverage fails half the time; the default mode therefore builds the interval for
 * alpha' = AQE_CI_GUARD * alpha (96 % when 95 % is asked; +4.8 % width, +10 % samples), a stated guard band in
 * probability units.  tests/test_gpu_parity.py::test_coverage_sweep_config4 holds the table of all three modes. */
#define AQE_CI_GUARD 0.8
typedef enum aqe_ci_mode {
    AQE_CI_DEFAULT = 0,        /* = AQE_CI_STEIN_GUARDED */
    AQE_CI_PLAIN = 1,          /* z * s_r / sqrt(n_r) at the stopping look (the round-1 interval without its 1.05 factor) */
    AQE_CI_STEIN = 2,          /* t(df) * max(s_{r-1}, s_r) / sqrt(n_r) */
    AQE_CI_STEIN_GUARDED = 3   /* the same at alpha' = AQE_CI_GUARD * alpha */
} aqe_ci_mode;

/* Draw designs of the persistent CLT kernel (K4). */
typedef enum aqe_design {
    AQE_DESIGN_SRS = 0,    /* Philox simple random sampling with replacement over the shard */
    AQE_DESIGN_BLOCK = 1   /* Philox-chosen contiguous tiles of block_size rows (cluster sample) */
} aqe_design;

typedef struct aqe_approx_spec {
    int32_t  agg;                /* aqe_agg */
    int32_t  design;             /* aqe_design */
    int32_t  agg_col;            /* AQE_COL_AMOUNT (f64) */
    int32_t  pred_col;           /* AQE_COL_NONE or predicate column (ratio estimator) */
    double   lo, hi;
    double   error_percent;      /* stop when z*SE/|estimate| * 100 <= error_percent */
    double   confidence_level;   /* 0.90 / 0.95 / 0.99 or any in (0,1) */
    uint64_t seed;
    uint64_t min_samples;        /* first look; 0 = default */
    uint64_t max_samples;        /* budget; 0 = default (N) */
    uint32_t block_size;         /* AQE_DESIGN_BLOCK tile rows; 0 = 1000 (block_sample default) */
    uint32_t ci_mode;            /* aqe_ci_mode; 0 = default */
} aqe_approx_spec;

/* status values follow CustomApproximationStatus (custom_scheduler.hpp:8-13). */
typedef enum aqe_approx_status {
    AQE_STABLE = 0, AQE_DRIFTING = 1, AQE_INSUFFICIENT_DATA = 2, AQE_ERROR = 3
} aqe_approx_status;

/* Result fields = the reference's CustomValidationResult + QueryResult
 * (custom_scheduler.hpp:15-22, executor.h:8-12) with the correct SUM interval. */
typedef struct aqe_approx_result {
    double   estimate;
    double   ci_lower, ci_upper;
    double   error_margin;       /* achieved relative half width z*SE/|estimate| */
    double   confidence_level;
    uint64_t n_samples;          /* rows gathered */
    uint64_t n_units;            /* sampling units (rows for SRS, tiles for BLOCK) */
    uint64_t population;         /* N of the shard */
    double   mean, m2;           /* unit-level moments (merge across shards with Chan's formula) */
    uint32_t rounds;
    int32_t  status;             /* aqe_approx_status */
    double   elapsed_us;         /* device time of the persistent kernel (CUDA events) */
    double   pass_fraction;      /* estimated fraction of rows passing the predicate (1 without one): the weight of this shard's
                                    AVG ... WHERE estimate when independent shard results are merged (aqe_approx_merge) */
} aqe_approx_result;

/* Synthetic "sales-shaped" generator distributions. */
typedef enum aqe_synth { AQE_SYNTH_UNIFORM = 0, AQE_SYNTH_LOGNORMAL = 1 } aqe_synth;

/* ------------------------------------------------------------------------------------------------
 * Library
 * ---------------------------------------------------------------------------------------------- */
AQE_API int         aqe_abi_version(void);
AQE_API const char* aqe_last_error(void);          /* thread-local, never NULL */
AQE_API int         aqe_device_count(int* out);    /* AQE_ERR_CUDA when no driver / device */
/* Which exact-scan kernel the calling thread's last scan launched: name, template arguments, grid, shared memory (bench.py
 * reports it as roofline.kernel instead of a string literal).  Thread-local, never NULL. */
AQE_API const char* aqe_last_scan_kernel(void);
/* Number of this library's kernels launched since load (bench.py gpu_launches). */
AQE_API uint64_t    aqe_launch_count(void);
/* Page-locked host buffers for the host-column entry points (pageable memory also works, staged). */
AQE_API int         aqe_host_alloc(size_t bytes, void** out);
AQE_API int         aqe_host_free(void* p);

/* ------------------------------------------------------------------------------------------------
 * Lifecycle / ingest  (create_database :135, open_database :153, load_from_file :685,
 * save_to_file :665, insert_record :164, insert_batch :196, close_database :157)
 * ---------------------------------------------------------------------------------------------- */
/* Empty table bound to CUDA device `device` (no CUDA call is made until rows are needed on device). */
AQE_API int aqe_create(int device, aqe_db** out);
/* create + load_file of the whole file. */
AQE_API int aqe_open(const char* path, int device, aqe_db** out);
/* Replace contents with rows [first_row, first_row + n_rows) of the record file (n_rows = UINT64_MAX:
 * to the end).  Rows are (stably) ordered by id, as load_from_file's insert_batch does (:198-200). */
AQE_API int aqe_load_file(aqe_db* db, const char* path, uint64_t first_row, uint64_t n_rows);
/* Write header (total, height, count) + rows in ascending id (:665-683). */
AQE_API int aqe_save_file(aqe_db* db, const char* path);
/* Append host rows (insert_record / insert_batch); the device copy is rebuilt lazily. */
AQE_API int aqe_append_records(aqe_db* db, const aqe_record* rows, size_t n);
/* Replace contents with n host rows (AoS, pageable or pinned), chunked H2D + AoS->SoA on device. */
AQE_API int aqe_from_host_records(aqe_db* db, const aqe_record* rows, size_t n);
/* Borrow device-resident columns (Torch hand-off through data_ptr()); any pointer may be NULL if the
 * column is never queried.  The caller keeps ownership and must keep them alive. */
AQE_API int aqe_attach_device_columns(aqe_db* db, const int64_t* id, const double* amount, const int32_t* region,
                              const int32_t* product_id, const int64_t* timestamp, uint64_t n);
/* Fill the shard on the device with rows [first_row, first_row+n_rows) of the synthetic table
 * (Philox4x32-10, key = seed, counter = global row).  columns_mask bit c = materialise column c. */
AQE_API int aqe_generate_synthetic(aqe_db* db, uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist,
                           uint32_t columns_mask);
/* Host twin of the generator (same bits), for files / oracles.  rows[i] = global row first_row+i. */
AQE_API int aqe_synth_rows_host(uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist, aqe_record* rows);
AQE_API int aqe_close(aqe_db* db); /* frees device + host memory; db invalid afterwards */

/* ---- the whole table over several GPUs of THIS process (aqe_open(path, n_gpus, &db); 8e: "single process,
 * 8 devices") ----
 * The table is cut into contiguous row ranges [N*g/G, N*(g+1)/G) (the reference's own region split,
 * custom_bplus_db.cpp:925-926), one per device; G = min(n_devices, max(1, N / AQE_MIN_SHARD_ROWS)) so that small tables
 * stay on one GPU (environment AQE_MIN_SHARD_ROWS, default 2^24).  Each query launches one kernel per device from one host
 * thread per device; the 64-byte shard partials (or the SQL accumulators, or the per-look moments of the CLT kernel) are
 * exchanged INSIDE the kernels through peer-mapped mailboxes (cudaDeviceEnablePeerAccess; NVLink stores) and folded in
 * rank order, exactly as the one-process-per-GPU exchange does -- the same bits as aqe_merge_partials over the shards.
 * Where the devices are not distinct peers (tests put several shards on one GPU) the shards are merged on the host.
 * devices = NULL: devices 0 .. n_devices-1; n_devices = 0: every visible device. */
AQE_API int aqe_create_sharded(const int* devices, int n_devices, aqe_db** out);
AQE_API int aqe_open_sharded(const char* path, int n_gpus, aqe_db** out);   /* create_sharded(NULL, n_gpus) + load_file */
AQE_API int aqe_shard_count(const aqe_db* db);                /* shards in use for the current contents (1 for a plain handle) */
AQE_API aqe_db* aqe_shard(aqe_db* db, int g);                 /* borrowed handle of shard g (the handle itself for a plain one) */
AQE_API uint64_t aqe_shard_first_row(const aqe_db* db, int g);/* first table row of shard g; g = shard_count: N */
AQE_API int aqe_shards_fused(const aqe_db* db);               /* 1: the shards exchange inside the kernels (distinct peer devices) */

/* Host only: where the reference's B+ tree puts rows that SHARE an id.  ids[0, n) in arrival order; op_rows / op_kinds split them into the
 * calls that inserted them (kind 0: one insert_batch / load_from_file / insert_record -- std::sort by id :198-200, then per row a leaf insert in
 * front of equal keys :32-37, leaves splitting 127 / 128 :43-58; kind 1: rows already in table order; n_ops = 0: one batch of all rows).
 * perm[k] = arrival number of the row at position k of the table (collect_all_records :660).  The engine orders tables with duplicate ids
 * this way (up to 2^27 rows; larger ones keep equal ids in arrival order); tables without duplicates are simply ascending by id. */
AQE_API int aqe_reference_order(const int64_t* ids, uint64_t n, const uint64_t* op_rows, const int* op_kinds, size_t n_ops, uint64_t* perm);

AQE_API uint64_t aqe_count(const aqe_db* db);                 /* get_total_records :646 */
AQE_API uint64_t aqe_node_count(const aqe_db* db);            /* get_node_count :654 (N/255+1) */
AQE_API uint64_t aqe_tree_height(const aqe_db* db);           /* get_tree_height :650 (bulk-load shape) */
AQE_API int      aqe_device(const aqe_db* db);
/* Device pointer of a column (NULL if absent) -- zero-copy export to torch / cupy. */
AQE_API const void* aqe_column_device_ptr(aqe_db* db, int col);
/* Copy elements [first,first+n) of one column to host memory (pinned or pageable). */
AQE_API int aqe_read_column(aqe_db* db, int col, uint64_t first, uint64_t n, void* out);
/* Copy rows [first,first+n) back to host AoS (collect_all_records :660). */
AQE_API int aqe_read_records(aqe_db* db, uint64_t first, uint64_t n, aqe_record* out);

/* ------------------------------------------------------------------------------------------------
 * Exact full scans (K1/K2)  -- sum_amount :242, avg_amount :253, count_records :259,
 * sum_amount_where :263
 * ---------------------------------------------------------------------------------------------- */
AQE_API int aqe_scan(aqe_db* db, const aqe_scan_spec* spec, aqe_partial* out);
/* Asynchronous form: launches on `stream` (a cudaStream_t; 0 = the handle's own non-blocking stream -- to
 * target CUDA's default stream pass cudaStreamLegacy / cudaStreamPerThread explicitly) and leaves
 * the 64-byte partial in device memory at `partial_dev` (e.g. a torch tensor's data_ptr()). */
AQE_API int aqe_scan_async(aqe_db* db, const aqe_scan_spec* spec, void* partial_dev, void* stream);
/* Scan host-resident column data through the device: chunked, double-buffered H2D from `host_col`
 * (n elements of agg column type; pinned recommended) overlapped with the reduction.  This is the
 * end-to-end (host buffers in, scalar out) form of sum_amount / sum_amount_where. */
AQE_API int aqe_scan_host_column(int device, const void* host_col, int col_kind, uint64_t n, double lo, double hi,
                         int use_pred, aqe_partial* out);
/* The same through several devices of this process (one host thread per device): chunks are handed out from one counter, so
 * a device behind a slower link takes fewer of them, and merged in chunk order -- the result does not depend on which device
 * took which chunk and equals aqe_scan_host_column's bit for bit.  `devices`: n_devices distinct CUDA device numbers. */
AQE_API int aqe_scan_host_column_multi(const int* devices, int n_devices, const void* host_col, int col_kind, uint64_t n,
                               double lo, double hi, int use_pred, aqe_partial* out);
/* Fused cross-GPU exchange (one process per GPU on one NVLink/NVSwitch box).  aqe_exchange_init allocates this
 * rank's mailbox and returns its 64-byte CUDA IPC handle; after the ranks have all-gathered the handles
 * (any transport), aqe_exchange_connect maps every peer's mailbox.  aqe_scan_exchange[_async] then runs the
 * scan and, inside the same kernel, stores the shard's 64-byte partial into every rank's mailbox over NVLink,
 * waits for all ranks and folds them in rank order: the result is the TABLE-level partial on every rank,
 * bit-identical to aqe_merge_partials over the per-shard results.  All ranks must issue the same sequence
 * of exchange scans.  world <= 16. */
AQE_API int aqe_exchange_init(aqe_db* db, int rank, int world, void* ipc_handle_out /* 64 bytes */);
AQE_API int aqe_exchange_connect(aqe_db* db, const void* all_handles /* world x 64 bytes, rank order */);
AQE_API int aqe_scan_exchange(aqe_db* db, const aqe_scan_spec* spec, aqe_partial* out);
AQE_API int aqe_scan_exchange_async(aqe_db* db, const aqe_scan_spec* spec, void* merged_dev, void* stream);
/* After synchronising the stream of asynchronous exchange scans: AQE_ERR_CUDA if a peer never showed up. */
AQE_API int aqe_exchange_check(aqe_db* db);
/* Fixed-order merge of per-shard partials (rank order) -- pure host code. */
AQE_API int aqe_merge_partials(const aqe_partial* parts, int n, int is_integer, aqe_partial* out);
/* Conveniences over aqe_scan: */
AQE_API int aqe_sum_f64(aqe_db* db, int col, double* out);                                   /* :242 */
AQE_API int aqe_sum_where_f64(aqe_db* db, int col, double lo, double hi, double* sum, uint64_t* count); /* :263 */
AQE_API int aqe_sum_i128(aqe_db* db, int col, uint64_t* lo64, int64_t* hi64);                /* new */

/* ------------------------------------------------------------------------------------------------
 * Sample plans (Appendix A index generators) -- host arithmetic only, no device needed
 * ---------------------------------------------------------------------------------------------- */
AQE_API void aqe_sample_params_default(aqe_sample_params* p, int method); /* pybind defaults */
/* Build the position list of `method` for a table of n_rows rows.  Methods that depend on data
 * (ADAPTIVE_BLOCK, STRATIFIED_BLOCK, CLT_VALIDATED_DUAL_POINTER) need `db` (may be NULL otherwise). */
AQE_API int  aqe_plan_build(aqe_db* db, uint64_t n_rows, int method, const aqe_sample_params* p, aqe_plan** out);
AQE_API int  aqe_plan_from_indices(const int64_t* idx, uint64_t n, aqe_plan** out);
AQE_API uint64_t aqe_plan_count(const aqe_plan* plan);
AQE_API uint32_t aqe_plan_num_segments(const aqe_plan* plan);  /* 0 => explicit index list */
AQE_API int  aqe_plan_segments(const aqe_plan* plan, aqe_segment* out, uint32_t cap);
AQE_API int  aqe_plan_indices(const aqe_plan* plan, int64_t* out, uint64_t cap); /* expand on host */
AQE_API int  aqe_plan_sorted_by_amount(const aqe_plan* plan);  /* 1: positions index the amount-sorted order */
AQE_API void aqe_plan_free(aqe_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Sampled aggregates (K3/K5/K6)
 * ---------------------------------------------------------------------------------------------- */
/* Moments of column `col` (as double) over the plan's positions. */
AQE_API int aqe_stats_from_plan(aqe_db* db, const aqe_plan* plan, int col, aqe_stats* out);
/* Same, but sampled rows failing lo <= pred_col <= hi contribute 0 (parallel_sum_where_sample :317-343). */
AQE_API int aqe_stats_from_plan_where(aqe_db* db, const aqe_plan* plan, int col, int pred_col, double lo, double hi,
                              aqe_stats* out);
/* Same for a caller-supplied host index list ("same sample index list" parity, E1/E2). */
AQE_API int aqe_stats_from_indices(aqe_db* db, const int64_t* idx, uint64_t n, int col, aqe_stats* out);
/* Range-sharded tables, one process per GPU (sharded.ShardedTable): this handle holds rows [window_first, window_first +
 * aqe_count(db)) of the table the plan was built for.  Every rank walks the same plan; positions outside its window belong
 * to another rank and are skipped, so the gathers split across the GPUs and no index list is cut up or exchanged.
 * aqe_stats_window leaves the shard's mergeable sums, aqe_stats_merge (pure host code, rank order) finishes them;
 * aqe_gather_window writes rows [k_first, k_first + k_count) of the plan that fall into the window into out[k - k_first]
 * and leaves the other slots untouched (the ranks' buffers, zero-filled beforehand, add up to the sample).
 * Replaces the per-thread region loops of parallel_pointer_sample :814-854, parallel_block_sample :1218-1271,
 * multithreaded_memory_stride_sample :1880-1960 across GPUs. */
AQE_API int aqe_stats_window(aqe_db* db, const aqe_plan* plan, int col, int pred_col, double lo, double hi, uint64_t window_first,
                             aqe_stats_partial* out);
AQE_API int aqe_stats_merge(const aqe_stats_partial* parts, int n, aqe_stats* out);
AQE_API int aqe_gather_window(aqe_db* db, const aqe_plan* plan, uint64_t window_first, uint64_t k_first, uint64_t k_count,
                              aqe_record* out, uint64_t* n_local);
/* Rows of the table the plan was built for (0: an explicit index list, validated against the table at use). */
AQE_API uint64_t aqe_plan_table_rows(const aqe_plan* plan);
/* Rows at the plan's positions, AoS, in plan order (the legacy list[Record] return path). */
AQE_API int aqe_gather_plan(aqe_db* db, const aqe_plan* plan, aqe_record* out, uint64_t cap);
AQE_API int aqe_gather_records(aqe_db* db, const int64_t* idx, uint64_t n, aqe_record* out);
/* fast_aggregated_memory_stride_sum :1962 -- raw (unscaled) sample sum, plus the sample count. */
AQE_API int aqe_fast_aggregated_sum(aqe_db* db, const aqe_sample_params* p, double* sum, uint64_t* n);

/* CLI estimators (enhanced_aqe_cli.py:188-200 random, 257-291 clt) from sample moments.
 * legacy_ci != 0 reproduces the reference's SUM interval (margin * N/n). */
AQE_API int aqe_estimate(const aqe_stats* s, uint64_t population, int agg, double z, int legacy_ci,
                 double* estimate, double* ci_lower, double* ci_upper);

/* ------------------------------------------------------------------------------------------------
 * Persistent CLT kernel (K4): Philox draws, Welford partials, in-kernel stop rule
 * ---------------------------------------------------------------------------------------------- */
AQE_API int aqe_approx(aqe_db* db, const aqe_approx_spec* spec, aqe_approx_result* out);
/* Multi-GPU form (after aqe_exchange_connect + aqe_exchange_set_total_rows on every rank; shards must be the
 * contiguous ranges [N*g/G, N*(g+1)/G)): shards are strata with proportional allocation and ONE global stop
 * rule -- after every look each rank's kernel publishes its cumulative moments to all ranks' mailboxes over
 * NVLink and every rank evaluates the same stratified estimate (sum_g U_g mean_g, Var = sum_g U_g^2 s_g^2/n_g),
 * so all ranks stop at the same look and return the identical table-level result.  Replaces the reference's
 * should_stop / current_mean atomics between its fast and slow threads (custom_bplus_db.cpp:901-904). */
AQE_API int aqe_exchange_set_total_rows(aqe_db* db, uint64_t total_rows);
AQE_API int aqe_approx_exchange(aqe_db* db, const aqe_approx_spec* spec, aqe_approx_result* out);
/* Merge per-shard results (stratified by shard, rank order) into the table-level estimate. */
AQE_API int aqe_approx_merge(const aqe_approx_result* parts, int n, int agg, double confidence_level,
                     aqe_approx_result* out);
/* z for a two-sided confidence level: the reference's table 2.576/1.96/1.645 (:911-912) when
 * exact == 0, else the inverse normal CDF. */
AQE_API double aqe_z_score(double confidence_level, int exact);

/* ------------------------------------------------------------------------------------------------
 * SQL-string path on the columnar table: run_query / run_query_groupby /
 * run_query_with_ci / run_query_groupby_with_ci  (bindings.cpp:126-136 -> executor.cpp:28-338,
 * parser.cpp:20-75).  The reference turns `SELECT agg(col) FROM t [WHERE ...] [GROUP BY g]` into SQLite
 * statements over a SQLite file, sampling with `rowid % (100/p) = 0` and scaling SUM/COUNT by 100/p.
 * Here the same query runs as ONE grouped-scan kernel (k_sql_agg) over the HBM-resident columns:
 * rowid = id; WHERE = AND / OR / parentheses over comparisons, BETWEENs and IN lists of columns with numeric literals,
 * compiled to at most AQE_SQL_MAX_ALT OR-ed conjunctions of one closed interval (+ optional "!=" value) per column; GROUP BY on an integer column with a dense
 * key range of at most AQE_SQL_MAX_GROUPS values.  Sums are accumulated in 128-bit fixed point
 * (order-independent, so results are bit-reproducible and shard merges are exact).
 * ---------------------------------------------------------------------------------------------- */
#define AQE_SQL_MAX_GROUPS 4096

/* One conjunct per column after merging:  lo <= col <= hi  [and col != ne]. */
typedef struct aqe_sql_term {
    int32_t col;          /* aqe_column */
    int32_t has_ne;       /* 1: also requires col != ne / ine */
    double  lo, hi;       /* bounds when col is f64 */
    int64_t ilo, ihi;     /* bounds when col is an integer column (ilo > ihi: empty) */
    double  ne;
    int64_t ine;
} aqe_sql_term;

/* Parsed + compiled query: the reference's `struct Query` (parser.h:17-24) with names resolved.  The WHERE clause
 * is in disjunctive normal form: up to AQE_SQL_MAX_ALT conjunctions OR-ed together, each one closed interval
 * (+ optional "!=" value) per column. */
#define AQE_SQL_MAX_ALT 8
typedef struct aqe_sql_query {
    int32_t agg;              /* aqe_agg */
    int32_t agg_col;          /* aqe_column; AQE_COL_NONE for COUNT(*) */
    int32_t group_col;        /* aqe_column or AQE_COL_NONE */
    int32_t sample_percent;   /* as passed; step = 100 / p for 0 < p < 100 (executor.cpp:20-26) */
    int32_t n_alt;            /* 0: every row passes; else the number of OR-ed conjunctions */
    int32_t always_false;     /* WHERE is unsatisfiable */
    int32_t top_level_or;     /* the clause text has an OR outside all parentheses.  The reference builds its statements by
                                 pasting text around the clause -- `group = 'k' AND <where> AND rowid % step = 0`
                                 (executor.cpp:38-42, :95-98) -- so with a top-level OR the group filter binds to the first
                                 branch only and the sampling filter to the last.  Such queries are refused
                                 (AQE_ERR_UNSUPPORTED) for sampled or grouped calls; write `WHERE (a OR b)`. */
    int32_t _pad;
    int32_t n_terms[AQE_SQL_MAX_ALT];      /* terms per conjunction, 1..5 */
    aqe_sql_term terms[AQE_SQL_MAX_ALT][5];
    char agg_text[32], column[64], table[64], group_by[64];
    char where[512];          /* raw WHERE text as parser.cpp:36-51 extracts it */
} aqe_sql_query;

/* How the *_with_ci results are formed. */
typedef enum aqe_sql_mode {
    AQE_SQL_VALUE = 0,        /* run_query / run_query_groupby: value only (ci_lower = ci_upper = value) */
    AQE_SQL_CI_REFERENCE = 1, /* run_query_with_ci / run_query_groupby_with_ci exactly as executor.cpp:177-338
                                 (SUM reports mean*100/p, margin 1.96*SE*100/p) */
    AQE_SQL_CI_CORRECT = 2    /* additive: SUM = sample sum * 100/p with margin 1.96 * sqrt(n) * s * 100/p */
} aqe_sql_mode;

typedef struct aqe_sql_row {
    int64_t  key;             /* group key (0 when the query has no GROUP BY) */
    double   value, ci_lower, ci_upper;
    uint64_t count;           /* sampled rows that passed WHERE (unscaled) */
    double   sum, sumsq;      /* their sum / sum of squares (unscaled) */
    uint64_t isum_lo;         /* exact sum of an integer aggregate column, two's complement 128 bit */
    int64_t  isum_hi;
    int32_t  is_null;         /* 1: SUM/AVG over no rows -- SQLite yields NULL and the reference's std::stod throws */
    int32_t  _pad;
} aqe_sql_row;

/* Host-only: parse + compile (no device needed).  AQE_ERR_INVALID for what parser.cpp rejects
 * (std::runtime_error there), AQE_ERR_UNSUPPORTED for SQL that SQLite would accept but this engine
 * does not (OR, expressions, GROUP BY on amount, ...). */
AQE_API int aqe_sql_parse(const char* sql, int sample_percent, aqe_sql_query* out);
/* Execute on one shard.  rows[0..*n_rows) in ascending key order; *n_rows may exceed cap (then only cap
 * rows were written).  Without GROUP BY exactly one row comes back. */
AQE_API int aqe_sql_execute(aqe_db* db, const aqe_sql_query* q, int mode, aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows);
/* parse + execute. */
AQE_API int aqe_sql_run(aqe_db* db, const char* sql, int sample_percent, int mode, aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows);
/* ---- pieces of aqe_sql_execute, exposed so that shards of several GPUs can be merged exactly ----
 * The kernel leaves, per group, five 64-bit words {count, sum_lo, sum_hi, sq_lo, sq_hi}: (sum_hi:sum_lo) is the
 * two's-complement 128-bit sum of round(x * 2^sum_shift) (x itself for integer columns, shift 0), likewise the
 * squares.  Integer adds commute, so summing these words over shards (with carry) gives the table-level
 * accumulators bit-for-bit whatever the shard count; all shards must use the same key range and shifts. */
typedef struct aqe_sql_facts {    /* what a shard knows about the columns a query touches */
    int64_t key_min, key_max;     /* GROUP BY column range on this shard (0,0 without GROUP BY; min > max: no rows) */
    double  agg_absmax;           /* max |x| of the aggregate column on this shard */
    int32_t agg_is_integer;
    int32_t _pad;
} aqe_sql_facts;
typedef struct aqe_sql_layout {   /* agreed by all shards before the scan */
    int64_t  key_min;
    uint32_t n_groups;            /* key_max - key_min + 1  (1 without GROUP BY), <= AQE_SQL_MAX_GROUPS */
    int32_t  sum_shift, sq_shift;
    int32_t  is_integer;
} aqe_sql_layout;
#define AQE_SQL_MOMENTS   1       /* also accumulate squares (the *_with_ci forms) */
#define AQE_SQL_UNSAMPLED 2       /* ignore sample_percent and count only: which groups pass WHERE at all
                                     (executor.cpp:68-79 lists groups with SELECT DISTINCT, unsampled) */
AQE_API int aqe_sql_facts_of(aqe_db* db, const aqe_sql_query* q, aqe_sql_facts* out);
/* Host only: fixed-point shifts for values of magnitude <= agg_absmax (|x| * 2^sum_shift < 2^62). */
AQE_API int aqe_sql_shifts(double agg_absmax, int agg_is_integer, int* sum_shift, int* sq_shift);
/* Host only: merge the facts of all shards into the common layout.  The scale is set for min(largest agg_absmax, the bound the
 * WHERE clause puts on the aggregate column when every OR branch has one): `amount BETWEEN 0 AND 1e-5` keeps 62 bits below 1e-5
 * whatever else the column holds.  Every shard must call it with the same query and the same facts. */
AQE_API int aqe_sql_layout_of(const aqe_sql_query* q, const aqe_sql_facts* facts, int n_shards, aqe_sql_layout* out);
AQE_API int aqe_sql_scan(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* layout, int flags,
                         uint64_t* acc /* n_groups x 5 words, host */);
/* Multi-GPU form (after aqe_exchange_connect on every rank; all ranks issue the same sequence of calls with the same
 * layout): the scan kernel's last CTA stores this shard's accumulators into every rank's mailbox over NVLink, waits for
 * all ranks and adds them with 128-bit carries inside the same kernel -- `acc` receives the TABLE-level accumulators on
 * every rank, bit-identical to aqe_sql_merge over the per-shard results.  Replaces the all-gather + host merge. */
AQE_API int aqe_sql_scan_exchange(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* layout, int flags,
                                  uint64_t* acc /* n_groups x 5 words, host */);
/* Host only: acc[i] += other[i] over the 5-word groups with 128-bit carries. */
AQE_API int aqe_sql_merge(uint64_t* acc, const uint64_t* other, uint32_t n_groups);
/* Host only: the arithmetic of executor.cpp on merged accumulators.  `exists` (n_groups x 5 words from an
 * AQE_SQL_UNSAMPLED scan, or NULL = "every group with sampled rows, plus nothing else") decides which groups
 * are reported. */
AQE_API int aqe_sql_finish(const aqe_sql_query* q, int mode, const aqe_sql_layout* layout, const uint64_t* acc,
                           const uint64_t* exists, aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows);

#ifdef __cplusplus
}
#endif
#endif /* AQE_B200_H */
