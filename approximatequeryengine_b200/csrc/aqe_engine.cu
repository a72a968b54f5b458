// aqe_engine.cu -- host side of libaqe_b200.so: the C-ABI of include/aqe_b200.h over the kernels in
// aqe_kernels.cuh.  One handle = one shard of the record table as five HBM-resident columns on one GPU.
//
// What it replaces in the reference (src/aqe_backend/core/custom_bplus_db.cpp): the B+ tree
// (insert_record :164, insert_batch :196), the "mmap" cache cached_records_ (:186-191), the per-query
// O(N) copies collect_all_records :660 / collect_leaf_records :715, the serial scans sum_amount :242 /
// sum_amount_where :263 and the std::async sampler fan-outs.  No CPU fallback exists: without a CUDA
// device every data-path entry point returns AQE_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cub/device/device_radix_sort.cuh>

#include "aqe_b200.h"
#include "aqe_kernels.cuh"
#include "aqe_order.hpp"
#include "aqe_plan.hpp"
#include "aqe_sql.hpp"
#include "aqe_sql_kernels.cuh"

using namespace aqe;

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local std::string g_scan_kernel;   // what the last exact scan of this thread launched (aqe_last_scan_kernel)
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return e == cudaErrorMemoryAllocation ? AQE_ERR_NOMEM : AQE_ERR_CUDA;
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct Slot {  // mapped pinned result area: kernels write it, the host reads it after a stream sync
    aqe_partial partial;
    aqe_stats stats;
    aqe_approx_result approx;
    unsigned int flags[4];
    aqe_stats_partial stats_raw;       // k_plan_stats over a window of a sharded table
    unsigned long long gathered;       // k_plan_gather: rows written (window form)
};

struct Group;  // aqe_group.inl: a table range-sharded over several GPUs of this process

struct aqe_db {
    Group* group = nullptr;   // set on the handle aqe_create_sharded returns; its shards are plain handles
    int device = 0;
    bool cuda_ready = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;

    uint64_t n = 0;  // rows resident on the device
    MutColumns col{nullptr, nullptr, nullptr, nullptr, nullptr};
    bool owned = true;

    // host-side rows appended through insert_record / insert_batch and not yet on the device
    std::vector<aqe_record> host_rows;
    std::vector<aqe::OrderOp> host_ops;  // how host_rows came about, call by call (aqe_order.hpp); empty: one insert_batch of all of them
    bool host_authoritative = false;  // host_rows is the whole table and the device copy is stale

    // scratch
    ScanAcc* scan_partials = nullptr;
    StatAcc* stat_partials = nullptr;
    ApproxAcc* approx_slots = nullptr;
    unsigned int* tickets = nullptr;  // [4]
    Slot* slot_host = nullptr;
    Slot* slot_dev = nullptr;
    aqe_record* gather_buf = nullptr; uint64_t gather_cap = 0;
    void* plan_buf = nullptr; size_t plan_cap = 0;
    int64_t* amount_perm = nullptr;  // rows ordered by amount (stratified_block_sample)
    int max_grid = 0;

    // fused cross-GPU exchange (see Exchange in aqe_kernels.cuh)
    int ex_rank = 0, ex_world = 0;
    ExSlot* ex_mailbox = nullptr;
    aqe_partial* ex_local = nullptr;     // this shard's partial on its way from the scan kernel to k_scan_merge
    ExSlot* ex_peers[kMaxRanks] = {nullptr};
    unsigned long long ex_seq = 0;
    unsigned long long ax_msg = 0;       // next message index of the sampled-estimate exchange (lock-step on all ranks)
    unsigned long long sqlx_seq = 0;     // sequence number of the SQL-path exchange (lock-step on all ranks)
    unsigned long long* sql_local = nullptr;  // [AQE_SQL_MAX_GROUPS][5]: this shard's accumulators on their way into the exchange
    uint64_t ex_total_rows = 0;          // rows of the whole table (all shards)
    bool ex_connected = false;
    bool ex_ipc = false;                 // peers were opened as CUDA IPC handles (another process' memory) and must be closed as such

    // SQL-string path (aqe_sql_*): lazily computed column statistics + accumulators of the grouped scan
    struct ColStat { bool valid = false; unsigned long long min_key = 0, max_key = 0; bool dense = false; long long first_id = 0; };
    ColStat col_stat[5];
    unsigned long long* sql_acc = nullptr;       // [n_groups][5], device, all zero between launches
    unsigned long long* sql_out_host = nullptr;  // same shape, mapped pinned: the last CTA writes it
    unsigned long long* sql_out_dev = nullptr;
    unsigned long long* sql_stat_dev = nullptr;  // [3] min key, max key, not-dense flag
    unsigned int* sql_ticket = nullptr;
};

static const int kMaxGrid = 148 * 16;

// ---- a table sharded over several GPUs of this process (aqe_group.inl, included at the end of this file) ----
static int group_close(aqe_db* db);
static int group_ensure_device(aqe_db* db);
static int group_load_file(aqe_db* db, const char* path, uint64_t first_row, uint64_t n_rows);
static int group_from_host_records(aqe_db* db, const aqe_record* rows, size_t n);
static int group_generate(aqe_db* db, uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist, uint32_t mask);
static int group_read_records(aqe_db* db, uint64_t first, uint64_t n, aqe_record* out);
static int group_read_column(aqe_db* db, int col, uint64_t first, uint64_t n, void* out);
static int group_scan_range(aqe_db* db, const aqe_scan_spec* spec, uint64_t first, uint64_t n, bool moments, aqe_partial* out);
static int group_stats(aqe_db* db, const aqe_plan* pl, int col, int pred_col, double lo, double hi, aqe_stats* out);
static int group_gather(aqe_db* db, const aqe_plan* pl, aqe_record* out, uint64_t cap);
static int group_approx(aqe_db* db, const aqe_approx_spec* S, aqe_approx_result* out);
static int group_sql_facts(aqe_db* db, const aqe_sql_query* q, aqe_sql_facts* out);
static int group_sql_scan(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* L, int flags, uint64_t* acc);
static int group_amount_view(aqe_db* db, aqe::GlobalF64* out);
static int group_amount_perm(aqe_db* db, const int64_t** perm);
static int group_unsupported(const char* what) {
    return fail(AQE_ERR_UNSUPPORTED, std::string(what) + " works on one shard: use aqe_shard(db, g) of a sharded table");
}

static int db_init_cuda(aqe_db* db) {
    if (db->cuda_ready) { CU(cudaSetDevice(db->device)); return AQE_OK; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (db->device < 0 || db->device >= ndev) return fail(AQE_ERR_CUDA, "no such CUDA device " + std::to_string(db->device));
    CU(cudaSetDevice(db->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, db->device));
    if (prop.major < 10) return fail(AQE_ERR_CUDA, std::string("libaqe_b200 is built for sm_100a (B200); found ") + prop.name);
    db->sm_count = prop.multiProcessorCount;
    db->max_grid = kMaxGrid;
    CU(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&db->ev0));
    CU(cudaEventCreate(&db->ev1));
    CU(cudaMalloc(&db->scan_partials, sizeof(ScanAcc) * kMaxGrid));
    CU(cudaMalloc(&db->stat_partials, sizeof(StatAcc) * kMaxGrid));
    CU(cudaMalloc(&db->approx_slots, sizeof(ApproxAcc) * 2 * kMaxGrid));
    CU(cudaMalloc(&db->tickets, sizeof(unsigned int) * 4));
    CU(cudaMemset(db->tickets, 0, sizeof(unsigned int) * 4));
    CU(cudaMalloc(&db->ex_local, sizeof(aqe_partial)));
    CU(cudaHostAlloc(&db->slot_host, sizeof(Slot), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&db->slot_dev, db->slot_host, 0));
    db->cuda_ready = true;
    return AQE_OK;
}

static void free_columns(aqe_db* db) {
    if (db->owned) {
        cudaFree(db->col.id); cudaFree(db->col.amount); cudaFree(db->col.region); cudaFree(db->col.product_id); cudaFree(db->col.timestamp);
    }
    db->col = MutColumns{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (db->amount_perm) { cudaFree(db->amount_perm); db->amount_perm = nullptr; }
    for (auto& st : db->col_stat) st.valid = false;
    db->n = 0; db->owned = true;
}

static int alloc_columns(aqe_db* db, uint64_t n, uint32_t mask) {
    free_columns(db);
    const size_t m = n ? n : 1;
    if (mask & (1u << AQE_COL_ID)) CU(cudaMalloc(&db->col.id, m * 8));
    if (mask & (1u << AQE_COL_AMOUNT)) CU(cudaMalloc(&db->col.amount, m * 8));
    if (mask & (1u << AQE_COL_REGION)) CU(cudaMalloc(&db->col.region, m * 4));
    if (mask & (1u << AQE_COL_PRODUCT_ID)) CU(cudaMalloc(&db->col.product_id, m * 4));
    if (mask & (1u << AQE_COL_TIMESTAMP)) CU(cudaMalloc(&db->col.timestamp, m * 8));
    db->n = n; db->owned = true;
    return AQE_OK;
}

static Columns const_cols(const aqe_db* db) { return Columns{db->col.id, db->col.amount, db->col.region, db->col.product_id, db->col.timestamp}; }

static int grid_for(const aqe_db* db, uint64_t work_items, int per_thread, int threads, int blocks_per_sm) {
    uint64_t want = (work_items + (uint64_t)threads * per_thread - 1) / ((uint64_t)threads * per_thread);
    uint64_t cap = (uint64_t)db->sm_count * blocks_per_sm;
    if (cap > (uint64_t)db->max_grid) cap = db->max_grid;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------------
// ingest: host AoS rows -> device columns (K7), chunked through pinned staging
// ------------------------------------------------------------------------------------------------
// Replaces load_from_file's per-record ifstream::read + tree rebuild (custom_bplus_db.cpp:700-710).  W worker threads
// each run an independent pipeline over the chunks c = w, w+W, ... of the row range: fill a pinned buffer (pread from
// the file / memcpy from caller memory) -> cudaMemcpyAsync -> k_aos_to_soa on the worker's own stream, two buffers
// per worker so the next fill overlaps the previous copy.  Page-cache reads run at a few GB/s per thread, the H2D
// link at ~55 GB/s, so several readers are needed to approach the link (8 readers: 44 GB/s from the page cache with 4 MiB chunks; copying out of
// an mmap of the file instead of pread() measured slower, 25 GB/s).  Out-of-order ids are detected inside a
// chunk on the device and across chunk boundaries on the host.
// rows per ingest chunk (AQE_INGEST_CHUNK_MB, read once: the pooled staging buffers have this size).  4 MiB: a 3.2 GB file from the
// page cache loads at 44 GB/s with 8 readers (16 MiB chunks: 32 GB/s, 8: 38, 2: 29-44) and a first load in a fresh process allocates
// 64 MB instead of 256 MB of pinned staging (profiles/r2_ingest_cold.json)
static size_t ingest_chunk_rows() {
    static const size_t rows = (size_t)std::min(64, std::max(1, env_int("AQE_INGEST_CHUNK_MB", 4))) << 15;
    return rows;
}

struct IngestWorker {
    bool ready = false;   // staging allocated (by the worker's own thread, so that the workers of a first load allocate side by side)
    aqe_record* pinned[2] = {nullptr, nullptr};
    aqe_record* dev[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    int rc = AQE_OK;
    std::string err;
    int init() {
        if (ready) return AQE_OK;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaHostAlloc(&pinned[i], ingest_chunk_rows() * sizeof(aqe_record), cudaHostAllocDefault));
            CU(cudaMalloc(&dev[i], ingest_chunk_rows() * sizeof(aqe_record)));
            CU(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        ready = true;
        return AQE_OK;
    }
    ~IngestWorker() {
        for (int i = 0; i < 2; ++i) { if (pinned[i]) cudaFreeHost(pinned[i]); if (dev[i]) cudaFree(dev[i]); if (done[i]) cudaEventDestroy(done[i]); }
        if (stream) cudaStreamDestroy(stream);
    }
};

struct IngestPool {
    static std::mutex& mu() { static std::mutex m; return m; }
    static std::vector<IngestWorker*>& idle(int device) { static std::vector<IngestWorker*> pool[16]; return pool[device & 15]; }
    struct Lease {
        int device, rc = AQE_OK;
        std::vector<IngestWorker*> workers;
        Lease(int dev, int count) : device(dev) {
            {
                std::lock_guard<std::mutex> lock(mu());
                auto& p = idle(device);
                while ((int)workers.size() < count && !p.empty()) { workers.push_back(p.back()); p.pop_back(); }
            }
            while ((int)workers.size() < count) {
                IngestWorker* w = new (std::nothrow) IngestWorker();
                if (!w) { rc = fail(AQE_ERR_NOMEM, "out of host memory"); return; }
                workers.push_back(w);   // its staging is allocated by the thread that will use it (IngestWorker::init in ingest_rows)
            }
        }
        ~Lease() {
            std::lock_guard<std::mutex> lock(mu());
            auto& p = idle(device);
            for (IngestWorker* w : workers) p.push_back(w);
        }
    };
};

// Feeds `n` rows obtained chunk-wise from the thread-safe `fill(dst, first, count)` into the columns.
// *unsorted_out is set if ids were found out of order or repeated (not strictly ascending).
// max_workers > 0 caps the reader threads (a sharded table loads its shards side by side and divides the readers among them).
template <typename Fill>
static int ingest_rows(aqe_db* db, uint64_t n, Fill fill, bool* unsorted_out, int max_workers = 0) {
    int rc = alloc_columns(db, n, 0x1f);
    if (rc) return rc;
    *unsorted_out = false;
    if (n == 0) return AQE_OK;
    const uint64_t chunk_rows = ingest_chunk_rows();
    const uint64_t nchunks = (n + chunk_rows - 1) / chunk_rows;
    int W = env_int("AQE_INGEST_THREADS", 0);
    if (W <= 0) W = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    if (max_workers > 0) W = std::min(W, max_workers);
    W = (int)std::min<uint64_t>((uint64_t)W, nchunks);
    unsigned int* unsorted = nullptr;
    CU(cudaMalloc(&unsorted, 4));
    CU(cudaMemset(unsorted, 0, 4));
    // worker resources (2 x 16 MiB pinned + 2 x 16 MiB device + a stream each) are pooled per device: allocating them costs
    // ~15 ms per worker, a third of a 6.4 GB load when done per call
    IngestPool::Lease lease(db->device, W);
    if (lease.rc) { cudaFree(unsorted); return lease.rc; }
    std::vector<IngestWorker*>& workers = lease.workers;
    for (IngestWorker* w : workers) { w->rc = AQE_OK; w->err.clear(); }
    std::vector<int64_t> first_id, last_id;
    try { first_id.resize(nchunks); last_id.resize(nchunks); } catch (const std::bad_alloc&) { cudaFree(unsorted); return fail(AQE_ERR_NOMEM, "out of host memory"); }
    const int device = db->device;
    const MutColumns cols = db->col;
    const int sm_count = db->sm_count;
    auto body = [&](int wi) {
        IngestWorker& w = *workers[wi];
        if (cudaSetDevice(device) != cudaSuccess) { w.rc = AQE_ERR_CUDA; w.err = "cudaSetDevice failed in ingest worker"; return; }
        if (!w.ready && w.init() != AQE_OK) { w.rc = AQE_ERR_NOMEM; w.err = "ingest worker: cannot allocate staging buffers: " + g_err; return; }
        int buf = 0;
        for (uint64_t c = (uint64_t)wi; c < nchunks; c += (uint64_t)W, buf ^= 1) {
            const uint64_t first = c * chunk_rows, cnt = std::min<uint64_t>(chunk_rows, n - first);
            cudaError_t e = cudaEventSynchronize(w.done[buf]);  // the pinned buffer is free again
            if (e == cudaSuccess && !fill(w.pinned[buf], first, cnt)) { w.rc = AQE_ERR_IO; w.err = "short read while loading rows"; return; }
            first_id[c] = w.pinned[buf][0].id; last_id[c] = w.pinned[buf][cnt - 1].id;
            if (e == cudaSuccess) e = cudaMemcpyAsync(w.dev[buf], w.pinned[buf], cnt * sizeof(aqe_record), cudaMemcpyHostToDevice, w.stream);
            if (e == cudaSuccess) {
                const uint64_t want = (cnt + 1023) / 1024;
                const int grid = (int)std::min<uint64_t>(want, (uint64_t)sm_count * 8);
                k_aos_to_soa<<<grid, 256, 0, w.stream>>>(w.dev[buf], cnt, cols, first, 0, 0, unsorted);
                LAUNCHED();
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaEventRecord(w.done[buf], w.stream);
            if (e != cudaSuccess) { w.rc = AQE_ERR_CUDA; w.err = std::string("ingest worker: ") + cudaGetErrorString(e); return; }
        }
        if (cudaStreamSynchronize(w.stream) != cudaSuccess) { w.rc = AQE_ERR_CUDA; w.err = "ingest worker: stream sync failed"; }
    };
    std::vector<std::thread> threads;
    for (int wi = 1; wi < W; ++wi) threads.emplace_back(body, wi);
    body(0);
    for (auto& t : threads) t.join();
    CU(cudaSetDevice(db->device));
    // a worker that gave up (short read, CUDA error) may still have copies / kernels in flight on its stream: drain every stream
    // before the staging buffers go back to the pool and the flag buffer is freed
    bool any_failed = false;
    for (IngestWorker* w : workers) any_failed = any_failed || w->rc != AQE_OK;
    if (any_failed)
        for (IngestWorker* w : workers) cudaStreamSynchronize(w->stream);
    for (IngestWorker* w : workers)
        if (w->rc) { cudaFree(unsorted); cudaGetLastError(); return fail(w->rc, w->err); }
    unsigned int flag = 0;
    CU(cudaMemcpy(&flag, unsorted, 4, cudaMemcpyDeviceToHost));
    cudaFree(unsorted);
    bool bad = flag != 0;   // an id below (bit 0) or equal to (bit 1) its predecessor: either way the host decides the order
    for (uint64_t c = 1; c < nchunks && !bad; ++c) bad = first_id[c] <= last_id[c - 1];
    *unsorted_out = bad;
    return AQE_OK;
}

// rows[0, n) in arrival order -> `sorted` in the order the reference's table would hold them: ascending id (load_from_file ->
// insert_batch orders rows by id, custom_bplus_db.cpp:198-200); rows that share an id sit where the reference's B+ tree puts them
// after the same history of inserts (`ops`; none: one insert_batch of all rows) -- aqe_order.cpp.  Tables above
// kReferenceOrderMaxRows keep equal ids in arrival order.
static int order_like_reference(const aqe_record* rows, uint64_t n, const std::vector<aqe::OrderOp>& ops, std::vector<aqe_record>& sorted) {
    try {
        sorted.assign(rows, rows + n);
        std::stable_sort(sorted.begin(), sorted.end(), [](const aqe_record& a, const aqe_record& b) { return a.id < b.id; });
        bool dup = false;
        for (uint64_t i = 1; i < n && !dup; ++i) dup = sorted[i].id == sorted[i - 1].id;
        if (!dup || n > aqe::kReferenceOrderMaxRows) return AQE_OK;
        std::vector<int64_t> ids(n);
        for (uint64_t i = 0; i < n; ++i) ids[i] = rows[i].id;
        std::vector<uint64_t> perm(n);
        const aqe::OrderOp whole{n, aqe::ORDER_OP_BATCH};
        const bool ok = ops.empty() ? aqe::reference_order(ids.data(), n, &whole, 1, perm.data())
                                    : aqe::reference_order(ids.data(), n, ops.data(), ops.size(), perm.data());
        if (!ok) return fail(AQE_ERR_STATE, "insert history does not cover the rows on the host");
        for (uint64_t i = 0; i < n; ++i) sorted[i] = rows[perm[i]];
    } catch (const std::bad_alloc&) { return fail(AQE_ERR_NOMEM, "out of host memory while ordering rows by id"); }
    return AQE_OK;
}

static int upload_host_rows(aqe_db* db, const aqe_record* rows, uint64_t n, const std::vector<aqe::OrderOp>& ops = {}) {
    bool unsorted = false;
    int rc = ingest_rows(db, n, [&](aqe_record* dst, uint64_t first, uint64_t cnt) { std::memcpy(dst, rows + first, cnt * sizeof(aqe_record)); return true; }, &unsorted);
    if (rc) return rc;
    if (unsorted) {
        std::vector<aqe_record> sorted;
        if ((rc = order_like_reference(rows, n, ops, sorted))) return rc;
        rc = ingest_rows(db, n, [&](aqe_record* dst, uint64_t first, uint64_t cnt) { std::memcpy(dst, sorted.data() + first, cnt * sizeof(aqe_record)); return true; }, &unsorted);
    }
    return rc;
}

// Brings rows appended on the host (insert_record / insert_batch) onto the device.
static int ensure_device(aqe_db* db) {
    if (db->group) return group_ensure_device(db);
    int rc = db_init_cuda(db);
    if (rc) return rc;
    if (!db->host_authoritative) return AQE_OK;
    rc = upload_host_rows(db, db->host_rows.data(), db->host_rows.size(), db->host_ops);
    if (rc) return rc;
    db->host_authoritative = false;
    return AQE_OK;
}

static int col_kind_of(int col) {  // 0 f64, 1 i64, 2 i32
    switch (col) {
        case AQE_COL_AMOUNT: return 0;
        case AQE_COL_ID: case AQE_COL_TIMESTAMP: return 1;
        case AQE_COL_REGION: case AQE_COL_PRODUCT_ID: return 2;
        default: return -1;
    }
}

// ------------------------------------------------------------------------------------------------
// library-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" {

int aqe_abi_version(void) { return AQE_ABI_VERSION; }
const char* aqe_last_scan_kernel(void) { return g_scan_kernel.c_str(); }
const char* aqe_last_error(void) { return g_err.c_str(); }
uint64_t aqe_launch_count(void) { return g_launches.load(); }

int aqe_device_count(int* out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { if (out) *out = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    if (out) *out = n;
    return AQE_OK;
}

int aqe_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(AQE_ERR_INVALID, "NULL out");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return AQE_OK;
}
int aqe_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return AQE_OK;
}

int aqe_create(int device, aqe_db** out) {
    if (!out) return fail(AQE_ERR_INVALID, "out is NULL");
    aqe_db* db = new (std::nothrow) aqe_db();
    if (!db) return fail(AQE_ERR_NOMEM, "out of host memory");
    db->device = device;
    *out = db;
    return AQE_OK;
}

int aqe_close(aqe_db* db) {
    if (!db) return AQE_OK;
    if (db->group) group_close(db);
    if (db->cuda_ready) {
        cudaSetDevice(db->device);
        cudaStreamSynchronize(db->stream);
        free_columns(db);
        cudaFree(db->scan_partials); cudaFree(db->stat_partials); cudaFree(db->approx_slots); cudaFree(db->tickets);
        cudaFree(db->gather_buf); cudaFree(db->plan_buf);
        cudaFree(db->sql_acc); cudaFree(db->sql_stat_dev); cudaFree(db->sql_ticket); cudaFree(db->sql_local);
        if (db->sql_out_host) cudaFreeHost(db->sql_out_host);
        for (int r = 0; r < db->ex_world; ++r)
            if (db->ex_connected && db->ex_ipc && r != db->ex_rank && db->ex_peers[r]) cudaIpcCloseMemHandle(db->ex_peers[r]);
        cudaFree(db->ex_mailbox); cudaFree(db->ex_local);
        cudaFreeHost(db->slot_host);
        cudaEventDestroy(db->ev0); cudaEventDestroy(db->ev1);
        cudaStreamDestroy(db->stream);
    }
    delete db;
    return AQE_OK;
}

// The record file: header (total, height, count) + count 32-byte rows (custom_bplus_db.cpp:665-683).
struct RecordFile {
    int fd = -1;
    uint64_t total = 0;   // record_count; the first two header words are ignored (custom_bplus_db.cpp:692-698)
    ~RecordFile() { if (fd >= 0) ::close(fd); }
    bool read_at(void* dst, size_t bytes, uint64_t off) const {
        char* p = static_cast<char*>(dst);
        while (bytes) {
            const ssize_t k = ::pread(fd, p, bytes, (off_t)off);
            if (k <= 0) return false;
            p += k; off += (uint64_t)k; bytes -= (size_t)k;
        }
        return true;
    }
    bool read_rows(aqe_record* dst, uint64_t first, uint64_t cnt) const { return read_at(dst, cnt * sizeof(aqe_record), 24 + first * 32); }
    int open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(AQE_ERR_IO, std::string("cannot open ") + path);
        uint64_t hdr[3];
        if (!read_at(hdr, 24, 0)) return fail(AQE_ERR_IO, std::string("short header in ") + path);
        struct stat st;
        // the header's row count sizes every allocation below: it must fit the file
        if (::fstat(fd, &st) != 0 || (uint64_t)st.st_size < 24 || hdr[2] > ((uint64_t)st.st_size - 24) / 32)
            return fail(AQE_ERR_IO, std::string("record count in the header exceeds the file: ") + path);
        total = hdr[2];
        return AQE_OK;
    }
};

int aqe_load_file(aqe_db* db, const char* path, uint64_t first_row, uint64_t n_rows) {
    if (!db || !path) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_load_file(db, path, first_row, n_rows);
    RecordFile f;
    int rc = f.open(path);
    if (rc) return rc;
    rc = db_init_cuda(db);
    if (rc) return rc;
    if (first_row > f.total) first_row = f.total;
    const uint64_t n = std::min<uint64_t>(n_rows, f.total - first_row);
    db->host_rows.clear(); db->host_ops.clear(); db->host_authoritative = false;
    bool unsorted = false;
    rc = ingest_rows(db, n, [&](aqe_record* dst, uint64_t first, uint64_t cnt) { return f.read_rows(dst, first_row + first, cnt); }, &unsorted);
    if (rc == AQE_OK && unsorted) {
        std::vector<aqe_record> rows;
        try { rows.resize(n); } catch (const std::bad_alloc&) { return fail(AQE_ERR_NOMEM, "out of host memory while ordering rows by id"); }
        if (!f.read_rows(rows.data(), first_row, n)) rc = fail(AQE_ERR_IO, "re-read failed");
        else rc = upload_host_rows(db, rows.data(), n);
    }
    return rc;
}

int aqe_open(const char* path, int device, aqe_db** out) {
    int rc = aqe_create(device, out);
    if (rc) return rc;
    rc = aqe_load_file(*out, path, 0, UINT64_MAX);
    if (rc) { std::string keep = g_err; aqe_close(*out); *out = nullptr; g_err = keep; }
    return rc;
}

int aqe_append_records(aqe_db* db, const aqe_record* rows, size_t n) {
    if (!db || (!rows && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (!db->host_authoritative) {
        // first append after a load: pull the table back so that host_rows is the whole table (rows that already are in table order)
        if (db->host_rows.size() != db->n || db->host_ops.empty()) {
            db->host_rows.clear(); db->host_ops.clear();
            if (db->n) {
                db->host_rows.resize(db->n);
                int rc = aqe_read_records(db, 0, db->n, db->host_rows.data());
                if (rc) return rc;
                db->host_ops.push_back(aqe::OrderOp{db->n, aqe::ORDER_OP_RESTORE});
            }
        }   // else: host_rows still holds the rows of the device table in arrival order, host_ops their history
        db->host_authoritative = true;
    }
    db->host_rows.insert(db->host_rows.end(), rows, rows + n);
    if (n) db->host_ops.push_back(aqe::OrderOp{(uint64_t)n, aqe::ORDER_OP_BATCH});   // one insert_batch, or one insert_record (a batch of one)
    return AQE_OK;
}

int aqe_from_host_records(aqe_db* db, const aqe_record* rows, size_t n) {
    if (!db || (!rows && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_from_host_records(db, rows, n);
    int rc = db_init_cuda(db);
    if (rc) return rc;
    db->host_rows.clear(); db->host_ops.clear(); db->host_authoritative = false;
    return upload_host_rows(db, rows, n);
}

int aqe_attach_device_columns(aqe_db* db, const int64_t* id, const double* amount, const int32_t* region,
                              const int32_t* product_id, const int64_t* timestamp, uint64_t n) {
    if (!db) return fail(AQE_ERR_INVALID, "NULL handle");
    if (db->group) return group_unsupported("aqe_attach_device_columns");
    int rc = db_init_cuda(db);
    if (rc) return rc;
    free_columns(db);
    db->host_rows.clear(); db->host_ops.clear(); db->host_authoritative = false;
    db->col = MutColumns{const_cast<int64_t*>(id), const_cast<double*>(amount), const_cast<int32_t*>(region),
                         const_cast<int32_t*>(product_id), const_cast<int64_t*>(timestamp)};
    db->owned = false;
    db->n = n;
    return AQE_OK;
}

int aqe_generate_synthetic(aqe_db* db, uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist, uint32_t columns_mask) {
    if (!db) return fail(AQE_ERR_INVALID, "NULL handle");
    if (db->group) return group_generate(db, seed, first_row, n_rows, dist, columns_mask ? columns_mask : 0x1f);
    int rc = db_init_cuda(db);
    if (rc) return rc;
    db->host_rows.clear(); db->host_ops.clear(); db->host_authoritative = false;
    rc = alloc_columns(db, n_rows, columns_mask ? columns_mask : 0x1f);
    if (rc) return rc;
    if (n_rows == 0) return AQE_OK;
    const int grid = grid_for(db, n_rows, 4, 256, 8);
    k_synth<<<grid, 256, 0, db->stream>>>(seed, first_row, n_rows, dist, db->col);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(db->stream));
    return AQE_OK;
}

int aqe_synth_rows_host(uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist, aqe_record* rows) {
    if (!rows && n_rows) return fail(AQE_ERR_INVALID, "NULL rows");
    for (uint64_t i = 0; i < n_rows; ++i) synth_row(seed, first_row + i, dist, rows[i]);
    return AQE_OK;
}

int aqe_reference_order(const int64_t* ids, uint64_t n, const uint64_t* op_rows, const int* op_kinds, size_t n_ops, uint64_t* perm) {
    if ((!ids && n) || (!perm && n) || (n_ops && (!op_rows || !op_kinds))) return fail(AQE_ERR_INVALID, "NULL argument");
    try {
        std::vector<aqe::OrderOp> ops(n_ops);
        for (size_t i = 0; i < n_ops; ++i) {
            if (op_kinds[i] != aqe::ORDER_OP_BATCH && op_kinds[i] != aqe::ORDER_OP_RESTORE) return fail(AQE_ERR_INVALID, "bad op kind");
            ops[i] = aqe::OrderOp{op_rows[i], op_kinds[i]};
        }
        if (ops.empty()) ops.push_back(aqe::OrderOp{n, aqe::ORDER_OP_BATCH});
        if (!aqe::reference_order(ids, n, ops.data(), ops.size(), perm)) return fail(AQE_ERR_INVALID, "ops do not cover the rows, or too many rows to replay");
    } catch (const std::bad_alloc&) { return fail(AQE_ERR_NOMEM, "out of host memory"); }
    return AQE_OK;
}

uint64_t aqe_count(const aqe_db* db) { return db ? (db->host_authoritative ? db->host_rows.size() : db->n) : 0; }
uint64_t aqe_node_count(const aqe_db* db) { return aqe_count(db) / 255 + 1; }  // custom_bplus_db.cpp:654-658
uint64_t aqe_tree_height(const aqe_db* db) { return tree_height(aqe_count(db)); }
int aqe_device(const aqe_db* db) { return db ? db->device : -1; }

const void* aqe_column_device_ptr(aqe_db* db, int col) {
    if (!db || ensure_device(db)) return nullptr;
    if (db->group) { group_unsupported("aqe_column_device_ptr"); return nullptr; }
    switch (col) {
        case AQE_COL_ID: return db->col.id;
        case AQE_COL_AMOUNT: return db->col.amount;
        case AQE_COL_REGION: return db->col.region;
        case AQE_COL_PRODUCT_ID: return db->col.product_id;
        case AQE_COL_TIMESTAMP: return db->col.timestamp;
        default: return nullptr;
    }
}

static int ensure_gather_buf(aqe_db* db, uint64_t rows) {
    if (rows <= db->gather_cap) return AQE_OK;
    if (db->gather_buf) cudaFree(db->gather_buf);
    db->gather_buf = nullptr; db->gather_cap = 0;
    const uint64_t cap = std::max<uint64_t>(rows, 1u << 16);
    CU(cudaMalloc(&db->gather_buf, cap * sizeof(aqe_record)));
    db->gather_cap = cap;
    return AQE_OK;
}

int aqe_read_records(aqe_db* db, uint64_t first, uint64_t n, aqe_record* out) {
    if (!db || (!out && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->host_authoritative) {
        if (first + n > db->host_rows.size()) return fail(AQE_ERR_INVALID, "row range out of bounds");
        // rows are reported in ascending id, as collect_all_records does (custom_bplus_db.cpp:660)
        int rc = ensure_device(db);
        if (rc) return rc;
    }
    if (db->group) return group_read_records(db, first, n, out);
    int rc = db_init_cuda(db);
    if (rc) return rc;
    if (first + n > db->n) return fail(AQE_ERR_INVALID, "row range out of bounds");
    const uint64_t chunk = 1u << 22;
    rc = ensure_gather_buf(db, std::min<uint64_t>(n, chunk));
    if (rc) return rc;
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - off);
        k_soa_to_aos<<<grid_for(db, cnt, 4, 256, 8), 256, 0, db->stream>>>(const_cols(db), db->gather_buf, first + off, cnt);
        LAUNCHED();
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + off, db->gather_buf, cnt * sizeof(aqe_record), cudaMemcpyDeviceToHost, db->stream));
        CU(cudaStreamSynchronize(db->stream));
    }
    return AQE_OK;
}

int aqe_read_column(aqe_db* db, int col, uint64_t first, uint64_t n, void* out) {
    if (!db || (!out && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    if (db->group) return group_read_column(db, col, first, n, out);
    const int k = col_kind_of(col);
    const char* base = static_cast<const char*>(aqe_column_device_ptr(db, col));
    if (k < 0 || !base) return fail(AQE_ERR_STATE, "column is not resident on the device");
    if (first + n > db->n) return fail(AQE_ERR_INVALID, "row range out of bounds");
    const size_t esz = k == 2 ? 4 : 8;
    CU(cudaMemcpyAsync(out, base + first * esz, n * esz, cudaMemcpyDeviceToHost, db->stream));
    CU(cudaStreamSynchronize(db->stream));
    return AQE_OK;
}

int aqe_save_file(aqe_db* db, const char* path) {
    if (!db || !path) return fail(AQE_ERR_INVALID, "NULL argument");
    const uint64_t n = aqe_count(db);
    // rows stream through a bounded buffer (k_soa_to_aos -> D2H -> fwrite per 4 M rows): a 1 B-row table never needs its
    // 32 GB on the host at once
    const uint64_t chunk = 1u << 22;
    std::vector<aqe_record> buf;
    try { buf.resize((size_t)std::min<uint64_t>(n, chunk)); } catch (const std::bad_alloc&) { return fail(AQE_ERR_NOMEM, "out of host memory"); }
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(AQE_ERR_IO, std::string("cannot create ") + path);
    const uint64_t hdr[3] = {n, tree_height(n), n};  // custom_bplus_db.cpp:669-676
    bool ok = std::fwrite(hdr, 8, 3, f) == 3;
    int rc = AQE_OK;
    for (uint64_t off = 0; ok && rc == AQE_OK && off < n; off += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - off);
        rc = aqe_read_records(db, off, cnt, buf.data());
        if (rc == AQE_OK) ok = std::fwrite(buf.data(), sizeof(aqe_record), cnt, f) == cnt;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (rc) return rc;
    return ok ? AQE_OK : fail(AQE_ERR_IO, std::string("write failed: ") + path);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// exact scans
// ------------------------------------------------------------------------------------------------
enum ColKind { K_F64 = 0, K_I64 = 1, K_I32 = 2 };
static int col_kind(int col) {
    switch (col) {
        case AQE_COL_AMOUNT: return K_F64;
        case AQE_COL_ID: case AQE_COL_TIMESTAMP: return K_I64;
        case AQE_COL_REGION: case AQE_COL_PRODUCT_ID: return K_I32;
        default: return -1;
    }
}
static const void* col_ptr(const aqe_db* db, int col) {
    switch (col) {
        case AQE_COL_ID: return db->col.id;
        case AQE_COL_AMOUNT: return db->col.amount;
        case AQE_COL_REGION: return db->col.region;
        case AQE_COL_PRODUCT_ID: return db->col.product_id;
        case AQE_COL_TIMESTAMP: return db->col.timestamp;
        default: return nullptr;
    }
}
static size_t kind_size(int k) { return k == K_I32 ? 4 : 8; }

struct ScanTuning { int variant, bps, unroll, stages, chunk_kb, minb; };
static ScanTuning scan_tuning() {
    // Defaults = the configuration that won the round-1 sweep (profiles/r1_scan_sweep*.jsonl).  The knobs exist
    // for tools/scan_sweep.py; they are read per call so one process can sweep them.
    ScanTuning t = {env_int("AQE_SCAN_VARIANT", 0), env_int("AQE_SCAN_BPS", 0), env_int("AQE_SCAN_UNROLL", 4),
                    env_int("AQE_SCAN_STAGES", 4), env_int("AQE_SCAN_CHUNK_KB", 16), env_int("AQE_SCAN_MINB", 1)};
    return t;
}

// occupancy (and the one-time dynamic-smem opt-in) per kernel, looked up once
static int kernel_occupancy(const void* kernel, int threads, size_t smem) {
    struct Key { const void* k; int dev; size_t smem; int occ; };
    static std::mutex mu;
    static std::vector<Key> cache;
    int dev = 0;
    cudaGetDevice(&dev);  // function attributes are per device
    std::lock_guard<std::mutex> lock(mu);
    size_t opted = 0;
    for (auto& e : cache) if (e.k == kernel && e.dev == dev) { if (e.smem == smem) return e.occ; opted = std::max(opted, e.smem); }
    if (smem > opted) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    occ = std::max(occ, 1);
    cache.push_back(Key{kernel, dev, smem, occ});
    return occ;
}

template <typename T> static const char* type_name() { return std::is_same_v<T, double> ? "double" : (sizeof(T) == 8 ? "int64" : "int32"); }
template <typename K> static int launch_regs_kernel(const aqe_db* db, K kernel, const ScanArgs& a, int W, int U, int bps_req, cudaStream_t s) {
    const int occ = kernel_occupancy((const void*)kernel, kScanThreads, 0);
    const int bps = bps_req > 0 ? std::min(bps_req, occ) : occ;
    g_scan_kernel += " grid=" + std::to_string(grid_for(db, a.n / W, U, kScanThreads, bps)) + " x " + std::to_string(kScanThreads) + " threads";
    kernel<<<grid_for(db, a.n / W, U, kScanThreads, bps), kScanThreads, 0, s>>>(a);
    LAUNCHED();
    return AQE_OK;
}
template <typename K> static int launch_ring_kernel(const aqe_db* db, K kernel, const ScanArgs& a_in, int stages, int rows_per_tile, int bps_req, cudaStream_t s) {
    ScanArgs a = a_in;
    size_t smem = (size_t)stages * kStageBytes;
    // programmatic dependent launch: exactly two CTAs fit an SM (3 x 76 KiB > 227 KiB), see ScanArgs::pdl_tail
    if (a.pdl_tail && (bps_req == 0 || bps_req == 2)) smem = std::max<size_t>(smem, 76 * 1024);
    const int occ = kernel_occupancy((const void*)kernel, kBulkThreads, smem);
    const int bps = bps_req > 0 ? std::min(bps_req, occ) : occ;
    const uint64_t ntiles = (a.n + rows_per_tile - 1) / rows_per_tile;
    int grid = (int)std::min<uint64_t>((uint64_t)db->sm_count * bps, std::max<uint64_t>(ntiles, 1));
    if (grid > db->max_grid) grid = db->max_grid;
    // tile schedule (ScanArgs::even_rounds): plain round-robin, or -- two CTAs per SM, back-to-back scans -- the first half of the grid
    // takes `skew` more tiles than the second so that the two CTAs of an SM do not hand over to the next scan at the same time
    const uint64_t n_main_tiles = ((a.n & ~3ull) + rows_per_tile - 1) / rows_per_tile;
    a.even_rounds = ~0ull; a.long_ctas = (unsigned)grid;
    // 16 tiles apart (~10 us of streaming): measured at 125 M rows, back to back: 0.1396 ms per scan without skew, 0.1379 with 6,
    // 0.1349 with 16, 0.1347 with 24 (profiles/r2_pdl_ab.jsonl) -- the 1 B-row rate; shorter tables take a proportional skew
    uint64_t skew = (uint64_t)std::max(0, env_int("AQE_SCAN_SKEW", 16));
    skew = std::min<uint64_t>(skew, n_main_tiles / ((uint64_t)grid * 4));
    if (a.pdl_tail && skew && bps == 2 && grid == 2 * db->sm_count) {
        a.long_ctas = (unsigned)(grid / 2);
        a.even_rounds = (n_main_tiles - (uint64_t)a.long_ctas * skew) / (uint64_t)grid;
    }
    g_scan_kernel += " grid=" + std::to_string(grid) + " x " + std::to_string(kBulkThreads) + " threads (" + std::to_string(bps) + " CTAs/SM), " + std::to_string(smem) +
                     " B dynamic smem" + (a.pdl_tail ? ", programmatic dependent launch" : "");
    if (a.pdl_tail) {
        // programmatic stream serialization: this launch may begin while the previous kernel of the stream is still draining (it
        // said so with griddepcontrol.launch_dependents); see ScanArgs::pdl_tail for what keeps that safe
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kBulkThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, kernel, a));
    } else {
        kernel<<<grid, kBulkThreads, smem, s>>>(a);
    }
    LAUNCHED();
    return AQE_OK;
}

// Kernel selection.  Default for every aligned scan: the TMA-staged ring, 4 stages x 16 KiB, 2 CTAs/SM
// (128 KiB in flight per SM) -- 7.2-7.4 TB/s on the headline query in the round-1 sweeps vs 7.0-7.3 for the
// register-staged kernel, and immune to ptxas sinking loads (which cost the integer-predicate register kernel half
// its bandwidth).  AQE_SCAN_VARIANT: 0 ring (default) | 2 ring with AQE_SCAN_STAGES in {2,3,4,6} | 4 register-staged
// LDG.256 kernel | 1 register-staged LDG.128 kernel.  Read per call so tools/scan_sweep.py can sweep in one process.
template <typename AggT, int PRED, typename PredT, bool MOMENTS>
static int launch_scan_t(const aqe_db* db, const ScanArgs& a, bool aligned, cudaStream_t s) {
    const ScanTuning t = scan_tuning();
    {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "<%s, PRED=%d, %s, MOMENTS=%d>", type_name<AggT>(), PRED, type_name<PredT>(), MOMENTS ? 1 : 0);
        g_scan_kernel = std::string(!aligned ? "aqe::k_scan_unaligned" : (t.variant == 0 || t.variant == 2 ? "aqe::k_scan_ring (TMA bulk-copy ring)" : "aqe::k_scan (register-staged)")) + buf;
        if (aligned && (t.variant == 0 || t.variant == 2)) g_scan_kernel += " STAGES=" + std::to_string(t.variant == 0 ? 4 : t.stages) + " x 16 KiB";
    }
    if (!aligned) {
        const int grid = grid_for(db, a.n, 8, kScanThreads, 8);
        k_scan_unaligned<AggT, PRED, PredT, MOMENTS><<<grid, kScanThreads, 0, s>>>(a);
        LAUNCHED();
        return AQE_OK;
    }
    using G = RingGeom<AggT, PRED, PredT>;
    if (t.variant == 0) return launch_ring_kernel(db, k_scan_ring<AggT, PRED, PredT, 4, MOMENTS>, a, 4, G::kRows, t.bps > 0 ? t.bps : 2, s);
    if constexpr (std::is_same_v<AggT, double> && PRED != 2 && !MOMENTS) {
        if (t.variant == 2) {
            if (t.stages == 2) return launch_ring_kernel(db, k_scan_ring<AggT, PRED, PredT, 2, false>, a, 2, G::kRows, t.bps, s);
            if (t.stages == 3) return launch_ring_kernel(db, k_scan_ring<AggT, PRED, PredT, 3, false>, a, 3, G::kRows, t.bps, s);
            if (t.stages == 6) return launch_ring_kernel(db, k_scan_ring<AggT, PRED, PredT, 6, false>, a, 6, G::kRows, t.bps, s);
            return launch_ring_kernel(db, k_scan_ring<AggT, PRED, PredT, 4, false>, a, 4, G::kRows, t.bps, s);
        }
        if (t.variant == 1) return launch_regs_kernel(db, k_scan<AggT, PRED, PredT, 2, 8, false, 2>, a, 2, 8, t.bps, s);
        if (t.variant == 4 && t.unroll == 6) return launch_regs_kernel(db, k_scan<AggT, PRED, PredT, 4, 6, false, 2>, a, 4, 6, t.bps, s);
        if (t.variant == 4 && t.minb == 3) return launch_regs_kernel(db, k_scan<AggT, PRED, PredT, 4, 4, false, 3>, a, 4, 4, t.bps, s);
    }
    // register-staged kernel: all-4-byte scans read 8 elements per 256-bit load, everything else 4 per unit
    constexpr int W = (sizeof(AggT) == 4 && (PRED != 2 || sizeof(PredT) == 4)) ? 8 : 4;
    return launch_regs_kernel(db, k_scan<AggT, PRED, PredT, W, 4, MOMENTS, 1>, a, W, 4, t.bps, s);
}

// {v in int64 : lo <= (double)v <= hi} is an interval because int64 -> double conversion is monotone; its end points
// are found by bisection on that monotone predicate, so the integer compare in the kernel selects exactly the rows
// the reference-shaped double compare would (also beyond 2^53 where the conversion rounds).
static void integer_bounds(double lo, double hi, long long* ilo, long long* ihi) {
    *ilo = 1; *ihi = 0;  // empty
    if (!(lo == lo) || !(hi == hi) || lo > hi) return;
    auto first_true = [](auto pred) -> unsigned long long {  // smallest offset u in [0, 2^64) with pred(u); 2^64-1 assumed true
        unsigned long long a = 0, b = ~0ull;
        while (a < b) { const unsigned long long m = a + (b - a) / 2; if (pred(m)) b = m; else a = m + 1; }
        return a;
    };
    auto from_off = [](unsigned long long u) { return (long long)(u ^ 0x8000000000000000ull); };  // order-preserving u64 -> i64
    if ((double)INT64_MAX < lo || (double)INT64_MIN > hi) return;
    const unsigned long long ul = first_true([&](unsigned long long u) { return (double)from_off(u) >= lo; });
    // largest v with (double)v <= hi  =  (smallest v with (double)v > hi) - 1, or INT64_MAX if none
    long long up;
    if ((double)INT64_MAX <= hi) up = INT64_MAX;
    else up = from_off(first_true([&](unsigned long long u) { return (double)from_off(u) > hi; })) - 1;
    *ilo = from_off(ul); *ihi = up;
}

template <typename AggT, bool MOMENTS>
static int launch_scan_pred(const aqe_db* db, const ScanArgs& a, int pred_mode, int pred_kind, bool aligned, cudaStream_t s) {
    if (pred_mode == 0) return launch_scan_t<AggT, 0, AggT, MOMENTS>(db, a, aligned, s);
    if (pred_mode == 1) return launch_scan_t<AggT, 1, AggT, MOMENTS>(db, a, aligned, s);
    switch (pred_kind) {
        case K_F64: return launch_scan_t<AggT, 2, double, MOMENTS>(db, a, aligned, s);
        case K_I64: return launch_scan_t<AggT, 2, int64_t, MOMENTS>(db, a, aligned, s);
        default: return launch_scan_t<AggT, 2, int32_t, MOMENTS>(db, a, aligned, s);
    }
}

// Launches the scan of rows [first, first+n) of the handle's columns; result lands at out_dev.
static int scan_launch(aqe_db* db, const aqe_scan_spec* spec, uint64_t first, uint64_t n, bool moments, aqe_partial* out_dev,
                       cudaStream_t s, bool exchange = false) {
    const int ak = col_kind(spec->agg_col);
    if (ak < 0) return fail(AQE_ERR_INVALID, "bad aggregate column");
    if (ak != K_F64 && n > (1ull << 32))  // the split 32/32-bit integer accumulators are exact up to 2^32 rows per launch
        return fail(AQE_ERR_UNSUPPORTED, "integer aggregates over more than 2^32 rows per shard: split the shard");
    const char* agg = static_cast<const char*>(col_ptr(db, spec->agg_col));
    if (!agg && n) return fail(AQE_ERR_STATE, "aggregate column is not resident on the device");
    int pred_mode = 0, pk = ak;
    const char* pred = nullptr;
    if (spec->pred_col != AQE_COL_NONE) {
        pk = col_kind(spec->pred_col);
        if (pk < 0) return fail(AQE_ERR_INVALID, "bad predicate column");
        if (spec->pred_col == spec->agg_col) pred_mode = 1;
        else {
            pred_mode = 2;
            pred = static_cast<const char*>(col_ptr(db, spec->pred_col));
            if (!pred && n) return fail(AQE_ERR_STATE, "predicate column is not resident on the device");
        }
    }
    ScanArgs a;
    std::memset(&a.ex, 0, sizeof(a.ex));
    if (exchange) {
        if (!db->ex_connected) return fail(AQE_ERR_STATE, "aqe_exchange_connect has not been called");
        a.ex.world = db->ex_world; a.ex.rank = db->ex_rank; a.ex.is_integer = ak != K_F64;
        a.ex.split = env_int("AQE_EXCHANGE_SPLIT", 1) != 0 ? 1 : 0;
        a.ex.local = db->ex_local;
        a.ex.seq = ++db->ex_seq;
        a.ex.timeout_cycles = (unsigned long long)env_int("AQE_EXCHANGE_TIMEOUT_MS", 5000) * 2000000ull;
        for (int r = 0; r < db->ex_world; ++r) a.ex.peers[r] = db->ex_peers[r];
        a.ex.status = &db->slot_dev->flags[0];
    }
    a.agg = agg ? agg + first * kind_size(ak) : nullptr;
    a.pred = pred ? pred + first * kind_size(pk) : nullptr;
    a.n = n; a.lo = spec->lo; a.hi = spec->hi;
    integer_bounds(spec->lo, spec->hi, &a.ilo, &a.ihi);
    a.partials = db->scan_partials; a.ticket = db->tickets + 0; a.out = out_dev;
    // back-to-back scans overlap their tails (programmatic dependent launch) when the columns are the handle's own: borrowed columns
    // may have been written by the caller's previous kernel on this stream, which only a full stream dependency orders
    a.pdl_tail = db->owned && env_int("AQE_SCAN_PDL", 1) != 0 ? 1u : 0u;
    // 256-bit vector loads on every column
    auto aligned_for = [](const void* p, int) { return ((uintptr_t)p % 32) == 0; };
    const bool aligned = aligned_for(a.agg, ak) && (pred_mode != 2 || aligned_for(a.pred, pk));
    int rc;
    if (ak == K_F64) rc = moments ? launch_scan_pred<double, true>(db, a, pred_mode, pk, aligned, s) : launch_scan_pred<double, false>(db, a, pred_mode, pk, aligned, s);
    else if (ak == K_I64) rc = launch_scan_pred<int64_t, false>(db, a, pred_mode, pk, aligned, s);
    else rc = launch_scan_pred<int32_t, false>(db, a, pred_mode, pk, aligned, s);
    if (rc) return rc;
    CU(cudaGetLastError());
    if (exchange && a.ex.split) {
        // the wait for the peers + the rank-order fold: one warp, right behind the scan (k_scan_merge, aqe_kernels.cuh)
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr; cfg.numAttrs = a.pdl_tail ? 1 : 0;
        CU(cudaLaunchKernelEx(&cfg, k_scan_merge, a.ex, out_dev));
        LAUNCHED();
    }
    return AQE_OK;
}

static int scan_sync(aqe_db* db, const aqe_scan_spec* spec, uint64_t first, uint64_t n, bool moments, aqe_partial* out) {
    if (db->group) return group_scan_range(db, spec, first, n, moments, out);
    const uint64_t kSeg = 1ull << 32;  // integer aggregates: one launch is exact up to 2^32 rows; longer shards go in segments
    if (col_kind(spec->agg_col) > K_F64 && n > kSeg) {
        std::vector<aqe_partial> parts;
        for (uint64_t off = 0; off < n; off += kSeg) {
            aqe_partial p;
            const int rc = scan_sync(db, spec, first + off, std::min(kSeg, n - off), moments, &p);
            if (rc) return rc;
            parts.push_back(p);
        }
        return aqe_merge_partials(parts.data(), (int)parts.size(), 1, out);
    }
    int rc = scan_launch(db, spec, first, n, moments, &db->slot_dev->partial, db->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(db->stream));
    *out = db->slot_host->partial;
    return AQE_OK;
}

extern "C" {

int aqe_scan(aqe_db* db, const aqe_scan_spec* spec, aqe_partial* out) {
    if (!db || !spec || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    return scan_sync(db, spec, 0, db->n, true, out);
}

int aqe_scan_async(aqe_db* db, const aqe_scan_spec* spec, void* partial_dev, void* stream) {
    if (!db || !spec || !partial_dev) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_scan_async");
    int rc = ensure_device(db);
    if (rc) return rc;
    return scan_launch(db, spec, 0, db->n, false, static_cast<aqe_partial*>(partial_dev), stream ? (cudaStream_t)stream : db->stream);
}

int aqe_exchange_init(aqe_db* db, int rank, int world, void* ipc_handle_out) {
    if (!db || !ipc_handle_out) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_exchange_init (the shards of a sharded handle are already connected)");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(AQE_ERR_INVALID, "rank/world out of range (at most 16 ranks)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    int rc = db_init_cuda(db);
    if (rc) return rc;
    if (!db->ex_mailbox) {
        // [0, 2*kMaxRanks): scan partials; [2*kMaxRanks, 4*kMaxRanks): per-look messages of the sampled estimators
        // behind them: the sequence flags and accumulator slots of the SQL exchange (aqe_sql_kernels.cuh)
        CU(cudaMalloc(&db->ex_mailbox, kMailboxBytes));
    }
    CU(cudaMemset(db->ex_mailbox, 0, kMailboxBytes));
    db->ex_rank = rank; db->ex_world = world; db->ex_seq = 0; db->ax_msg = 0; db->sqlx_seq = 0; db->ex_connected = false;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, db->ex_mailbox));
    std::memcpy(ipc_handle_out, &h, sizeof(h));
    return AQE_OK;
}

int aqe_exchange_connect(aqe_db* db, const void* all_handles) {
    if (!db || !all_handles) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_exchange_connect");
    if (!db->ex_mailbox) return fail(AQE_ERR_STATE, "call aqe_exchange_init first");
    CU(cudaSetDevice(db->device));
    for (int r = 0; r < db->ex_world; ++r) {
        if (r == db->ex_rank) { db->ex_peers[r] = db->ex_mailbox; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(all_handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        db->ex_peers[r] = static_cast<ExSlot*>(p);
    }
    db->ex_connected = true; db->ex_ipc = true;
    return AQE_OK;
}

int aqe_scan_exchange_async(aqe_db* db, const aqe_scan_spec* spec, void* merged_dev, void* stream) {
    if (!db || !spec || !merged_dev) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_scan_exchange_async");
    int rc = ensure_device(db);
    if (rc) return rc;
    return scan_launch(db, spec, 0, db->n, false, static_cast<aqe_partial*>(merged_dev), stream ? (cudaStream_t)stream : db->stream, db->ex_world > 1);
}

int aqe_exchange_check(aqe_db* db) {
    if (db && db->group) return AQE_OK;
    if (!db || !db->cuda_ready) return fail(AQE_ERR_STATE, "no device state");
    CU(cudaSetDevice(db->device));
    CU(cudaStreamSynchronize(db->stream));   // the flag lives in the mapped pinned slot: no copy, but the kernels must have finished
    if (db->slot_host->flags[0]) { db->slot_host->flags[0] = 0; return fail(AQE_ERR_CUDA, "fused exchange timed out waiting for a peer rank"); }
    return AQE_OK;
}

int aqe_scan_exchange(aqe_db* db, const aqe_scan_spec* spec, aqe_partial* out) {
    if (!db || !spec || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return aqe_scan(db, spec, out);   // a sharded handle exchanges inside every scan
    int rc = ensure_device(db);
    if (rc) return rc;
    rc = scan_launch(db, spec, 0, db->n, true, &db->slot_dev->partial, db->stream, db->ex_world > 1);
    if (rc) return rc;
    CU(cudaStreamSynchronize(db->stream));
    if (db->slot_host->flags[0]) { db->slot_host->flags[0] = 0; return fail(AQE_ERR_CUDA, "fused exchange timed out waiting for a peer rank"); }
    *out = db->slot_host->partial;
    return AQE_OK;
}

int aqe_merge_partials(const aqe_partial* parts, int n, int is_integer, aqe_partial* out) {
    if ((!parts && n > 0) || !out || n < 0) return fail(AQE_ERR_INVALID, "bad argument");   // (no partials: the empty result)
    aqe_partial r;
    std::memset(&r, 0, sizeof(r));
    r.minv = INFINITY; r.maxv = -INFINITY;
    __int128 isum = 0;
    // double-double fold in rank order (TwoSum), so the merged sum does not depend on the shard count
    // beyond the last bit
    double s = 0.0, c = 0.0, q = 0.0;
    for (int i = 0; i < n; ++i) {
        const aqe_partial& p = parts[i];
        r.count += p.count;
        isum += ((__int128)p.isum_hi << 64) + (__int128)p.isum_lo;
        volatile double t = s + p.sum;
        volatile double z = t - s;
        volatile double e = (s - (t - z)) + (p.sum - z);
        c += p.comp + e;
        s = t;
        q += p.sumsq;
        if (p.count) { r.minv = std::min(r.minv, p.minv); r.maxv = std::max(r.maxv, p.maxv); }
    }
    volatile double t = s + c;
    r.comp = c - (t - s);
    r.sum = t;
    r.sumsq = q;
    r.isum_lo = (uint64_t)isum; r.isum_hi = (int64_t)(isum >> 64);
    if (is_integer) { r.sum = (double)isum; r.comp = 0.0; }
    *out = r;
    return AQE_OK;
}

int aqe_sum_f64(aqe_db* db, int col, double* out) {
    if (!out) return fail(AQE_ERR_INVALID, "NULL out");
    if (col_kind(col) != K_F64) return fail(AQE_ERR_INVALID, "aqe_sum_f64 needs an f64 column");
    if (aqe_count(db) == 0) { *out = 0.0; return db ? AQE_OK : fail(AQE_ERR_INVALID, "NULL handle"); }
    int rc = ensure_device(db);
    if (rc) return rc;
    aqe_scan_spec sp{col, AQE_COL_NONE, 0.0, 0.0};
    aqe_partial p;
    rc = scan_sync(db, &sp, 0, db->n, false, &p);
    if (rc) return rc;
    *out = p.sum;
    return AQE_OK;
}

int aqe_sum_where_f64(aqe_db* db, int col, double lo, double hi, double* sum, uint64_t* count) {
    if (col_kind(col) != K_F64) return fail(AQE_ERR_INVALID, "aqe_sum_where_f64 needs an f64 column");
    if (aqe_count(db) == 0) { if (sum) *sum = 0.0; if (count) *count = 0; return db ? AQE_OK : fail(AQE_ERR_INVALID, "NULL handle"); }
    int rc = ensure_device(db);
    if (rc) return rc;
    aqe_scan_spec sp{col, col, lo, hi};
    aqe_partial p;
    rc = scan_sync(db, &sp, 0, db->n, false, &p);
    if (rc) return rc;
    if (sum) *sum = p.sum;
    if (count) *count = p.count;
    return AQE_OK;
}

int aqe_sum_i128(aqe_db* db, int col, uint64_t* lo64, int64_t* hi64) {
    const int k = col_kind(col);
    if (k != K_I64 && k != K_I32) return fail(AQE_ERR_INVALID, "aqe_sum_i128 needs an integer column");
    if (aqe_count(db) == 0) { if (lo64) *lo64 = 0; if (hi64) *hi64 = 0; return db ? AQE_OK : fail(AQE_ERR_INVALID, "NULL handle"); }
    int rc = ensure_device(db);
    if (rc) return rc;
    aqe_scan_spec sp{col, AQE_COL_NONE, 0.0, 0.0};
    aqe_partial p;
    rc = scan_sync(db, &sp, 0, db->n, false, &p);
    if (rc) return rc;
    if (lo64) *lo64 = p.isum_lo;
    if (hi64) *hi64 = p.isum_hi;
    return AQE_OK;
}

// ---- host-resident column through the device (end-to-end form) -------------------------------------------
struct HostScanCtx {
    int device = -1;
    aqe_db* db = nullptr;           // scratch (partials, tickets) + stream
    cudaStream_t streams[2] = {nullptr, nullptr};
    void* dev[2] = {nullptr, nullptr};
    void* pinned[2] = {nullptr, nullptr};
    cudaEvent_t freed[2] = {nullptr, nullptr};
    aqe_partial* parts_dev = nullptr;
    ScanAcc* partials2 = nullptr;   // second scratch so the two streams never share a partial array
    unsigned int* tickets2 = nullptr;
    size_t chunk_bytes = 0;
    size_t max_chunks = 0;
};
static std::mutex g_hs_mu;
static HostScanCtx g_hs[16];

static int host_scan_ctx(int device, HostScanCtx** out) {
    if (device < 0 || device >= 16) return fail(AQE_ERR_INVALID, "device out of range");
    HostScanCtx& c = g_hs[device];
    if (c.device == device) { CU(cudaSetDevice(device)); *out = &c; return AQE_OK; }
    int rc = aqe_create(device, &c.db);
    if (rc) return rc;
    rc = db_init_cuda(c.db);
    if (rc) return rc;
    c.chunk_bytes = (size_t)env_int("AQE_E2E_CHUNK_MB", 64) << 20;
    c.max_chunks = 4096;
    for (int i = 0; i < 2; ++i) {
        CU(cudaStreamCreateWithFlags(&c.streams[i], cudaStreamNonBlocking));
        CU(cudaMalloc(&c.dev[i], c.chunk_bytes));
        CU(cudaEventCreateWithFlags(&c.freed[i], cudaEventDisableTiming));
    }
    CU(cudaMalloc(&c.parts_dev, sizeof(aqe_partial) * c.max_chunks));
    CU(cudaMalloc(&c.partials2, sizeof(ScanAcc) * kMaxGrid));
    CU(cudaMalloc(&c.tickets2, 16));
    CU(cudaMemset(c.tickets2, 0, 16));
    c.device = device;
    *out = &c;
    return AQE_OK;
}

// Chunks of a host column through ONE device: copy in on two alternating streams, scan, 64-byte partial per chunk.  `next` hands out
// chunk numbers (shared by the devices of aqe_scan_host_column_multi: a device whose link is slower simply takes fewer chunks);
// parts[k] receives chunk k's partial whichever device scanned it -- every B200 runs the same grid over the same bytes, so the
// partial of a chunk, and with it the merged result, does not depend on who took it.
static int host_scan_chunks(HostScanCtx* c, const void* host_col, int col_kind_id, uint64_t n, uint64_t per_chunk, uint64_t nchunks, bool pinned,
                            double lo, double hi, int use_pred, std::atomic<uint64_t>* next, aqe_partial* parts) {
    CU(cudaSetDevice(c->device));
    const size_t esz = kind_size(col_kind_id);
    if (!pinned) {
        for (int i = 0; i < 2; ++i)
            if (!c->pinned[i]) CU(cudaHostAlloc(&c->pinned[i], c->chunk_bytes, cudaHostAllocDefault));
    }
    aqe_db* db = c->db;
    std::vector<uint64_t> mine;
    // whatever goes wrong below, the copies and scans already queued on this device's two streams finish before the buffers are used again
    struct Quiesce { HostScanCtx* c; bool armed = true; ~Quiesce() { if (armed) for (int i = 0; i < 2; ++i) cudaStreamSynchronize(c->streams[i]); } } quiesce{c};
    int rc = AQE_OK;
    for (;;) {
        const uint64_t k = next->fetch_add(1, std::memory_order_relaxed);
        if (k >= nchunks) break;
        const size_t slot = mine.size();
        const int b = (int)(slot & 1);
        const uint64_t first = k * per_chunk, cnt = std::min<uint64_t>(per_chunk, n - first);
        const char* src = static_cast<const char*>(host_col) + first * esz;
        if (!pinned) {
            CU(cudaEventSynchronize(c->freed[b]));
            std::memcpy(c->pinned[b], src, cnt * esz);
            src = static_cast<const char*>(c->pinned[b]);
        }
        CU(cudaMemcpyAsync(c->dev[b], src, cnt * esz, cudaMemcpyHostToDevice, c->streams[b]));
        ScanArgs a;   // a one-column view of the chunk
        std::memset(&a.ex, 0, sizeof(a.ex));
        a.agg = c->dev[b]; a.pred = nullptr; a.n = cnt; a.lo = lo; a.hi = hi; a.pdl_tail = 0;   // the chunk was just copied in on this stream
        integer_bounds(lo, hi, &a.ilo, &a.ihi);
        a.partials = b ? c->partials2 : db->scan_partials;
        a.ticket = b ? c->tickets2 : db->tickets + 1;
        a.out = c->parts_dev + slot;
        const int pm = use_pred ? 1 : 0;
        if (col_kind_id == K_F64) rc = launch_scan_pred<double, false>(db, a, pm, K_F64, true, c->streams[b]);
        else if (col_kind_id == K_I64) rc = launch_scan_pred<int64_t, false>(db, a, pm, K_I64, true, c->streams[b]);
        else rc = launch_scan_pred<int32_t, false>(db, a, pm, K_I32, true, c->streams[b]);
        if (rc) return rc;
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->freed[b], c->streams[b]));
        mine.push_back(k);
    }
    for (int i = 0; i < 2; ++i) CU(cudaStreamSynchronize(c->streams[i]));
    quiesce.armed = false;
    if (mine.empty()) return AQE_OK;
    std::vector<aqe_partial> got(mine.size());
    CU(cudaMemcpy(got.data(), c->parts_dev, sizeof(aqe_partial) * mine.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < mine.size(); ++i) parts[mine[i]] = got[i];
    return AQE_OK;
}

static bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

int aqe_scan_host_column(int device, const void* host_col, int col_kind_id, uint64_t n, double lo, double hi, int use_pred,
                         aqe_partial* out) {
    if (!out || (!host_col && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (col_kind_id < 0 || col_kind_id > 2) return fail(AQE_ERR_INVALID, "col_kind: 0 f64, 1 i64, 2 i32");
    std::lock_guard<std::mutex> lock(g_hs_mu);
    HostScanCtx* c = nullptr;
    int rc = host_scan_ctx(device, &c);
    if (rc) return rc;
    const uint64_t per_chunk = c->chunk_bytes / kind_size(col_kind_id);
    const uint64_t nchunks = n ? (n + per_chunk - 1) / per_chunk : 0;
    if (nchunks > c->max_chunks) return fail(AQE_ERR_UNSUPPORTED, "column too large for aqe_scan_host_column; raise AQE_E2E_CHUNK_MB");
    std::vector<aqe_partial> parts(nchunks);
    std::atomic<uint64_t> next{0};
    rc = host_scan_chunks(c, host_col, col_kind_id, n, per_chunk, nchunks, n && host_pointer_is_pinned(host_col), lo, hi, use_pred, &next, parts.data());
    if (rc) return rc;
    return aqe_merge_partials(parts.data(), (int)nchunks, col_kind_id != K_F64, out);
}

// The same column through SEVERAL devices of this process: one host thread per device, chunks handed out from one counter, the
// chunks' partials merged in chunk order.  Links that share a PCIe root or sit far from the column's NUMA node run slower (24-35 GB/s
// per GPU with eight at once against 55 alone); with equal shards the step ends with the slowest of them, here every link
// stays busy until the column is through.
int aqe_scan_host_column_multi(const int* devices, int n_devices, const void* host_col, int col_kind_id, uint64_t n, double lo, double hi,
                               int use_pred, aqe_partial* out) {
    if (!out || !devices || (!host_col && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (n_devices < 1 || n_devices > 16) return fail(AQE_ERR_INVALID, "n_devices: 1 .. 16");
    if (col_kind_id < 0 || col_kind_id > 2) return fail(AQE_ERR_INVALID, "col_kind: 0 f64, 1 i64, 2 i32");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(AQE_ERR_INVALID, "devices: every device once");
    std::lock_guard<std::mutex> lock(g_hs_mu);
    std::vector<HostScanCtx*> ctx((size_t)n_devices, nullptr);
    for (int i = 0; i < n_devices; ++i) {
        const int rc = host_scan_ctx(devices[i], &ctx[i]);
        if (rc) return rc;
    }
    const uint64_t per_chunk = ctx[0]->chunk_bytes / kind_size(col_kind_id);
    const uint64_t nchunks = n ? (n + per_chunk - 1) / per_chunk : 0;
    if (nchunks > ctx[0]->max_chunks) return fail(AQE_ERR_UNSUPPORTED, "column too large for aqe_scan_host_column_multi; raise AQE_E2E_CHUNK_MB");
    std::vector<aqe_partial> parts(nchunks);
    std::atomic<uint64_t> next{0};
    const bool pinned = n && host_pointer_is_pinned(host_col);
    std::vector<int> rcs((size_t)n_devices, AQE_OK);
    std::vector<std::string> errs((size_t)n_devices);
    std::vector<std::thread> th;
    for (int i = 1; i < n_devices; ++i)
        th.emplace_back([&, i] {
            rcs[i] = host_scan_chunks(ctx[i], host_col, col_kind_id, n, per_chunk, nchunks, pinned, lo, hi, use_pred, &next, parts.data());
            if (rcs[i]) errs[i] = g_err;   // (thread-local)
        });
    rcs[0] = host_scan_chunks(ctx[0], host_col, col_kind_id, n, per_chunk, nchunks, pinned, lo, hi, use_pred, &next, parts.data());
    if (rcs[0]) errs[0] = g_err;
    for (auto& t : th) t.join();
    for (int i = 0; i < n_devices; ++i)
        if (rcs[i]) return fail(rcs[i], errs[i]);
    return aqe_merge_partials(parts.data(), (int)nchunks, col_kind_id != K_F64, out);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// plans on the device, sampled aggregates
// ------------------------------------------------------------------------------------------------
static int plan_to_device(aqe_db* db, const aqe_plan* pl, PlanDev* out) {
    PlanDev P{};
    P.count = pl->count; P.nseg = (uint32_t)pl->segs.size(); P.perm = nullptr;
    const size_t seg_bytes = pl->segs.size() * sizeof(aqe_segment);
    const size_t start_bytes = pl->seg_start.size() * sizeof(uint64_t);
    const size_t idx_bytes = pl->segs.empty() ? pl->idx.size() * sizeof(int64_t) : 0;
    const size_t need = ((seg_bytes + 255) & ~(size_t)255) + ((start_bytes + 255) & ~(size_t)255) + idx_bytes + 256;
    if (need > db->plan_cap) {
        if (db->plan_buf) cudaFree(db->plan_buf);
        db->plan_buf = nullptr; db->plan_cap = 0;
        const size_t cap = std::max<size_t>(need, 1u << 20);
        CU(cudaMalloc(&db->plan_buf, cap));
        db->plan_cap = cap;
    }
    char* base = static_cast<char*>(db->plan_buf);
    if (!pl->segs.empty()) {
        CU(cudaMemcpyAsync(base, pl->segs.data(), seg_bytes, cudaMemcpyHostToDevice, db->stream));
        char* st = base + ((seg_bytes + 255) & ~(size_t)255);
        CU(cudaMemcpyAsync(st, pl->seg_start.data(), start_bytes, cudaMemcpyHostToDevice, db->stream));
        P.segs = reinterpret_cast<const aqe_segment*>(base);
        P.seg_start = reinterpret_cast<const uint64_t*>(st);
    } else if (idx_bytes) {
        CU(cudaMemcpyAsync(base, pl->idx.data(), idx_bytes, cudaMemcpyHostToDevice, db->stream));
        P.idx = reinterpret_cast<const int64_t*>(base);
    }
    *out = P;
    return AQE_OK;
}

static int check_plan_bounds(const aqe_db* db, const aqe_plan* pl) {
    // A generated plan is in range for the table size it was built for (and for every larger table); used on a smaller table
    // -- after a reload, or built with an explicit n_rows -- its segments would expand to rows the columns do not have.
    if (pl->n_rows > db->n) return fail(AQE_ERR_INVALID, "sample plan was built for " + std::to_string(pl->n_rows) + " rows, the table holds " + std::to_string(db->n));
    // explicit lists supplied by callers are validated
    if (pl->n_rows == 0)
        for (int64_t v : pl->idx)
            if (v < 0 || (uint64_t)v >= db->n) return fail(AQE_ERR_INVALID, "sample index out of range");
    return AQE_OK;
}

static int ensure_amount_perm(aqe_db* db);  // below (stratified)

// The rows of the table this handle holds, as a sample plan sees them: a plain handle holds the whole table; a shard of a
// range-sharded table holds the window [first, first + n) and reads the amount-sorted permutation of the WHOLE table.
struct PlanWindow {
    uint64_t first = 0, n = ~0ull;
    const int64_t* perm = nullptr;   // by_amount_order plans of a sharded table: the table-level permutation (peer mapped)
};

// Moments over the plan's positions: finished (`out`, whole table on this handle) or, with `raw`, the mergeable sums of the
// positions inside the window (aqe_stats_merge).  Synchronous.
static int stats_launch(aqe_db* db, const aqe_plan* pl, int col, aqe_stats* out, int pred_col = AQE_COL_NONE, double lo = 0.0, double hi = 0.0,
                        const PlanWindow* win = nullptr, aqe_stats_partial* raw = nullptr) {
    if (col_kind(col) < 0) return fail(AQE_ERR_INVALID, "bad column");
    if (pred_col != AQE_COL_NONE && (col_kind(pred_col) < 0 || (!col_ptr(db, pred_col) && db->n))) return fail(AQE_ERR_INVALID, "bad predicate column");
    if (!col_ptr(db, col) && pl->count && db->n) return fail(AQE_ERR_STATE, "column is not resident on the device");
    if (raw) std::memset(raw, 0, sizeof(*raw));
    if (pl->count == 0 || (raw && db->n == 0)) { if (out) { out->n = 0; out->mean = 0; out->m2 = 0; out->sum = 0; } return AQE_OK; }
    StatArgs a;
    int rc = plan_to_device(db, pl, &a.plan);
    if (rc) return rc;
    if (pl->by_amount_order) {
        if (win && win->perm) a.plan.perm = win->perm;
        else { rc = ensure_amount_perm(db); if (rc) return rc; a.plan.perm = db->amount_perm; }
    }
    a.cols = const_cols(db); a.col = col; a.pred_col = pred_col; a.lo = lo; a.hi = hi; a.partials = db->stat_partials; a.ticket = db->tickets + 2; a.out = &db->slot_dev->stats;
    a.win_first = win ? win->first : 0; a.win_n = win ? std::min<uint64_t>(win->n, db->n) : ~0ull;
    a.raw_out = raw ? &db->slot_dev->stats_raw : nullptr;
    bool tiles = !pl->segs.empty();   // every segment a run of contiguous rows?
    for (const aqe_segment& sg : pl->segs) tiles = tiles && sg.kind == 0 && sg.inner_len >= 32;
    if (pl->by_amount_order) tiles = false;
    const int grid = grid_for(db, pl->count, tiles ? 4 : 8, 256, 8);
    if (tiles) k_plan_stats<4><<<grid, 256, 0, db->stream>>>(a);
    else k_plan_stats<8><<<grid, 256, 0, db->stream>>>(a);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(db->stream));
    if (raw) *raw = db->slot_host->stats_raw;
    else *out = db->slot_host->stats;
    return AQE_OK;
}

// Rows [k_first, k_first + cnt) of the plan that fall into the window -> dst[k - k_first] (device-visible memory), on the
// handle's stream, not synchronised.  `counter` (device-visible, may be NULL) receives the number of rows written.
static int gather_kernel(aqe_db* db, const PlanDev& P, aqe_record* dst, uint64_t k_first, uint64_t cnt, const PlanWindow& win, unsigned long long* counter) {
    k_plan_gather<<<grid_for(db, cnt, 2, 256, 8), 256, 0, db->stream>>>(P, const_cols(db), dst, k_first, cnt, win.first, std::min<uint64_t>(win.n, db->n), counter);
    LAUNCHED();
    CU(cudaGetLastError());
    return AQE_OK;
}

static int gather_launch(aqe_db* db, const aqe_plan* pl, aqe_record* out, uint64_t cap) {
    const uint64_t n = std::min<uint64_t>(pl->count, cap);
    if (n == 0) return AQE_OK;
    PlanDev P;
    int rc = plan_to_device(db, pl, &P);
    if (rc) return rc;
    if (pl->by_amount_order) { rc = ensure_amount_perm(db); if (rc) return rc; P.perm = db->amount_perm; }
    const uint64_t chunk = 1u << 22;
    rc = ensure_gather_buf(db, std::min<uint64_t>(n, chunk));
    if (rc) return rc;
    PlanWindow whole;
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - off);
        if ((rc = gather_kernel(db, P, db->gather_buf, off, cnt, whole, nullptr))) return rc;
        CU(cudaMemcpyAsync(out + off, db->gather_buf, cnt * sizeof(aqe_record), cudaMemcpyDeviceToHost, db->stream));
        CU(cudaStreamSynchronize(db->stream));
    }
    return AQE_OK;
}

// ---- amount-sorted permutation (stratified_block_sample sorts the table by amount, cbd:1343-1345) ----
__global__ void k_iota(int64_t* p, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (int64_t)i;
}
// perm[r] = row with the r-th smallest amount among `keys[0, n)` (stable: ties keep ascending row order), on db's device / stream.
static int sort_rows_by_amount(aqe_db* db, const double* keys, uint64_t n, int64_t** perm_out) {
    int64_t *iota = nullptr, *perm = nullptr;
    double* keys_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&](bool keep_perm) { cudaFree(tmp); cudaFree(iota); cudaFree(keys_out); if (!keep_perm) cudaFree(perm); };
#define CUS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(false); return cuda_fail(e_, #call); } } while (0)
    CUS(cudaMalloc(&iota, n * 8));
    CUS(cudaMalloc(&perm, n * 8));
    CUS(cudaMalloc(&keys_out, n * 8));
    k_iota<<<grid_for(db, n, 4, 256, 8), 256, 0, db->stream>>>(iota, n);
    LAUNCHED();
    // stable LSD radix sort on the f64 keys
    CUS(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, iota, perm, (int64_t)n, 0, 64, db->stream));
    CUS(cudaMalloc(&tmp, tmp_bytes));
    CUS(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, iota, perm, (int64_t)n, 0, 64, db->stream));
    LAUNCHED();
    CUS(cudaStreamSynchronize(db->stream));
#undef CUS
    cleanup(true);
    *perm_out = perm;
    return AQE_OK;
}
static int ensure_amount_perm(aqe_db* db) {
    if (db->amount_perm || db->n == 0) return AQE_OK;
    if (!db->col.amount) return fail(AQE_ERR_STATE, "amount column is not resident on the device");
    return sort_rows_by_amount(db, db->col.amount, db->n, &db->amount_perm);
}

// ---- lock-step stop step of clt_validated_dual_pointer_sample on the device ---------------------------------
// One CTA per sampler thread q walks that thread's stride sequence in chunks of blockDim.x samples, keeps
// running sums of d = x - K and d^2 (K = first sample) with a block scan, evaluates the reference's check
// (custom_bplus_db.cpp:936-961 fast, :993-1016 slow) at every checkpoint inside the chunk and records per
// thread the first step at which ITS OWN rule would fire given the published fast mean; the host then
// resolves the global first stop in (step, thread) order.
struct CltArgs {
    GlobalF64 amount;   // the whole table's amount column (one part on a plain handle, one per shard on a sharded one)
    CltThread th[64];
    int nthreads;
    int64_t check_interval, T;
    double z, max_err;
    // per (thread, checkpoint) outputs for the host resolution
    double* mean_out;   // [nthreads][max_checks]
    double* err_out;    // fast: error percent ; slow: unused
    uint64_t max_checks;
    uint64_t max_steps; // only steps <= max_steps are evaluated in this launch
};

__global__ void __launch_bounds__(1024) k_clt_prefix(const CltArgs a) {
    __shared__ double s_d[32], s_dd[32];
    __shared__ double carry_d, carry_dd;
    const int q = blockIdx.x;
    const CltThread t = a.th[q];
    const uint64_t every = t.fast ? (uint64_t)a.check_interval : (uint64_t)(a.check_interval / 2);
    const uint64_t minn = t.fast ? 30 : 20;
    const uint64_t steps = t.len < a.max_steps ? t.len : a.max_steps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double K = t.len ? a.amount.at(t.first) : 0.0;
    if (threadIdx.x == 0) { carry_d = 0.0; carry_dd = 0.0; }
    __syncthreads();
    for (uint64_t base = 0; base < steps; base += blockDim.x) {
        const uint64_t k = base + threadIdx.x;  // 0-based sample number
        double d = 0.0, dd = 0.0;
        if (k < steps) { d = a.amount.at(t.first + k * t.step) - K; dd = d * d; }
        // inclusive block scan of (d, dd)
        double pd = d, pdd = dd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double x = __shfl_up_sync(0xffffffffu, pd, o), y = __shfl_up_sync(0xffffffffu, pdd, o);
            if (lane >= o) { pd += x; pdd += y; }
        }
        if (lane == 31) { s_d[warp] = pd; s_dd[warp] = pdd; }
        __syncthreads();
        if (warp == 0) {
            double x = lane < nw ? s_d[lane] : 0.0, y = lane < nw ? s_dd[lane] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, x, o), v = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) { x += u; y += v; }
            }
            if (lane < nw) { s_d[lane] = x; s_dd[lane] = y; }
        }
        __syncthreads();
        const double off_d = carry_d + (warp ? s_d[warp - 1] : 0.0), off_dd = carry_dd + (warp ? s_dd[warp - 1] : 0.0);
        const double Sd = off_d + pd, Sdd = off_dd + pdd;
        const uint64_t cnt = k + 1;
        if (k < steps && cnt % every == 0 && cnt >= minn) {
            const double n = (double)cnt;
            const double mean = K + Sd / n;
            double var = (Sdd - Sd * Sd / n) / (n - 1.0);
            if (var < 0.0) var = 0.0;
            const uint64_t ci = cnt / every - 1;
            if (ci < a.max_checks) {
                a.mean_out[(size_t)q * a.max_checks + ci] = mean;
                a.err_out[(size_t)q * a.max_checks + ci] = (a.z * sqrt(var / n) / mean) * 100.0;
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry_d = Sd; carry_dd = Sdd; }
        __syncthreads();
    }
}

// `db`: a plain handle, or the sharded handle (then the walk runs on shard 0's GPU and reads the other shards' rows through
// the peer mapping -- a latency-bound walk of a few thousand samples, not worth splitting).
static int clt_resolve(aqe_db* db, const aqe_sample_params& P, PlanData& data) {
    std::vector<CltThread> th;
    int64_t T = 0;
    std::string err;
    int rc = clt_threads(db->n, P, th, T, err);
    if (rc) return fail(rc, err);
    data.clt_kstop = 0; data.clt_stopper = -1;
    if (th.empty()) return AQE_OK;
    if (th.size() > 64) return fail(AQE_ERR_UNSUPPORTED, "clt_validated_dual_pointer_sample: at most 64 threads");
    GlobalF64 view;
    std::memset(&view, 0, sizeof(view));
    if (db->group) {
        if ((rc = group_amount_view(db, &view))) return rc;
        db = aqe_shard(db, 0);
        CU(cudaSetDevice(db->device));
    } else {
        if (!db->col.amount) return fail(AQE_ERR_STATE, "amount column is not resident on the device");
        view.base[0] = db->col.amount; view.first[0] = 0; view.first[1] = db->n; view.parts = 1;
    }
    const int64_t ci = P.check_interval;
    const double z = P.confidence_level >= 0.99 ? 2.576 : (P.confidence_level >= 0.95 ? 1.96 : 1.645);  // cbd:911-912
    const int F = (int)(P.num_threads / 2);
    uint64_t maxlen = 0;
    for (auto& t : th) maxlen = std::max<uint64_t>(maxlen, t.len);
    // windows of steps; stop as soon as one window contains the stop step
    const uint64_t every_min = std::max<int64_t>(1, ci / 2);
    double current_mean = 0.0;
    uint64_t sample_count = 0;
    uint64_t window = 1u << 15, begin = 0;
    std::vector<double> mean_h, err_h;
    double *mean_d = nullptr, *err_d = nullptr;
    uint64_t cap_checks = 0;
    while (begin < maxlen) {
        const uint64_t end = std::min<uint64_t>(maxlen, begin ? begin * 4 : window);
        const uint64_t checks = end / every_min + 1;
        if (checks > cap_checks) {
            if (mean_d) { cudaFree(mean_d); cudaFree(err_d); }
            CU(cudaMalloc(&mean_d, th.size() * checks * 8));
            CU(cudaMalloc(&err_d, th.size() * checks * 8));
            cap_checks = checks;
        }
        CltArgs a;
        a.amount = view; a.nthreads = (int)th.size();
        for (size_t q = 0; q < th.size(); ++q) a.th[q] = th[q];
        a.check_interval = ci; a.T = T; a.z = z; a.max_err = P.max_error_percent;
        a.mean_out = mean_d; a.err_out = err_d; a.max_checks = cap_checks; a.max_steps = end;
        k_clt_prefix<<<(int)th.size(), 1024, 0, db->stream>>>(a);
        LAUNCHED();
        CU(cudaGetLastError());
        mean_h.resize(th.size() * cap_checks); err_h.resize(th.size() * cap_checks);
        CU(cudaMemcpyAsync(mean_h.data(), mean_d, mean_h.size() * 8, cudaMemcpyDeviceToHost, db->stream));
        CU(cudaMemcpyAsync(err_h.data(), err_d, err_h.size() * 8, cudaMemcpyDeviceToHost, db->stream));
        CU(cudaStreamSynchronize(db->stream));
        // replay the lock-step schedule over steps (begin, end]: fast threads first, thread order
        for (uint64_t k = begin + 1; k <= end; ++k) {
            for (size_t q = 0; q < th.size(); ++q) {
                if (k > th[q].len) continue;
                const bool fast = (int)q < F;
                const uint64_t every = fast ? (uint64_t)ci : (uint64_t)(ci / 2), minn = fast ? 30 : 20;
                if (k % every != 0 || k < minn) continue;
                const uint64_t c = k / every - 1;
                const double mean = mean_h[q * cap_checks + c];
                if (fast) {
                    current_mean = mean; sample_count = k;                              // cbd:949-951
                    if (err_h[q * cap_checks + c] <= P.max_error_percent && k >= 50) {   // cbd:958
                        data.clt_kstop = (int64_t)k; data.clt_stopper = (int64_t)q; goto done;
                    }
                } else if (current_mean > 0) {
                    const double md = std::fabs(mean - current_mean) / current_mean;   // cbd:1008
                    if (md <= P.max_error_percent / 100.0 && sample_count >= (uint64_t)(T / 2)) {
                        data.clt_kstop = (int64_t)k; data.clt_stopper = (int64_t)q; goto done;
                    }
                }
            }
        }
        begin = end;
    }
done:
    if (mean_d) { cudaFree(mean_d); cudaFree(err_d); }
    return AQE_OK;
}

extern "C" {

void aqe_sample_params_default(aqe_sample_params* p, int method) {  // bindings.cpp:56-101
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->sample_percent = 1.0;
    p->step_size = method == AQE_M_RANDOM_START_NTH ? 10 : 2;
    p->num_threads = 4;
    p->block_size = 1000;
    if (method == AQE_M_PAGE) p->block_size = 4096;
    if (method == AQE_M_ADAPTIVE_BLOCK) p->block_size = 500;
    if (method == AQE_M_MEMORY_STRIDE || method == AQE_M_RANDOM_START_MEMORY_STRIDE) p->block_size = 0;
    p->block_size_max = method == AQE_M_STRATIFIED_BLOCK ? 4 : 2000;
    p->check_interval = method == AQE_M_OPTIMIZED_CLT ? 20 : 10;
    p->confidence_level = 0.95;
    p->max_error_percent = 2.0;
    p->seed = 42;
}

int aqe_plan_build(aqe_db* db, uint64_t n_rows, int method, const aqe_sample_params* p, aqe_plan** out) {
    if (!p || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    PlanData data;
    double zone_var[10];
    const bool needs_data = method == AQE_M_ADAPTIVE_BLOCK || method == AQE_M_CLT_VALIDATED_DUAL_POINTER;
    if (needs_data) {
        if (!db) return fail(AQE_ERR_STATE, "this sampler reads the table: pass a handle");
        int rc = ensure_device(db);
        if (rc) return rc;
        n_rows = db->n;
        const int64_t T = (int64_t)((double)n_rows * p->sample_percent / 100.0);
        if (method == AQE_M_ADAPTIVE_BLOCK && n_rows >= 10 && T > 0) {
            // zone variance = sum_sq/count - mean^2 (custom_bplus_db.cpp:1294-1304) from ten exact scans
            const uint64_t zs = n_rows / 10;
            for (int z = 0; z < 10; ++z) {
                const uint64_t s = z * zs, e = std::min<uint64_t>(s + zs, n_rows);
                aqe_scan_spec sp{AQE_COL_AMOUNT, AQE_COL_NONE, 0.0, 0.0};
                aqe_partial part;
                rc = scan_sync(db, &sp, s, e - s, true, &part);
                if (rc) return rc;
                const double cnt = (double)(e - s), mean = part.sum / cnt;
                zone_var[z] = part.sumsq / cnt - mean * mean;
            }
            data.zone_var = zone_var;
        } else if (method == AQE_M_ADAPTIVE_BLOCK) {
            for (double& v : zone_var) v = 1.0;
            data.zone_var = zone_var;
        }
        if (method == AQE_M_CLT_VALIDATED_DUAL_POINTER && n_rows > 0 && T > 0) {
            rc = clt_resolve(db, *p, data);
            if (rc) return rc;
        } else if (method == AQE_M_CLT_VALIDATED_DUAL_POINTER) {
            data.clt_kstop = 0;
        }
    } else if (db && n_rows == 0) {
        n_rows = aqe_count(db);
    }
    aqe_plan* pl = new (std::nothrow) aqe_plan();
    if (!pl) return fail(AQE_ERR_NOMEM, "out of host memory");
    std::string err;
    const int rc = plan_build(n_rows, method, *p, data, *pl, err);
    if (rc) { delete pl; return fail(rc, err); }
    pl->n_rows = n_rows;
    *out = pl;
    return AQE_OK;
}

int aqe_plan_from_indices(const int64_t* idx, uint64_t n, aqe_plan** out) {
    if (!out || (!idx && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    aqe_plan* pl = new (std::nothrow) aqe_plan();
    if (!pl) return fail(AQE_ERR_NOMEM, "out of host memory");
    pl->idx.assign(idx, idx + n);
    plan_finalize(*pl);
    *out = pl;
    return AQE_OK;
}

uint64_t aqe_plan_count(const aqe_plan* plan) { return plan ? plan->count : 0; }
uint64_t aqe_plan_table_rows(const aqe_plan* plan) { return plan ? plan->n_rows : 0; }
uint32_t aqe_plan_num_segments(const aqe_plan* plan) { return plan ? (uint32_t)plan->segs.size() : 0; }
int aqe_plan_segments(const aqe_plan* plan, aqe_segment* out, uint32_t cap) {
    if (!plan || (!out && cap)) return fail(AQE_ERR_INVALID, "NULL argument");
    const uint32_t n = std::min<uint32_t>(cap, (uint32_t)plan->segs.size());
    if (n) std::memcpy(out, plan->segs.data(), n * sizeof(aqe_segment));
    return AQE_OK;
}
int aqe_plan_indices(const aqe_plan* plan, int64_t* out, uint64_t cap) {
    if (!plan || (!out && cap)) return fail(AQE_ERR_INVALID, "NULL argument");
    const uint64_t n = std::min<uint64_t>(cap, plan->count);
    for (uint64_t k = 0; k < n; ++k) out[k] = plan_position_host(*plan, k);
    return AQE_OK;
}
int aqe_plan_sorted_by_amount(const aqe_plan* plan) { return plan && plan->by_amount_order ? 1 : 0; }
void aqe_plan_free(aqe_plan* plan) { delete plan; }

int aqe_stats_from_plan(aqe_db* db, const aqe_plan* plan, int col, aqe_stats* out) {
    if (!db || !plan || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    rc = check_plan_bounds(db, plan);
    if (rc) return rc;
    if (db->group) return group_stats(db, plan, col, AQE_COL_NONE, 0.0, 0.0, out);
    return stats_launch(db, plan, col, out);
}

int aqe_stats_from_plan_where(aqe_db* db, const aqe_plan* plan, int col, int pred_col, double lo, double hi, aqe_stats* out) {
    if (!db || !plan || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    rc = check_plan_bounds(db, plan);
    if (rc) return rc;
    if (db->group) return group_stats(db, plan, col, pred_col, lo, hi, out);
    return stats_launch(db, plan, col, out, pred_col, lo, hi);
}

// ---- one process per GPU: this handle is the window [window_first, window_first + n) of the plan's table ----
int aqe_stats_window(aqe_db* db, const aqe_plan* plan, int col, int pred_col, double lo, double hi, uint64_t window_first, aqe_stats_partial* out) {
    if (!db || !plan || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_stats_window");
    int rc = ensure_device(db);
    if (rc) return rc;
    if (plan->by_amount_order) return fail(AQE_ERR_UNSUPPORTED, "plans in amount order (stratified_block_sample) need the whole table's permutation: use a sharded handle (aqe_create_sharded)");
    if (plan->n_rows && window_first + db->n > plan->n_rows) return fail(AQE_ERR_INVALID, "window exceeds the table the plan was built for");
    PlanWindow w;
    w.first = window_first; w.n = db->n;
    return stats_launch(db, plan, col, nullptr, pred_col, lo, hi, &w, out);
}

int aqe_stats_merge(const aqe_stats_partial* parts, int n, aqe_stats* out) {
    if (!parts || !out || n < 0) return fail(AQE_ERR_INVALID, "bad argument");
    // rank order.  sum x: double-double fold; mean / M2: Chan, Golub & LeVeque's pairwise update (each shard's moments are about
    // its own shift K_g: mean_g = K_g + sd/n, M2_g = sdd - sd^2/n)
    uint64_t N = 0;
    double mean = 0.0, m2 = 0.0, s = 0.0, c = 0.0;
    for (int i = 0; i < n; ++i) {
        const aqe_stats_partial& p = parts[i];
        if (p.n == 0) continue;
        volatile double t = s + p.sum;
        volatile double z = t - s;
        volatile double e = (s - (t - z)) + (p.sum - z);
        c += p.sum_c + e;
        s = t;
        const double np = (double)p.n;
        const double sd = p.sd + p.sd_c;
        const double mean_p = p.shift + sd / np;
        double m2_p = (p.sdd + p.sdd_c) - sd * sd / np;
        if (m2_p < 0.0) m2_p = 0.0;
        if (N == 0) { mean = mean_p; m2 = m2_p; }
        else {
            const double na = (double)N, tot = na + np, delta = mean_p - mean;
            m2 += m2_p + delta * delta * (na * np / tot);
            mean += delta * (np / tot);
        }
        N += p.n;
    }
    out->n = N;
    out->sum = s + c;
    out->mean = N ? out->sum / (double)N : 0.0;
    out->m2 = m2;
    return AQE_OK;
}

int aqe_gather_window(aqe_db* db, const aqe_plan* plan, uint64_t window_first, uint64_t k_first, uint64_t k_count, aqe_record* out, uint64_t* n_local) {
    if (!db || !plan || (!out && k_count)) return fail(AQE_ERR_INVALID, "NULL argument");
    if (db->group) return group_unsupported("aqe_gather_window");
    int rc = ensure_device(db);
    if (rc) return rc;
    if (plan->by_amount_order) return fail(AQE_ERR_UNSUPPORTED, "plans in amount order (stratified_block_sample) need the whole table's permutation: use a sharded handle (aqe_create_sharded)");
    if (plan->n_rows && window_first + db->n > plan->n_rows) return fail(AQE_ERR_INVALID, "window exceeds the table the plan was built for");
    if (k_first > plan->count || k_count > plan->count - k_first) return fail(AQE_ERR_INVALID, "plan range out of bounds");
    if (plan->n_rows == 0)
        for (int64_t v : plan->idx) if (v < 0) return fail(AQE_ERR_INVALID, "sample index out of range");
    if (n_local) *n_local = 0;
    if (k_count == 0) return AQE_OK;
    PlanDev P;
    if ((rc = plan_to_device(db, plan, &P))) return rc;
    const uint64_t chunk = 1u << 22;
    if ((rc = ensure_gather_buf(db, std::min<uint64_t>(k_count, chunk)))) return rc;
    PlanWindow w;
    w.first = window_first; w.n = db->n;
    db->slot_host->gathered = 0;
    for (uint64_t off = 0; off < k_count; off += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, k_count - off);
        CU(cudaMemsetAsync(db->gather_buf, 0, cnt * sizeof(aqe_record), db->stream));   // slots of other ranks' rows stay zero
        if (db->n && (rc = gather_kernel(db, P, db->gather_buf, k_first + off, cnt, w, &db->slot_dev->gathered))) return rc;
        CU(cudaMemcpyAsync(out + off, db->gather_buf, cnt * sizeof(aqe_record), cudaMemcpyDeviceToHost, db->stream));
        CU(cudaStreamSynchronize(db->stream));
    }
    if (n_local) *n_local = db->slot_host->gathered;
    return AQE_OK;
}

int aqe_stats_from_indices(aqe_db* db, const int64_t* idx, uint64_t n, int col, aqe_stats* out) {
    if (!db || !out || (!idx && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    aqe_plan pl;
    pl.idx.assign(idx, idx + n);
    plan_finalize(pl);
    return aqe_stats_from_plan(db, &pl, col, out);
}

int aqe_gather_plan(aqe_db* db, const aqe_plan* plan, aqe_record* out, uint64_t cap) {
    if (!db || !plan || (!out && cap)) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    rc = check_plan_bounds(db, plan);
    if (rc) return rc;
    if (db->group) return group_gather(db, plan, out, cap);
    return gather_launch(db, plan, out, cap);
}

int aqe_gather_records(aqe_db* db, const int64_t* idx, uint64_t n, aqe_record* out) {
    if (!db || (!idx && n) || (!out && n)) return fail(AQE_ERR_INVALID, "NULL argument");
    aqe_plan pl;
    pl.idx.assign(idx, idx + n);
    plan_finalize(pl);
    return aqe_gather_plan(db, &pl, out, n);
}

int aqe_fast_aggregated_sum(aqe_db* db, const aqe_sample_params* p, double* sum, uint64_t* n) {
    if (!db || !p) return fail(AQE_ERR_INVALID, "NULL argument");
    if (aqe_count(db) == 0) { if (sum) *sum = 0.0; if (n) *n = 0; return AQE_OK; }
    int rc = ensure_device(db);
    if (rc) return rc;
    aqe_plan pl;
    std::string err;
    rc = plan_build(db->n, AQE_M_MULTITHREADED_MEMORY_STRIDE, *p, PlanData{}, pl, err);
    if (rc) return fail(rc, err);
    pl.n_rows = db->n;
    aqe_stats st;
    rc = db->group ? group_stats(db, &pl, AQE_COL_AMOUNT, AQE_COL_NONE, 0.0, 0.0, &st) : stats_launch(db, &pl, AQE_COL_AMOUNT, &st);
    if (rc) return rc;
    if (sum) *sum = st.n ? st.sum : 0.0;  // raw, unscaled sample sum (custom_bplus_db.cpp:2045-2047)
    if (n) *n = st.n;
    return AQE_OK;
}

int aqe_estimate(const aqe_stats* s, uint64_t population, int agg, double z, int legacy_ci, double* estimate,
                 double* ci_lower, double* ci_upper) {
    if (!s) return fail(AQE_ERR_INVALID, "NULL stats");
    if (s->n == 0) return fail(AQE_ERR_INVALID, "no samples");  // the CLI raises "No samples collected"
    const double n = (double)s->n, N = (double)population;
    double e;
    if (agg == AQE_AGG_SUM) e = s->sum * (N / n);        // enhanced_aqe_cli.py:190-193
    else if (agg == AQE_AGG_COUNT) e = N;                // :196-197
    else e = s->sum / n;                                 // :194-195
    const double var = s->n > 1 ? s->m2 / (n - 1.0) : 0.0;  // :279
    const double moe = z * std::sqrt(var) / std::sqrt(n);   // :281
    double m;
    if (agg == AQE_AGG_SUM) m = legacy_ci ? moe * (N / n) : moe * N;  // :285 vs the correct scaling
    else if (agg == AQE_AGG_COUNT) m = 0.0;
    else m = moe;
    if (estimate) *estimate = e;
    if (ci_lower) *ci_lower = e - m;
    if (ci_upper) *ci_upper = e + m;
    return AQE_OK;
}

double aqe_z_score(double conf, int exact) {
    if (!exact) return conf >= 0.99 ? 2.576 : (conf >= 0.95 ? 1.96 : 1.645);  // custom_bplus_db.cpp:911-912
    // inverse normal CDF (P. Acklam's rational approximation, |rel err| < 1.15e-9), p = 1 - (1-conf)/2
    const double p = 1.0 - (1.0 - conf) / 2.0;
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    if (p < 0.02425) {
        const double q = std::sqrt(-2 * std::log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    if (p <= 1 - 0.02425) {
        const double q = p - 0.5, r = q * q;
        return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    }
    const double q = std::sqrt(-2 * std::log(1 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
}

// ------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------
// normal quantile the interval of a ci_mode is built from (include/aqe_b200.h, aqe_ci_mode) and whether it is Stein-type
static double ci_z(double confidence_level, uint32_t ci_mode, int* stein) {
    const uint32_t m = ci_mode == AQE_CI_DEFAULT ? (uint32_t)AQE_CI_STEIN_GUARDED : ci_mode;
    if (stein) *stein = m != AQE_CI_PLAIN;
    const double alpha = 1.0 - confidence_level;
    return aqe_z_score(1.0 - (m == AQE_CI_STEIN_GUARDED ? AQE_CI_GUARD : 1.0) * alpha, 1);
}

static int approx_run(aqe_db* db, const aqe_approx_spec* S, aqe_approx_result* out, bool multi) {
    if (!db || !S || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    std::memset(out, 0, sizeof(*out));
    if (!(S->error_percent > 0.0)) return fail(AQE_ERR_INVALID, "error_percent must be > 0");
    if (!(S->confidence_level > 0.0 && S->confidence_level < 1.0)) return fail(AQE_ERR_INVALID, "confidence_level must be in (0,1)");
    if (S->agg < 0 || S->agg > 2 || S->design < 0 || S->design > 1) return fail(AQE_ERR_INVALID, "bad agg / design");
    if (S->ci_mode > AQE_CI_STEIN_GUARDED) return fail(AQE_ERR_INVALID, "bad ci_mode");
    const uint64_t N = aqe_count(db);
    const uint32_t B = S->design == AQE_DESIGN_BLOCK ? (S->block_size ? S->block_size : 1000) : 1;
    uint64_t rows_total = N, units_total = (N + B - 1) / B;
    if (multi) {
        if (!db->ex_connected) return fail(AQE_ERR_STATE, "aqe_exchange_connect has not been called");
        if (!db->ex_total_rows) return fail(AQE_ERR_STATE, "aqe_exchange_set_total_rows has not been called");
        rows_total = db->ex_total_rows;
        units_total = 0;
        for (int g = 0; g < db->ex_world; ++g) {  // contiguous shards [N g/G, N (g+1)/G)
            const uint64_t ng = (uint64_t)(((unsigned __int128)rows_total * (g + 1)) / db->ex_world) - (uint64_t)(((unsigned __int128)rows_total * g) / db->ex_world);
            if (g == db->ex_rank && ng != N) return fail(AQE_ERR_STATE, "this shard does not hold rows [N*rank/world, N*(rank+1)/world) of the table");
            units_total += (ng + B - 1) / B;
        }
    }
    out->population = rows_total; out->confidence_level = S->confidence_level;
    if (rows_total == 0) { out->status = AQE_INSUFFICIENT_DATA; return AQE_OK; }
    if (S->agg == AQE_AGG_COUNT && S->pred_col == AQE_COL_NONE) {  // enhanced_aqe_cli.py:196-197: COUNT is exact
        out->estimate = out->ci_lower = out->ci_upper = (double)rows_total; out->status = AQE_STABLE; out->pass_fraction = 1.0;
        return AQE_OK;
    }
    int rc = ensure_device(db);
    if (rc) return rc;
    if (S->agg != AQE_AGG_COUNT && col_kind(S->agg_col) < 0) return fail(AQE_ERR_INVALID, "bad aggregate column");
    if (N && S->agg != AQE_AGG_COUNT && !col_ptr(db, S->agg_col)) return fail(AQE_ERR_STATE, "aggregate column is not resident on the device");
    if (N && S->pred_col != AQE_COL_NONE && !col_ptr(db, S->pred_col)) return fail(AQE_ERR_STATE, "predicate column is not resident on the device");

    ApproxArgs a;
    std::memset(&a, 0, sizeof(a));
    a.cols = const_cols(db);
    a.n_rows = N;
    a.block_rows = B;
    a.units = (N + B - 1) / B;
    a.design = S->design; a.agg = S->agg; a.agg_col = S->agg_col; a.pred_col = S->pred_col;
    a.lo = S->lo; a.hi = S->hi; a.eps = S->error_percent; a.z = ci_z(S->confidence_level, S->ci_mode, &a.stein);
    a.seed = S->seed + (multi ? 0x9E3779B97F4A7C15ull * (uint64_t)db->ex_rank : 0ull);  // independent stream per stratum
    const uint64_t n0 = S->min_samples ? S->min_samples : (S->design == AQE_DESIGN_BLOCK ? 1024 : 16384);
    const uint64_t nmax = S->max_samples ? S->max_samples : units_total;
    a.n0 = std::min(n0, nmax); a.nmax = nmax;
    a.units_total = units_total; a.rows_total = rows_total; a.n0_total = a.n0; a.nmax_total = nmax;
    a.slots = db->approx_slots;
    a.out = &db->slot_dev->approx;
    if (units_total <= a.n0) {
        // the first look would already draw as many units as the table has: scan it exactly instead
        aqe_scan_spec sp{S->agg == AQE_AGG_COUNT ? (int32_t)(S->pred_col) : S->agg_col, S->pred_col, S->lo, S->hi};
        rc = scan_launch(db, &sp, 0, db->n, false, &db->slot_dev->partial, db->stream, multi);
        if (rc) return rc;
        CU(cudaStreamSynchronize(db->stream));
        if (multi) { rc = aqe_exchange_check(db); if (rc) return rc; }
        const aqe_partial part = db->slot_host->partial;
        double v;
        if (S->agg == AQE_AGG_COUNT) v = (double)part.count;
        else if (S->agg == AQE_AGG_SUM) v = part.sum;
        else v = part.count ? part.sum / (double)part.count : 0.0;
        out->estimate = out->ci_lower = out->ci_upper = v;
        out->n_samples = rows_total; out->n_units = units_total; out->status = AQE_STABLE;
        out->pass_fraction = rows_total ? (double)part.count / (double)rows_total : 0.0;
        return AQE_OK;
    }
    if (multi) {
        a.ex.world = db->ex_world; a.ex.rank = db->ex_rank; a.ex.seq = db->ax_msg;
        a.ex.timeout_cycles = (unsigned long long)env_int("AQE_EXCHANGE_TIMEOUT_MS", 5000) * 2000000ull;
        for (int r = 0; r < db->ex_world; ++r) a.ex.peers[r] = db->ex_peers[r];
        a.ex.status = &db->slot_dev->flags[0];
    }

    static int coop_blocks_per_sm = 0;
    if (!coop_blocks_per_sm) {
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_approx, 256, 0);
        coop_blocks_per_sm = std::max(1, occ);
    }
    int grid = db->sm_count * std::min(coop_blocks_per_sm, env_int("AQE_APPROX_BPS", 1));
    if (grid > db->max_grid) grid = db->max_grid;
    void* params[] = {&a};
    CU(cudaEventRecord(db->ev0, db->stream));
    CU(cudaLaunchCooperativeKernel((void*)k_approx, dim3(grid), dim3(256), params, 0, db->stream));
    LAUNCHED();
    CU(cudaEventRecord(db->ev1, db->stream));
    CU(cudaStreamSynchronize(db->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, db->ev0, db->ev1));
    *out = db->slot_host->approx;
    out->confidence_level = S->confidence_level;
    out->elapsed_us = (double)ms * 1000.0;
    if (multi) {
        db->ax_msg += out->rounds;  // every rank ran the same number of looks (one global decision per look)
        rc = aqe_exchange_check(db);
        if (rc) return rc;
    }
    return AQE_OK;
}

int aqe_approx(aqe_db* db, const aqe_approx_spec* S, aqe_approx_result* out) {
    if (db && db->group) return group_approx(db, S, out);
    return approx_run(db, S, out, false);
}

int aqe_exchange_set_total_rows(aqe_db* db, uint64_t total_rows) {
    if (!db) return fail(AQE_ERR_INVALID, "NULL handle");
    db->ex_total_rows = total_rows;
    return AQE_OK;
}

int aqe_approx_exchange(aqe_db* db, const aqe_approx_spec* S, aqe_approx_result* out) {
    if (!db) return fail(AQE_ERR_INVALID, "NULL handle");
    if (db->group) return group_approx(db, S, out);
    return approx_run(db, S, out, db->ex_world > 1);
}

int aqe_approx_merge(const aqe_approx_result* parts, int n, int agg, double confidence_level, aqe_approx_result* out) {
    if (!parts || !out || n <= 0) return fail(AQE_ERR_INVALID, "bad argument");
    // Shards are strata with independent draws: totals add, variances of the totals add.  AVG: each shard's estimate is the mean
    // of ITS matching rows, so the weights are the shards' (estimated) matching-row counts population_g * pass_fraction_g -- the
    // combined ratio estimator sum_g U_g ybar_g / sum_g U_g cbar_g the fused exchange evaluates (aqe_kernels.cuh, approx_global);
    // without a predicate pass_fraction = 1 and the weights are the shard sizes.
    const double z = ci_z(confidence_level, AQE_CI_DEFAULT, nullptr);
    aqe_approx_result r;
    std::memset(&r, 0, sizeof(r));
    double total = 0.0, var_total = 0.0, wsum = 0.0, pass_rows = 0.0;
    uint64_t pop = 0;
    int worst = AQE_STABLE;
    for (int i = 0; i < n; ++i) {
        const aqe_approx_result& p = parts[i];
        const double half = (p.ci_upper - p.ci_lower) * 0.5;
        const double w = agg == AQE_AGG_AVG ? (double)p.population * p.pass_fraction : 1.0;
        total += p.estimate * w;
        var_total += (half / z) * (half / z) * w * w;
        wsum += w; pass_rows += (double)p.population * p.pass_fraction;
        pop += p.population; r.n_samples += p.n_samples; r.n_units += p.n_units;
        r.rounds = std::max(r.rounds, p.rounds); r.elapsed_us = std::max(r.elapsed_us, p.elapsed_us);
        if (p.status > worst) worst = p.status;
    }
    if (agg == AQE_AGG_AVG) {
        if (wsum > 0.0) {
            total /= wsum;
            // the weights are estimates too: with e = y - R c (R the merged ratio) a shard contributes, besides w_g^2 Var(R_g), the term
            // (U_g (R_g - R))^2 Var(cbar_g), Var(cbar_g) = p_g (1 - p_g) / n_g -- what the linearised residual of the fused path
            // (approx_global, aqe_kernels.cuh) carries implicitly.  Zero without a predicate (p_g = 1).
            for (int i = 0; i < n; ++i) {
                const aqe_approx_result& p = parts[i];
                if (p.n_samples == 0) continue;
                const double u = (double)p.population * (p.estimate - total);
                var_total += u * u * p.pass_fraction * (1.0 - p.pass_fraction) / (double)p.n_samples;
            }
            var_total /= wsum * wsum;
        } else { total = 0.0; var_total = 0.0; }
    }
    const double half = z * std::sqrt(var_total);
    r.estimate = total; r.ci_lower = total - half; r.ci_upper = total + half;
    r.error_margin = total != 0.0 ? half / std::fabs(total) : 0.0;
    r.confidence_level = confidence_level; r.population = pop; r.status = worst;
    r.pass_fraction = pop ? pass_rows / (double)pop : 0.0;
    *out = r;
    return AQE_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// SQL-string path: run_query* of bindings.cpp:126-136 on the columnar table
// ------------------------------------------------------------------------------------------------
static int sql_init(aqe_db* db) {
    if (db->sql_acc) return AQE_OK;
    const size_t bytes = sizeof(unsigned long long) * 5 * AQE_SQL_MAX_GROUPS;
    CU(cudaMalloc(&db->sql_acc, bytes));
    CU(cudaMemset(db->sql_acc, 0, bytes));
    CU(cudaMalloc(&db->sql_local, bytes));
    CU(cudaMalloc(&db->sql_stat_dev, 3 * sizeof(unsigned long long)));
    CU(cudaMalloc(&db->sql_ticket, sizeof(unsigned int)));
    CU(cudaMemset(db->sql_ticket, 0, sizeof(unsigned int)));
    CU(cudaHostAlloc(&db->sql_out_host, bytes, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&db->sql_out_dev, db->sql_out_host, 0));
    return AQE_OK;
}

// min / max of a column (and, for id, whether ids are first_id + row), computed once per table version
static int sql_col_stat(aqe_db* db, int col, const aqe_db::ColStat** out) {
    aqe_db::ColStat& st = db->col_stat[col];
    if (!st.valid) {
        const void* ptr = col_ptr(db, col);
        if (!ptr && db->n) return fail(AQE_ERR_STATE, "a column the query needs is not resident on the device");
        st.min_key = ~0ull; st.max_key = 0ull; st.dense = false; st.first_id = 0;
        if (db->n) {
            const unsigned long long init[3] = {~0ull, 0ull, 0ull};
            CU(cudaMemcpyAsync(db->sql_stat_dev, init, sizeof(init), cudaMemcpyHostToDevice, db->stream));
            ColStatsArgs a;
            a.col = ptr; a.kind = col_kind(col); a.n = db->n; a.check_dense = col == AQE_COL_ID ? 1 : 0; a.first_id = 0; a.out = db->sql_stat_dev;
            if (a.check_dense) {
                CU(cudaMemcpyAsync(&a.first_id, ptr, 8, cudaMemcpyDeviceToHost, db->stream));
                CU(cudaStreamSynchronize(db->stream));
            }
            k_col_stats<<<grid_for(db, db->n, 8, 256, 8), 256, 0, db->stream>>>(a);
            LAUNCHED();
            CU(cudaGetLastError());
            unsigned long long res[3];
            CU(cudaMemcpyAsync(res, db->sql_stat_dev, sizeof(res), cudaMemcpyDeviceToHost, db->stream));
            CU(cudaStreamSynchronize(db->stream));
            st.min_key = res[0]; st.max_key = res[1];
            st.dense = a.check_dense && res[2] == 0;
            st.first_id = a.first_id;
        }
        st.valid = true;
    }
    *out = &st;
    return AQE_OK;
}

static double okey_f64(unsigned long long k) {
    const unsigned long long b = okey_to_bits_f64(k);
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}

// Dynamic shared memory of the SQL kernels varies per query (bins grow with the group count): opt every kernel in to
// the full 227 KiB once per device and ask the occupancy calculator per launch (host arithmetic, no driver round trip).
static int sql_occupancy(const void* kernel, int threads, size_t smem) {
    struct Key { const void* k; int dev; };
    static std::mutex mu;
    static std::vector<Key> opted;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        bool seen = false;
        for (auto& e : opted) seen = seen || (e.k == kernel && e.dev == dev);
        if (!seen) {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess)  // the opt-in limit covers static + dynamic
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
            cudaGetLastError();
            opted.push_back(Key{kernel, dev});
        }
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    return occ;
}

template <int MODE, bool MOMENTS> static int sql_launch_regs(const aqe_db* db, const SqlArgs& a, cudaStream_t s) {
    const size_t smem = SqlBins<MODE, MOMENTS, kSqlThreads>::smem_bytes(a.n_groups);
    const void* k = (const void*)k_sql_agg<MODE, MOMENTS>;
    const int occ = sql_occupancy(k, kSqlThreads, smem);
    if (occ < 1) return fail(AQE_ERR_UNSUPPORTED, "SQL path: group bins do not fit shared memory");
    const int grid = grid_for(db, a.count, 1, kSqlThreads, occ);
    k_sql_agg<MODE, MOMENTS><<<grid, kSqlThreads, smem, s>>>(a);
    LAUNCHED();
    return AQE_OK;
}

template <int MODE, bool MOMENTS> static int sql_launch_ring(const aqe_db* db, SqlRingArgs& ra, cudaStream_t s) {
    constexpr int T = kBulkConsumerWarps * 32;
    uint32_t row_bytes = 0;
    for (int i = 0; i < ra.q.ncols; ++i) row_bytes += ra.q.cols[i].kind == K_I32 ? 4 : 8;
    // Rows per consumer thread and tile, K (every row slot of a full tile is live either way).  A larger K halves the per-tile barrier
    // and bookkeeping cost, more CTAs per SM hide the shared-memory latency of the bin updates; measured on 1 B rows, 12 B/row:
    // ungrouped K = 8 / 4 CTAs 1.70 ms vs K = 4 2.25 ms; `GROUP BY region` K = 8 / 2 CTAs 2.19, K = 6 / 3 CTAs 1.95, K = 4 / 4 CTAs 2.4 ms;
    // the same with squares (registers allow two CTAs whatever K) K = 8 2.68, K = 6 2.94, K = 4 3.34 ms.
    // Rule: the largest K whose kernel still runs three CTAs per SM (shared memory AND registers), else the largest that runs two, else 4.
    const size_t bins = SqlBins<MODE, MOMENTS, T>::smem_bytes(ra.q.n_groups);
    auto ctas_with = [&](int k) -> int {
        const size_t smem_k = (size_t)2 * T * k * row_bytes + bins;
        if (smem_k > (size_t)(220 * 1024)) return 0;
        return k == 8 ? sql_occupancy((const void*)k_sql_ring<MODE, MOMENTS, 2, 8>, kBulkThreads, smem_k)
             : k == 6 ? sql_occupancy((const void*)k_sql_ring<MODE, MOMENTS, 2, 6>, kBulkThreads, smem_k)
                      : sql_occupancy((const void*)k_sql_ring<MODE, MOMENTS, 2, 4>, kBulkThreads, smem_k);
    };
    int K = 4;
    // (packed shared bins without squares: 56 registers, and at K = 6 the stages leave room for a FOURTH CTA -- 1000 groups over 1 B rows
    // 2.56 ms against 2.77 ms at K = 8 / three CTAs; the kernel is bound by latency, not by a pipe: one / two / three CTAs 6.4 / 3.5 / 2.8 ms)
    for (int want : {MODE == 3 && !MOMENTS ? 4 : 3, 3, 2}) {
        bool found = false;
        for (int k : {8, 6, 4}) if (!found && ctas_with(k) >= want) { K = k; found = true; }
        if (found) break;
    }
    if (row_bytes <= 8) K = 8;   // 16 KiB stages at most: always the largest tile
    // One int32 column and nothing else (COUNT [WHERE region ...] [GROUP BY region | product_id]): 16 KiB stages hold 16 rows per thread.  These scans
    // are bound by the consumers' instructions per row, a good half of which is per-tile bookkeeping: twice the rows per tile halves that share (1 B rows:
    // `COUNT(*) WHERE region = 1` 1.08 -> 0.77 ms, `COUNT ... GROUP BY region` 0.78 -> 0.64, `... GROUP BY product_id` 0.96 -> 0.89).  The shared-bin kernel
    // runs K = 16 under a three-CTA register budget (sql_ring_min_ctas): under the 56 registers of four CTAs it spills and loses (0.83 -> 0.95 ms).
    constexpr bool kHasK16 = !MOMENTS && (MODE == 0 || MODE == 2);
    if (kHasK16 && row_bytes == 4 && env_int("AQE_SQL_K16", 1)) K = 16;
    { const int forced = env_int("AQE_SQL_K", 0); if (forced == 4 || forced == 6 || forced == 8) K = forced; }   // experiments (tools/sql_bench.py)
    const uint32_t tile = (uint32_t)(T * K);
    ra.tile_rows = tile;
    uint32_t off = 0;
    for (int i = 0; i < ra.q.ncols; ++i) { ra.col_off[i] = off; off += tile * (ra.q.cols[i].kind == K_I32 ? 4u : 8u); }
    ra.stage_bytes = off;
    const uint64_t ntiles = (ra.q.count + tile - 1) / tile;
    // 2 stages x up to 4 CTAs/SM beat 4 stages x 2 CTAs/SM on every grouped query of tools/sql_bench.py (27 instead of
    // 18 consumer warps per SM hide the shared-memory latency of the bin updates; profiles/r1_sql_bench.json)
    constexpr int STAGES = 2;
    ra.ring_bytes = (uint32_t)STAGES * ra.stage_bytes;
    const size_t smem = (size_t)ra.ring_bytes + bins;
    auto go = [&](auto kernel) -> int {
        const int occ = sql_occupancy((const void*)kernel, kBulkThreads, smem);
        if (occ < 1) return fail(AQE_ERR_UNSUPPORTED, "SQL path: group bins do not fit shared memory");
        int grid = (int)std::min<uint64_t>((uint64_t)db->sm_count * std::min(occ, env_int("AQE_SQL_BPS", 4)), std::max<uint64_t>(ntiles, 1));
        if (grid > db->max_grid) grid = db->max_grid;
        kernel<<<grid, kBulkThreads, smem, s>>>(ra);
        LAUNCHED();
        return AQE_OK;
    };
    if constexpr (kHasK16) { if (K == 16) return go(k_sql_ring<MODE, MOMENTS, STAGES, 16>); }
    if (K == 8) return go(k_sql_ring<MODE, MOMENTS, STAGES, 8>);
    if (K == 6) return go(k_sql_ring<MODE, MOMENTS, STAGES, 6>);
    return go(k_sql_ring<MODE, MOMENTS, STAGES, 4>);
}

static int sql_scan_impl(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* L, int flags, uint64_t* acc, bool exchange = false) {
    const uint32_t G = L->n_groups;
    if (G < 1 || G > AQE_SQL_MAX_GROUPS) return fail(AQE_ERR_INVALID, "layout: n_groups out of range");
    std::memset(acc, 0, sizeof(uint64_t) * 5 * G);
    int rc = sql_init(db);
    if (rc) return rc;
    SqlExchange ex;
    std::memset(&ex, 0, sizeof(ex));
    if (exchange) {
        if (!db->ex_connected) return fail(AQE_ERR_STATE, "aqe_exchange_connect has not been called");
        ex.world = db->ex_world; ex.rank = db->ex_rank; ex.seq = ++db->sqlx_seq;
        ex.timeout_cycles = (unsigned long long)env_int("AQE_EXCHANGE_TIMEOUT_MS", 5000) * 2000000ull;
        for (int r = 0; r < db->ex_world; ++r) ex.peers[r] = reinterpret_cast<unsigned char*>(db->ex_peers[r]);
        ex.status = &db->slot_dev->flags[0]; ex.local = db->sql_local;
    }
    // a shard that has nothing to scan: without an exchange the zeros (or the metadata count) are the answer; with one it still
    // publishes them so that the other ranks' kernels are not left waiting
    auto finish_without_scan = [&](uint64_t count0) -> int {
        acc[0] = count0;
        if (!exchange) return AQE_OK;
        CU(cudaMemcpyAsync(db->sql_local, acc, sizeof(uint64_t) * 5 * G, cudaMemcpyHostToDevice, db->stream));
        k_sql_exchange_only<<<1, 256, 0, db->stream>>>(ex, G, db->sql_out_dev);
        LAUNCHED();
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(db->stream));
        std::memcpy(acc, db->sql_out_host, sizeof(uint64_t) * 5 * G);
        return aqe_exchange_check(db);
    };
    if (q->always_false || db->n == 0) return finish_without_scan(0);
    // Accumulator widths: a thread sums 32-bit halves into 64-bit words (fewer than 2^32 rows per thread), a CTA counts
    // rows per bin in 32 bits (fewer than 2^32 rows per CTA), everything above is 128-bit.  2^40 rows keeps all of that
    // far from overflow at any grid this library launches (and is 8 TB of one 8-byte column).
    if (db->n > (1ull << 40)) return fail(AQE_ERR_UNSUPPORTED, "SQL path: more than 2^40 rows per shard");
    const bool unsampled = (flags & AQE_SQL_UNSAMPLED) != 0;
    const bool moments = (flags & AQE_SQL_MOMENTS) != 0 && !unsampled;
    const bool sums = q->agg_col != AQE_COL_NONE && !unsampled && (q->agg != AQE_AGG_COUNT || moments);
    const int step = unsampled ? 0 : sql_sample_step(q->sample_percent);

    SqlRingArgs ra;
    std::memset(&ra, 0, sizeof(ra));
    SqlArgs& a = ra.q;
    a.agg_slot = -1; a.group_slot = -1;
    auto slot_of = [&](int col) -> int {
        const void* ptr = col_ptr(db, col);
        for (int i = 0; i < a.ncols; ++i) if (a.cols[i].ptr == ptr) return i;
        a.cols[a.ncols].ptr = ptr; a.cols[a.ncols].kind = col_kind(col);
        return a.ncols++;
    };
    auto need = [&](int col) -> int {
        if (col_kind(col) < 0) return fail(AQE_ERR_INVALID, "bad column in query");
        if (!col_ptr(db, col)) return fail(AQE_ERR_STATE, "a column the query needs is not resident on the device");
        return AQE_OK;
    };
    if (sums) {
        if ((rc = need(q->agg_col))) return rc;
        a.agg_slot = slot_of(q->agg_col); a.agg_kind = col_kind(q->agg_col);
    }
    if (q->group_col != AQE_COL_NONE) {
        if ((rc = need(q->group_col))) return rc;
        a.group_slot = slot_of(q->group_col);
    }
    if (q->n_alt < 0 || q->n_alt > AQE_SQL_MAX_ALT) return fail(AQE_ERR_INVALID, "query: n_alt out of range");
    a.n_alt = q->n_alt;
    for (int alt = 0; alt < q->n_alt; ++alt)
        if (q->n_terms[alt] < 1 || q->n_terms[alt] > 5) return fail(AQE_ERR_INVALID, "query: a WHERE branch needs 1..5 terms");
    // OR branches that differ in ONE integer column whose values span at most AQE_SQL_MAX_GROUPS keys (IN lists, NOT IN, chains of
    // !=, with or without further AND-ed terms shared by all branches) fold into one branch: that column's test becomes a membership
    // bitmap over [min, max] built from the column's statistics -- one predicate pass instead of one per branch (measured on 1 B
    // rows: `region IN (1, 3, 5, 7)` 3.56 ms as four branches, 1.73 ms folded).  Fewer than 64 keys: the bitmap is one register.
    int fold_col = AQE_COL_NONE;
    uint64_t fold_bits = 0, fold_span = 0;
    bool fold_any = false;
    int64_t fold_first = 0;
    if (q->n_alt > 1) {
        for (int c = 0; c < 5 && fold_col == AQE_COL_NONE; ++c) {
            if (col_kind(c) != K_I32 && col_kind(c) != K_I64) continue;
            bool ok = col_ptr(db, c) != nullptr;
            for (int alt = 0; alt < q->n_alt && ok; ++alt) {   // every branch: exactly one term on c, the other terms those of branch 0
                if (q->n_terms[alt] != q->n_terms[0]) { ok = false; break; }
                int on_c = 0;
                for (int t = 0; t < q->n_terms[alt]; ++t) {
                    const aqe_sql_term& x = q->terms[alt][t];
                    const aqe_sql_term& y = q->terms[0][t];
                    if (x.col == c) { ++on_c; if (y.col != c) ok = false; continue; }
                    if (x.col != y.col || x.has_ne != y.has_ne || std::memcmp(&x.lo, &y.lo, sizeof(double)) || std::memcmp(&x.hi, &y.hi, sizeof(double)) ||
                        x.ilo != y.ilo || x.ihi != y.ihi || (x.has_ne && (std::memcmp(&x.ne, &y.ne, sizeof(double)) || x.ine != y.ine))) ok = false;
                }
                if (on_c != 1) ok = false;
            }
            if (!ok) continue;
            const aqe_db::ColStat* st;
            if ((rc = sql_col_stat(db, c, &st))) return rc;
            const int64_t lo = okey_to_i64(st->min_key), hi = okey_to_i64(st->max_key);
            if (hi < lo || (uint64_t)hi - (uint64_t)lo >= (uint64_t)AQE_SQL_MAX_GROUPS) continue;
            fold_span = (uint64_t)hi - (uint64_t)lo + 1;
            for (uint64_t d = 0; d < fold_span; ++d) {
                const int64_t v = lo + (int64_t)d;
                bool pass = false;
                for (int alt = 0; alt < q->n_alt && !pass; ++alt)
                    for (int t = 0; t < q->n_terms[alt]; ++t) {
                        const aqe_sql_term& x = q->terms[alt][t];
                        if (x.col == c) pass = v >= x.ilo && v <= x.ihi && !(x.has_ne && v == x.ine);
                    }
                if (!pass) continue;
                fold_any = true;
                if (fold_span <= 64) fold_bits |= 1ull << d;
                else a.member_bits[d >> 5] |= 1u << (d & 31);
            }
            fold_col = c; fold_first = lo;
        }
    }
    if (fold_col != AQE_COL_NONE) {
        if (!fold_any) return finish_without_scan(0);   // no value the column holds passes
        a.n_alt = 1;
        a.member_used = fold_span > 64 ? 1 : 0;
    }
    for (int alt = 0; alt < a.n_alt; ++alt) {
        for (int t = 0; t < q->n_terms[alt]; ++t) {
            const aqe_sql_term& term = q->terms[alt][t];
            if ((rc = need(term.col))) return rc;
            const int slot = slot_of(term.col);
            SqlPred& c = a.cols[slot].pred[alt];
            if (term.col == fold_col) {
                c.has_ne = 0; c.ne = 0; c.lo = fold_first;
                if (fold_span <= 64) { c.has_pred = 2; c.hi = (long long)fold_bits; } else { c.has_pred = 3; c.hi = (long long)fold_span; }
                continue;
            }
            c.has_pred = 1; c.has_ne = term.has_ne;
            if (a.cols[slot].kind == K_F64) { std::memcpy(&c.lo, &term.lo, 8); std::memcpy(&c.hi, &term.hi, 8); std::memcpy(&c.ne, &term.ne, 8); }
            else { c.lo = term.ilo; c.hi = term.ihi; c.ne = term.ine; }
        }
    }
    // ---- rowid % step = 0 (executor.cpp:38-42); rowid = id ----
    a.first = 0; a.stride = 1; a.count = db->n;
    bool dense = false;
    long long phase = 0;  // row i is sampled iff (i + phase) % step == 0
    if (step > 1) {
        if ((rc = need(AQE_COL_ID))) return rc;
        const aqe_db::ColStat* ids;
        if ((rc = sql_col_stat(db, AQE_COL_ID, &ids))) return rc;
        dense = ids->dense;
        if (dense) phase = ((ids->first_id % step) + step) % step;
        else a.cols[slot_of(AQE_COL_ID)].mod_step = step;  // ids with gaps: read the id column and test it
    }
    if (a.ncols == 0) {  // COUNT without WHERE / GROUP BY: metadata (0 bytes per record)
        const uint64_t first = dense ? (uint64_t)((step - phase) % step) : 0;
        return finish_without_scan(dense ? (first < db->n ? (db->n - first + step - 1) / step : 0) : db->n);
    }
    a.key_min = L->key_min; a.n_groups = G;
    a.pair_bins = env_int("AQE_SQL_PAIR_BINS", 1);
    a.sum_scale = std::ldexp(1.0, L->sum_shift); a.sq_scale = std::ldexp(1.0, L->sq_shift);
    a.global_acc = db->sql_acc; a.out = db->sql_out_dev; a.ticket = db->sql_ticket; a.ex = ex;
    // 16 rows of headroom for the ragged tail rows a thread may add after its last check; AQE_SQL_DRAIN_ROWS is a test knob (drain early)
    const int packed_rows = (int)(a.group_slot >= 0 && G <= (uint32_t)kSqlPrivateMaxGroups && moments ? kSqlPackedRowsMoments : kSqlPackedRows);   // SqlBins, MODE 1
    a.drain_rows = (unsigned int)std::min<int>(std::max(env_int("AQE_SQL_DRAIN_ROWS", packed_rows), 16), packed_rows - 16);
    bool aligned16 = true;
    for (int i = 0; i < a.ncols; ++i) aligned16 = aligned16 && ((uintptr_t)a.cols[i].ptr % 16) == 0;
    // Visit plan.  Dense ids turn the sample into an arithmetic progression of row numbers.  The ring can stream everything and
    // filter on the row number, but it then converts and masks every row: measured at 1 B rows (tools/sql_bench.py sampled), the
    // strided visit of the register kernel wins from step 2 on without a WHERE clause (grouped CI at p = 50: 2.11 vs 3.53 ms) and
    // from step 4 on with one; only the predicate passes of ungrouped queries at steps 2-3 are cheaper out of the staged tile
    // (2.07 vs 3.00 ms).  profiles/r1_sql_sampled_small_steps.json holds the table.
    // AQE_SQL_VARIANT: 0 auto | 1 register-staged kernel only | 2 ring whenever it can run (steps < 8): tests compare them.
    const int variant = env_int("AQE_SQL_VARIANT", 0);
    const int ring_below = variant == 2 ? 8 : (q->n_alt > 0 && a.group_slot < 0 ? 4 : 2);
    const bool strided = dense && (step >= ring_below || !aligned16 || variant == 1);
    if (strided) {
        a.first = (uint64_t)((step - phase) % step); a.stride = (uint64_t)step;
        a.count = a.first < db->n ? (db->n - a.first + step - 1) / step : 0;
        if (a.count == 0) return finish_without_scan(0);
    } else if (dense) {
        ra.samp_step = (uint32_t)step; ra.samp_phase = (uint32_t)phase;
    }
    const bool ring = !strided && aligned16 && variant != 1;
    int mode = a.group_slot < 0 ? 0 : (G <= (uint32_t)kSqlPrivateMaxGroups ? 1 : 2);
    // COUNT-only queries over few groups: one ATOMS.POPC.INC per row on CTA-shared counters (the instruction adds up the lanes that name
    // the same address, so 32 rows on 8 keys cost no more than on 32) instead of a read-modify-write of a private bin
    if (mode == 1 && a.agg_slot < 0 && env_int("AQE_SQL_COUNT_SHARED", 1)) mode = 2;
    if (mode == 2 && a.agg_slot >= 0 && env_int("AQE_SQL_PACKED", 1)) {
        // Packed shared bins (SqlBins MODE 3, three atomics per row) hold sums of u = fx - bias with 0 <= u < 2^62: usable whenever the
        // fixed-point values of the aggregate column span fewer than 2^62 steps -- every floating-point column scaled for its own
        // magnitude and every integer column narrower than that.  The range comes from the cached column statistics; a WHERE-bounded
        // scale (sql_layout) under which the column's extremes saturate keeps the general form.
        const aqe_db::ColStat* st;
        if ((rc = sql_col_stat(db, q->agg_col, &st))) return rc;
        __int128 flo, fhi;
        bool fits = true;
        if (a.agg_kind == K_F64) {
            const double lo = okey_f64(st->min_key) * a.sum_scale, hi = okey_f64(st->max_key) * a.sum_scale;
            fits = std::fabs(lo) < 0x1p62 && std::fabs(hi) < 0x1p62;   // (false for NaN)
            flo = fits ? (__int128)std::llrint(lo) : 0; fhi = fits ? (__int128)std::llrint(hi) : 0;
        } else {
            flo = (__int128)okey_to_i64(st->min_key); fhi = (__int128)okey_to_i64(st->max_key);
        }
        const __int128 bias = flo < 0 ? flo : 0;
        // MODE 4 = MODE 3 + the ring kernel's walk over the pass bits of sparse tiles (72 registers, three CTAs per SM); without a WHERE
        // clause or a sample filter every row passes and MODE 3 (56 registers, four CTAs per SM at K = 6) is the faster one
        if (fits && fhi >= flo && fhi - bias < ((__int128)1 << 62)) { mode = (a.n_alt > 0 || step > 1) ? 4 : 3; a.fx_bias = (long long)bias; }
    }
    cudaStream_t s = db->stream;
    if (ring) {
        if (mode == 0) rc = moments ? sql_launch_ring<0, true>(db, ra, s) : sql_launch_ring<0, false>(db, ra, s);
        else if (mode == 1) rc = moments ? sql_launch_ring<1, true>(db, ra, s) : sql_launch_ring<1, false>(db, ra, s);
        else if (mode == 2) rc = moments ? sql_launch_ring<2, true>(db, ra, s) : sql_launch_ring<2, false>(db, ra, s);
        else if (mode == 3) rc = moments ? sql_launch_ring<3, true>(db, ra, s) : sql_launch_ring<3, false>(db, ra, s);
        else rc = moments ? sql_launch_ring<4, true>(db, ra, s) : sql_launch_ring<4, false>(db, ra, s);
    } else {
        if (!strided && dense) {  // register kernel has no row-number filter: visit the progression
            a.first = (uint64_t)((step - phase) % step); a.stride = (uint64_t)step;
            a.count = a.first < db->n ? (db->n - a.first + step - 1) / step : 0;
            if (a.count == 0) return finish_without_scan(0);
        }
        if (mode == 0) rc = moments ? sql_launch_regs<0, true>(db, a, s) : sql_launch_regs<0, false>(db, a, s);
        else if (mode == 1) rc = moments ? sql_launch_regs<1, true>(db, a, s) : sql_launch_regs<1, false>(db, a, s);
        else if (mode == 2) rc = moments ? sql_launch_regs<2, true>(db, a, s) : sql_launch_regs<2, false>(db, a, s);
        else rc = moments ? sql_launch_regs<3, true>(db, a, s) : sql_launch_regs<3, false>(db, a, s);
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(db->stream));
    std::memcpy(acc, db->sql_out_host, sizeof(uint64_t) * 5 * G);
    return exchange ? aqe_exchange_check(db) : AQE_OK;
}

static int sql_scan_any(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* L, int flags, uint64_t* acc) {
    return db->group ? group_sql_scan(db, q, L, flags, acc) : sql_scan_impl(db, q, L, flags, acc);
}

extern "C" {

int aqe_sql_parse(const char* sql, int sample_percent, aqe_sql_query* out) {
    if (!sql || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    std::string err;
    const int rc = sql_parse(sql, sample_percent, *out, err);
    return rc ? fail(rc, err) : AQE_OK;
}

int aqe_sql_facts_of(aqe_db* db, const aqe_sql_query* q, aqe_sql_facts* out) {
    if (!db || !q || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    std::memset(out, 0, sizeof(*out));
    int rc = ensure_device(db);
    if (rc) return rc;
    if (db->group) return group_sql_facts(db, q, out);
    rc = sql_init(db);
    if (rc) return rc;
    out->key_min = 0; out->key_max = db->n ? 0 : -1;
    if (q->group_col != AQE_COL_NONE) {
        if (col_kind(q->group_col) < 0 || col_kind(q->group_col) == K_F64) return fail(AQE_ERR_INVALID, "GROUP BY needs an integer column");
        const aqe_db::ColStat* st;
        if ((rc = sql_col_stat(db, q->group_col, &st))) return rc;
        if (db->n) { out->key_min = okey_to_i64(st->min_key); out->key_max = okey_to_i64(st->max_key); }
    }
    if (q->agg_col != AQE_COL_NONE) {
        const int k = col_kind(q->agg_col);
        if (k < 0) return fail(AQE_ERR_INVALID, "bad aggregate column");
        out->agg_is_integer = k != K_F64;
        const aqe_db::ColStat* st;
        if ((rc = sql_col_stat(db, q->agg_col, &st))) return rc;
        if (db->n) {
            const double lo = k == K_F64 ? okey_f64(st->min_key) : (double)okey_to_i64(st->min_key);
            const double hi = k == K_F64 ? okey_f64(st->max_key) : (double)okey_to_i64(st->max_key);
            out->agg_absmax = (lo != lo || hi != hi) ? NAN : std::max(std::fabs(lo), std::fabs(hi));
        }
    }
    return AQE_OK;
}

int aqe_sql_layout_of(const aqe_sql_query* q, const aqe_sql_facts* facts, int n_shards, aqe_sql_layout* out) {
    if (!q || !facts || !out || n_shards < 1) return fail(AQE_ERR_INVALID, "bad argument");
    std::string err;
    const int rc = sql_layout(*q, facts, n_shards, *out, err);
    return rc ? fail(rc, err) : AQE_OK;
}

int aqe_sql_shifts(double agg_absmax, int agg_is_integer, int* sum_shift, int* sq_shift) {
    int a = 0, b = 0;
    std::string err;
    const int rc = sql_shifts(agg_absmax, agg_is_integer != 0, a, b, err);
    if (rc) return fail(rc, err);
    if (sum_shift) *sum_shift = a;
    if (sq_shift) *sq_shift = b;
    return AQE_OK;
}

int aqe_sql_scan(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* layout, int flags, uint64_t* acc) {
    if (!db || !q || !layout || !acc) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    return sql_scan_any(db, q, layout, flags, acc);
}

int aqe_sql_scan_exchange(aqe_db* db, const aqe_sql_query* q, const aqe_sql_layout* layout, int flags, uint64_t* acc) {
    if (!db || !q || !layout || !acc) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = ensure_device(db);
    if (rc) return rc;
    if (db->group) return group_sql_scan(db, q, layout, flags, acc);
    return sql_scan_impl(db, q, layout, flags, acc, db->ex_world > 1);
}

int aqe_sql_merge(uint64_t* acc, const uint64_t* other, uint32_t n_groups) {
    if (!acc || !other) return fail(AQE_ERR_INVALID, "NULL argument");
    sql_merge(acc, other, n_groups);
    return AQE_OK;
}

int aqe_sql_finish(const aqe_sql_query* q, int mode, const aqe_sql_layout* layout, const uint64_t* acc, const uint64_t* exists,
                   aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows) {
    if (!q || !layout || !acc) return fail(AQE_ERR_INVALID, "NULL argument");
    std::string err;
    const int rc = sql_finish(*q, mode, *layout, acc, exists, rows, cap, n_rows, err);
    return rc ? fail(rc, err) : AQE_OK;
}

int aqe_sql_execute(aqe_db* db, const aqe_sql_query* q, int mode, aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows) {
    if (!db || !q) return fail(AQE_ERR_INVALID, "NULL argument");
    aqe_sql_facts facts;
    int rc = aqe_sql_facts_of(db, q, &facts);
    if (rc) return rc;
    aqe_sql_layout L;
    if ((rc = aqe_sql_layout_of(q, &facts, 1, &L))) return rc;
    std::vector<uint64_t> acc((size_t)L.n_groups * 5), exists;
    const int flags = sql_needs_moments(*q, mode) ? AQE_SQL_MOMENTS : 0;
    if ((rc = sql_scan_any(db, q, &L, flags, acc.data()))) return rc;
    const uint64_t* ex = nullptr;
    if (q->group_col != AQE_COL_NONE && sql_sample_step(q->sample_percent) > 1) {
        // a group none of whose rows were sampled still exists for the reference (SELECT DISTINCT runs unsampled)
        bool hole = false;
        for (uint32_t g = 0; g < L.n_groups && !hole; ++g) hole = acc[(size_t)g * 5] == 0;
        if (hole) {
            exists.resize(acc.size());
            if ((rc = sql_scan_any(db, q, &L, AQE_SQL_UNSAMPLED, exists.data()))) return rc;
            ex = exists.data();
        }
    }
    return aqe_sql_finish(q, mode, &L, acc.data(), ex, rows, cap, n_rows);
}

int aqe_sql_run(aqe_db* db, const char* sql, int sample_percent, int mode, aqe_sql_row* rows, uint32_t cap, uint32_t* n_rows) {
    aqe_sql_query q;
    const int rc = aqe_sql_parse(sql, sample_percent, &q);
    if (rc) return rc;
    return aqe_sql_execute(db, &q, mode, rows, cap, n_rows);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// the table over several GPUs of this process
// ------------------------------------------------------------------------------------------------
#include "aqe_group.inl"
