// aqe_group.inl -- the record table range-sharded over several GPUs of ONE process (included at the end of aqe_engine.cu).
//
// SURVEY 8b sketched `aqe_open(path, n_gpus, &db)`, 8e "single process, 8 devices": the reference's callers
// (enhanced_aqe_cli.py:165-186, 327-346) are single-process -- `CustomBPlusDB()` -> `open_database` -> query -- so the way to give
// them the whole box is a handle that IS the sharded table.  A Group owns one plain `aqe_db` per device; shard g holds the rows
// [N g / G, N (g+1) / G) (the reference's own region split, custom_bplus_db.cpp:925-926).  A query runs as one kernel per GPU,
// launched by one host thread per GPU (Workers), and the tiny partial results meet in one of two ways:
//   * peer mode (distinct devices with peer access -- an NVSwitch box): the kernels exchange through peer-mapped mailboxes,
//     exactly the one-process-per-GPU exchange of aqe_kernels.cuh / aqe_sql_kernels.cuh, with raw peer pointers in place of
//     CUDA IPC mappings.  The result every shard ends up with is the table-level one; shard 0's pinned slot is read.
//   * colocated mode (every shard on the same device: how a 1-GPU box tests this file; kernels of one GPU must not wait on one
//     another, B200_PROFILING.md): every shard's result is merged on the host in rank order (aqe_merge_partials, aqe_sql_merge,
//     aqe_stats_merge, aqe_approx_merge) -- the same arithmetic, so both modes return the same bits for exact queries.
// Sample plans index rows of the whole table; every shard walks the whole plan and keeps the positions inside its window, so
// the gathers split across the GPUs without cutting any index list (k_plan_stats / k_plan_gather, aqe_kernels.cuh).
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>

namespace {

// One host thread per shard beyond the first (the caller's thread serves shard 0).  Workers spin briefly after a job so that a
// stream of queries finds them awake, then sleep on a condition variable.
class Workers {
public:
    explicit Workers(int n) : slots_(n > 0 ? n : 1) {
        for (int g = 1; g < (int)slots_.size(); ++g) slots_[g].th = std::thread([this, g] { loop(g); });
    }
    ~Workers() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_work_.notify_all();
        for (auto& s : slots_) if (s.th.joinable()) s.th.join();
    }
    // fn(g) for g in [0, count) side by side; the first failure (lowest g) is returned with its message in g_err.
    int run(int count, const std::function<int(int)>& fn) {
        count = std::min<int>(count, (int)slots_.size());
        if (count <= 1) return count == 1 ? fn(0) : AQE_OK;
        job_ = &fn; count_ = count;
        pending_.store((int)slots_.size() - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lock(mu_);
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_work_.notify_all();
        slots_[0].rc = fn(0);
        if (slots_[0].rc) slots_[0].err = g_err;
        for (uint32_t spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins)
            if (spins > 20000) std::this_thread::yield();
        for (int g = 0; g < count; ++g)
            if (slots_[g].rc) { g_err = slots_[g].err; return slots_[g].rc; }
        return AQE_OK;
    }

private:
    struct Slot { std::thread th; int rc = AQE_OK; std::string err; };
    void loop(int g) {
        uint64_t seen = 0;
        for (;;) {
            bool woke = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t spins = 0; !woke; ++spins) {   // poll for ~200 us after the last job, then sleep
                woke = epoch_.load(std::memory_order_acquire) != seen;
                if (woke) break;
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                if ((spins & 63u) == 63u && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
            }
            if (!woke) {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
            }
            seen = epoch_.load(std::memory_order_acquire);
            if (stop_) return;
            slots_[g].rc = AQE_OK;
            if (g < count_) {
                slots_[g].rc = (*job_)(g);
                if (slots_[g].rc) slots_[g].err = g_err;
            }
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<Slot> slots_;
    std::mutex mu_;
    std::condition_variable cv_work_;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<int> pending_{0};
    const std::function<int(int)>* job_ = nullptr;
    int count_ = 0;
    bool stop_ = false;
};

inline uint64_t shard_edge(uint64_t n, int g, int G) { return (uint64_t)(((unsigned __int128)n * (unsigned)g) / (unsigned)G); }

}  // namespace

struct Group {
    std::vector<aqe_db*> shards;   // one plain handle per device slot
    int active = 1;                // shards holding rows: ranks 0 .. active-1
    std::vector<uint64_t> first;   // active + 1 row offsets into the table
    bool peer = false;             // distinct devices, peer access on: exchanges run inside the kernels
    int ready = 0;                 // shards whose GPU has been set up (context, scratch, peer access, mailbox): done on demand, so a
                                   // handle over 8 GPUs that only ever holds a small table touches one GPU
    int64_t* perm = nullptr;       // rows of the WHOLE table ordered by amount (stratified_block_sample), on shard 0's device
    std::unique_ptr<Workers> pool;
    uint64_t rows_of(int g) const { return first[g + 1] - first[g]; }
};

// ---- exchange state: fresh mailboxes and sequence numbers for `active` ranks (after every layout change or failed exchange) ----
// Sets up the GPUs of shards [ready, want): CUDA state of the handle, peer access to and from every shard already set up, the mailbox.
static int group_prepare(aqe_db* gdb, int want) {
    Group* G = gdb->group;
    for (int g = G->ready; g < want; ++g) {
        aqe_db* c = G->shards[g];
        int rc = db_init_cuda(c);
        if (rc) return rc;
        if (G->peer) {
            for (int h = 0; h < g; ++h) {
                for (int dir = 0; dir < 2; ++dir) {
                    const int from = dir ? G->shards[h]->device : c->device, to = dir ? c->device : G->shards[h]->device;
                    CU(cudaSetDevice(from));
                    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
                }
            }
            CU(cudaSetDevice(c->device));
            if (!c->ex_mailbox) CU(cudaMalloc(&c->ex_mailbox, kMailboxBytes));
        }
        G->ready = g + 1;
    }
    return AQE_OK;
}

static int group_reset_exchange(aqe_db* gdb) {
    Group* G = gdb->group;
    if (!G->peer) return AQE_OK;
    for (int g = 0; g < G->ready; ++g) {
        aqe_db* c = G->shards[g];
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemset(c->ex_mailbox, 0, kMailboxBytes));
        c->slot_host->flags[0] = 0;
        c->ex_rank = (int)g; c->ex_world = G->active; c->ex_seq = 0; c->ax_msg = 0; c->sqlx_seq = 0;
        c->ex_total_rows = gdb->n;
        c->ex_connected = (int)g < G->active && G->active > 1;
        for (int r = 0; r < kMaxRanks; ++r) c->ex_peers[r] = r < G->ready ? G->shards[r]->ex_mailbox : nullptr;
    }
    return AQE_OK;
}

// The table now has n rows: how many shards hold them, and where they start.
static int group_set_layout(aqe_db* gdb, uint64_t n) {
    Group* G = gdb->group;
    const uint64_t min_rows = (uint64_t)std::max(1, env_int("AQE_MIN_SHARD_ROWS", 1 << 24));
    const uint64_t want = std::max<uint64_t>(1, n / min_rows);
    G->active = (int)std::min<uint64_t>(want, G->shards.size());
    G->first.assign((size_t)G->active + 1, 0);
    for (int g = 0; g <= G->active; ++g) G->first[g] = shard_edge(n, g, G->active);
    gdb->n = n;
    int rc0 = group_prepare(gdb, G->active);
    if (rc0) return rc0;
    if (G->perm) { cudaSetDevice(G->shards[0]->device); cudaFree(G->perm); G->perm = nullptr; }
    for (size_t g = (size_t)G->active; g < G->shards.size(); ++g) {   // shards left without rows give their memory back
        aqe_db* c = G->shards[g];
        if (c->cuda_ready) { CU(cudaSetDevice(c->device)); free_columns(c); }
        c->host_rows.clear(); c->host_ops.clear(); c->host_authoritative = false;
    }
    return group_reset_exchange(gdb);
}

static int group_run(aqe_db* gdb, const std::function<int(int)>& fn) {
    Group* G = gdb->group;
    return G->pool->run(G->active, [&](int g) -> int {
        cudaError_t e = cudaSetDevice(G->shards[g]->device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        return fn(g);
    });
}

static int group_close(aqe_db* gdb) {
    Group* G = gdb->group;
    if (!G) return AQE_OK;
    G->pool.reset();
    if (G->perm && !G->shards.empty()) { cudaSetDevice(G->shards[0]->device); cudaFree(G->perm); }
    for (aqe_db* c : G->shards) aqe_close(c);
    delete G;
    gdb->group = nullptr;
    return AQE_OK;
}

// ---- ingest ---------------------------------------------------------------------------------------------------------------
// rows[0, n) are the whole table in host memory: every shard uploads its range; out-of-order or repeated ids anywhere (inside a
// shard: seen on the device; across a boundary: seen here) send the table through order_like_reference (custom_bplus_db.cpp:198-200).
static int group_upload(aqe_db* gdb, const aqe_record* rows, uint64_t n, const std::vector<aqe::OrderOp>& ops = {}) {
    Group* G = gdb->group;
    int rc = group_set_layout(gdb, n);
    if (rc) return rc;
    const int per_shard = std::max(1, 8 / G->active);
    std::vector<char> unsorted((size_t)G->active, 0);
    auto upload = [&](const aqe_record* src) {
        return group_run(gdb, [&](int g) -> int {
            aqe_db* c = G->shards[g];
            int r = db_init_cuda(c);
            if (r) return r;
            c->host_rows.clear(); c->host_ops.clear(); c->host_authoritative = false;
            bool uns = false;
            const uint64_t lo = G->first[g];
            r = ingest_rows(c, G->rows_of(g), [&](aqe_record* dst, uint64_t f, uint64_t cnt) { std::memcpy(dst, src + lo + f, cnt * sizeof(aqe_record)); return true; }, &uns, per_shard);
            unsorted[g] = uns ? 1 : 0;
            return r;
        });
    };
    if ((rc = upload(rows))) return rc;
    bool bad = false;
    for (int g = 0; g < G->active; ++g) bad = bad || unsorted[g];
    for (int g = 1; g < G->active && !bad; ++g)
        if (G->rows_of(g) && G->first[g] > 0) bad = rows[G->first[g]].id <= rows[G->first[g] - 1].id;
    if (!bad) return AQE_OK;
    std::vector<aqe_record> sorted;
    if ((rc = order_like_reference(rows, n, ops, sorted))) return rc;
    return upload(sorted.data());
}

static int group_from_host_records(aqe_db* gdb, const aqe_record* rows, size_t n) {
    gdb->host_rows.clear(); gdb->host_ops.clear(); gdb->host_authoritative = false;
    return group_upload(gdb, rows, n);
}

static int group_ensure_device(aqe_db* gdb) {
    if (!gdb->host_authoritative) return group_prepare(gdb, gdb->group->active);
    const int rc = group_upload(gdb, gdb->host_rows.data(), gdb->host_rows.size(), gdb->host_ops);
    if (rc) return rc;
    gdb->host_authoritative = false;
    return AQE_OK;
}

static int group_load_file(aqe_db* gdb, const char* path, uint64_t first_row, uint64_t n_rows) {
    Group* G = gdb->group;
    RecordFile f;
    int rc = f.open(path);
    if (rc) return rc;
    if (first_row > f.total) first_row = f.total;
    const uint64_t n = std::min<uint64_t>(n_rows, f.total - first_row);
    gdb->host_rows.clear(); gdb->host_ops.clear(); gdb->host_authoritative = false;
    if ((rc = group_set_layout(gdb, n))) return rc;
    const int per_shard = std::max(1, 8 / G->active);
    std::vector<char> unsorted((size_t)G->active, 0);
    rc = group_run(gdb, [&](int g) -> int {
        aqe_db* c = G->shards[g];
        int r = db_init_cuda(c);
        if (r) return r;
        c->host_rows.clear(); c->host_ops.clear(); c->host_authoritative = false;
        bool uns = false;
        const uint64_t lo = first_row + G->first[g];
        r = ingest_rows(c, G->rows_of(g), [&](aqe_record* dst, uint64_t fr, uint64_t cnt) { return f.read_rows(dst, lo + fr, cnt); }, &uns, per_shard);
        unsorted[g] = uns ? 1 : 0;
        return r;
    });
    if (rc) return rc;
    bool bad = false;
    for (int g = 0; g < G->active; ++g) bad = bad || unsorted[g];
    for (int g = 1; g < G->active && !bad; ++g) {
        if (!G->rows_of(g) || G->first[g] == 0) continue;
        aqe_record edge[2];
        if (!f.read_rows(edge, first_row + G->first[g] - 1, 2)) return fail(AQE_ERR_IO, "short read while loading rows");
        bad = edge[1].id <= edge[0].id;
    }
    if (!bad) return AQE_OK;
    std::vector<aqe_record> rows;
    try { rows.resize(n); } catch (const std::bad_alloc&) { return fail(AQE_ERR_NOMEM, "out of host memory while ordering rows by id"); }
    if (!f.read_rows(rows.data(), first_row, n)) return fail(AQE_ERR_IO, "re-read failed");
    return group_upload(gdb, rows.data(), n);
}

static int group_generate(aqe_db* gdb, uint64_t seed, uint64_t first_row, uint64_t n_rows, int dist, uint32_t mask) {
    Group* G = gdb->group;
    gdb->host_rows.clear(); gdb->host_ops.clear(); gdb->host_authoritative = false;
    int rc = group_set_layout(gdb, n_rows);
    if (rc) return rc;
    return group_run(gdb, [&](int g) { return aqe_generate_synthetic(G->shards[g], seed, first_row + G->first[g], G->rows_of(g), dist, mask); });
}

// ---- reading rows back --------------------------------------------------------------------------------------------------
template <typename F> static int group_each_range(aqe_db* gdb, uint64_t first, uint64_t n, F f) {  // f(g, local first, count, offset into [first, first+n))
    Group* G = gdb->group;
    if (first > gdb->n || n > gdb->n - first) return fail(AQE_ERR_INVALID, "row range out of bounds");
    return group_run(gdb, [&](int g) -> int {
        const uint64_t a = std::max(first, G->first[g]), b = std::min(first + n, G->first[g + 1]);
        if (a >= b) return AQE_OK;
        return f(g, a - G->first[g], b - a, a - first);
    });
}
static int group_read_records(aqe_db* gdb, uint64_t first, uint64_t n, aqe_record* out) {
    Group* G = gdb->group;
    return group_each_range(gdb, first, n, [&](int g, uint64_t lo, uint64_t cnt, uint64_t off) { return aqe_read_records(G->shards[g], lo, cnt, out + off); });
}
static int group_read_column(aqe_db* gdb, int col, uint64_t first, uint64_t n, void* out) {
    Group* G = gdb->group;
    const int k = col_kind_of(col);
    if (k < 0) return fail(AQE_ERR_INVALID, "bad column");
    const size_t esz = k == 2 ? 4 : 8;
    return group_each_range(gdb, first, n, [&](int g, uint64_t lo, uint64_t cnt, uint64_t off) {
        return aqe_read_column(G->shards[g], col, lo, cnt, static_cast<char*>(out) + off * esz);
    });
}

// ---- exact scans ----------------------------------------------------------------------------------------------------------
static aqe_partial identity_partial() {
    aqe_partial p;
    std::memset(&p, 0, sizeof(p));
    p.minv = INFINITY; p.maxv = -INFINITY;
    return p;
}

static int group_scan_range(aqe_db* gdb, const aqe_scan_spec* spec, uint64_t first, uint64_t n, bool moments, aqe_partial* out) {
    Group* G = gdb->group;
    if (first > gdb->n || n > gdb->n - first) return fail(AQE_ERR_INVALID, "row range out of bounds");
    if (G->active == 1) {
        CU(cudaSetDevice(G->shards[0]->device));
        return scan_sync(G->shards[0], spec, first, n, moments, out);
    }
    const int ak = col_kind(spec->agg_col);
    if (ak < 0) return fail(AQE_ERR_INVALID, "bad aggregate column");
    if (spec->pred_col != AQE_COL_NONE && col_kind(spec->pred_col) < 0) return fail(AQE_ERR_INVALID, "bad predicate column");
    bool fused = G->peer && first == 0 && n == gdb->n;
    for (int g = 0; g < G->active; ++g) {   // everything a launch could refuse is checked before any kernel waits for a peer
        const aqe_db* c = G->shards[g];
        if (!col_ptr(c, spec->agg_col) && c->n) return fail(AQE_ERR_STATE, "aggregate column is not resident on the device");
        if (spec->pred_col != AQE_COL_NONE && spec->pred_col != spec->agg_col && !col_ptr(c, spec->pred_col) && c->n)
            return fail(AQE_ERR_STATE, "predicate column is not resident on the device");
        if (ak != K_F64 && c->n > (1ull << 32)) fused = false;   // integer aggregates of long shards go in segments, merged here
    }
    if (fused) {
        int rc = group_run(gdb, [&](int g) -> int {
            aqe_db* c = G->shards[g];
            int r = scan_launch(c, spec, 0, c->n, moments, &c->slot_dev->partial, c->stream, true);
            if (r) return r;
            CU(cudaStreamSynchronize(c->stream));
            if (c->slot_host->flags[0]) return fail(AQE_ERR_CUDA, "fused exchange timed out waiting for a peer shard");
            return AQE_OK;
        });
        if (rc) { const std::string keep = g_err; group_reset_exchange(gdb); g_err = keep; return rc; }
        *out = G->shards[0]->slot_host->partial;
        return AQE_OK;
    }
    std::vector<aqe_partial> parts((size_t)G->active, identity_partial());
    int rc = group_each_range(gdb, first, n, [&](int g, uint64_t lo, uint64_t cnt, uint64_t) { return scan_sync(G->shards[g], spec, lo, cnt, moments, &parts[g]); });
    if (rc) return rc;
    return aqe_merge_partials(parts.data(), G->active, ak != K_F64, out);
}

// ---- sample plans over the sharded table -------------------------------------------------------------------------------------
static int group_amount_view(aqe_db* gdb, GlobalF64* out) {
    Group* G = gdb->group;
    std::memset(out, 0, sizeof(*out));
    out->parts = G->active;
    for (int g = 0; g < G->active; ++g) {
        if (!G->shards[g]->col.amount && G->rows_of(g)) return fail(AQE_ERR_STATE, "amount column is not resident on the device");
        out->base[g] = G->shards[g]->col.amount; out->first[g] = G->first[g];
    }
    out->first[G->active] = gdb->n;
    return AQE_OK;
}

// rows of the whole table ordered by amount (stratified_block_sample sorts the table, custom_bplus_db.cpp:1343): the shards' amount
// columns are copied next to each other on shard 0's GPU and sorted there once per table version; every shard reads the
// permutation through the peer mapping.
static int group_amount_perm(aqe_db* gdb, const int64_t** perm) {
    Group* G = gdb->group;
    if (!G->perm && gdb->n) {
        aqe_db* c0 = G->shards[0];
        CU(cudaSetDevice(c0->device));
        double* all = nullptr;
        CU(cudaMalloc(&all, gdb->n * 8));
        for (int g = 0; g < G->active; ++g) {
            if (!G->shards[g]->col.amount && G->rows_of(g)) { cudaFree(all); return fail(AQE_ERR_STATE, "amount column is not resident on the device"); }
            if (G->rows_of(g)) {
                cudaError_t e = cudaMemcpyAsync(all + G->first[g], G->shards[g]->col.amount, G->rows_of(g) * 8, cudaMemcpyDefault, c0->stream);
                if (e != cudaSuccess) { cudaFree(all); return cuda_fail(e, "cudaMemcpyAsync (peer copy of a shard's amount column)"); }
            }
        }
        const int rc = sort_rows_by_amount(c0, all, gdb->n, &G->perm);
        cudaFree(all);
        if (rc) return rc;
    }
    *perm = G->perm;
    return AQE_OK;
}

static int group_stats(aqe_db* gdb, const aqe_plan* pl, int col, int pred_col, double lo, double hi, aqe_stats* out) {
    Group* G = gdb->group;
    if (G->active == 1) {
        CU(cudaSetDevice(G->shards[0]->device));
        return stats_launch(G->shards[0], pl, col, out, pred_col, lo, hi);
    }
    const int64_t* perm = nullptr;
    int rc;
    if (pl->by_amount_order && (rc = group_amount_perm(gdb, &perm))) return rc;
    std::vector<aqe_stats_partial> parts((size_t)G->active);
    rc = group_run(gdb, [&](int g) -> int {
        PlanWindow w;
        w.first = G->first[g]; w.n = G->rows_of(g); w.perm = perm;
        return stats_launch(G->shards[g], pl, col, nullptr, pred_col, lo, hi, &w, &parts[g]);
    });
    if (rc) return rc;
    return aqe_stats_merge(parts.data(), G->active, out);
}

// Every shard writes the rows of its window straight into shard 0's gather buffer (peer stores, 32 bytes a row); one D2H follows.
static int group_gather(aqe_db* gdb, const aqe_plan* pl, aqe_record* out, uint64_t cap) {
    Group* G = gdb->group;
    if (G->active == 1) {
        CU(cudaSetDevice(G->shards[0]->device));
        return gather_launch(G->shards[0], pl, out, cap);
    }
    const uint64_t n = std::min<uint64_t>(pl->count, cap);
    if (n == 0) return AQE_OK;
    const int64_t* perm = nullptr;
    int rc;
    if (pl->by_amount_order && (rc = group_amount_perm(gdb, &perm))) return rc;
    aqe_db* c0 = G->shards[0];
    CU(cudaSetDevice(c0->device));
    const uint64_t chunk = 1u << 22;
    if ((rc = ensure_gather_buf(c0, std::min<uint64_t>(n, chunk)))) return rc;
    std::vector<PlanDev> P((size_t)G->active);
    rc = group_run(gdb, [&](int g) -> int {
        const int r = plan_to_device(G->shards[g], pl, &P[g]);
        P[g].perm = perm;
        return r;
    });
    if (rc) return rc;
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - off);
        rc = group_run(gdb, [&](int g) -> int {
            aqe_db* c = G->shards[g];
            PlanWindow w;
            w.first = G->first[g]; w.n = G->rows_of(g);
            if (w.n == 0) return AQE_OK;
            const int r = gather_kernel(c, P[g], c0->gather_buf, off, cnt, w, nullptr);
            if (r) return r;
            CU(cudaStreamSynchronize(c->stream));
            return AQE_OK;
        });
        if (rc) return rc;
        CU(cudaSetDevice(c0->device));
        CU(cudaMemcpyAsync(out + off, c0->gather_buf, cnt * sizeof(aqe_record), cudaMemcpyDeviceToHost, c0->stream));
        CU(cudaStreamSynchronize(c0->stream));
    }
    return AQE_OK;
}

// ---- fused estimators -------------------------------------------------------------------------------------------------------
static int group_approx(aqe_db* gdb, const aqe_approx_spec* S, aqe_approx_result* out) {
    Group* G = gdb->group;
    if (!S || !out) return fail(AQE_ERR_INVALID, "NULL argument");
    int rc = group_ensure_device(gdb);
    if (rc) return rc;
    if (G->active == 1) {
        CU(cudaSetDevice(G->shards[0]->device));
        return approx_run(G->shards[0], S, out, false);
    }
    std::vector<aqe_approx_result> res((size_t)G->active);
    if (G->peer) {
        // shards are strata, ONE stop rule: the persistent kernels exchange their cumulative moments after every look
        rc = group_run(gdb, [&](int g) { return approx_run(G->shards[g], S, &res[g], true); });
        if (rc) { const std::string keep = g_err; group_reset_exchange(gdb); g_err = keep; return rc; }
        *out = res[0];
        return AQE_OK;
    }
    // colocated shards cannot wait on one another: every shard runs to the same relative target on its own stream of draws and
    // the estimates are merged as strata (what sharded.ShardedTable.approx does without the fused exchange)
    rc = group_run(gdb, [&](int g) {
        aqe_approx_spec sp = *S;
        sp.seed = (S->seed << 8) + (uint64_t)g;
        return approx_run(G->shards[g], &sp, &res[g], false);
    });
    if (rc) return rc;
    return aqe_approx_merge(res.data(), G->active, S->agg, S->confidence_level, out);
}

// ---- SQL-string path --------------------------------------------------------------------------------------------------------
static int group_sql_facts(aqe_db* gdb, const aqe_sql_query* q, aqe_sql_facts* out) {
    Group* G = gdb->group;
    std::vector<aqe_sql_facts> facts((size_t)G->active);
    int rc = group_run(gdb, [&](int g) { return aqe_sql_facts_of(G->shards[g], q, &facts[g]); });
    if (rc) return rc;
    // the merge aqe_sql_layout_of does over the shards' facts (aqe_sql.cpp sql_layout): one range of keys, the largest magnitude
    aqe_sql_facts m;
    std::memset(&m, 0, sizeof(m));
    m.key_min = 0; m.key_max = -1;
    bool any = false;
    for (const aqe_sql_facts& f : facts) {
        if (!(f.agg_absmax <= m.agg_absmax)) m.agg_absmax = f.agg_absmax;   // NaN propagates
        m.agg_is_integer = m.agg_is_integer || f.agg_is_integer;
        if (f.key_min > f.key_max) continue;
        if (!any) { m.key_min = f.key_min; m.key_max = f.key_max; any = true; }
        else { m.key_min = std::min(m.key_min, f.key_min); m.key_max = std::max(m.key_max, f.key_max); }
    }
    *out = m;
    return AQE_OK;
}

static int group_sql_scan(aqe_db* gdb, const aqe_sql_query* q, const aqe_sql_layout* L, int flags, uint64_t* acc) {
    Group* G = gdb->group;
    if (G->active == 1) {
        CU(cudaSetDevice(G->shards[0]->device));
        return sql_scan_impl(G->shards[0], q, L, flags, acc);
    }
    const uint32_t ng = L->n_groups;
    if (ng < 1 || ng > AQE_SQL_MAX_GROUPS) return fail(AQE_ERR_INVALID, "layout: n_groups out of range");
    std::vector<uint64_t> accs((size_t)G->active * 5 * ng);
    int rc = group_run(gdb, [&](int g) { return sql_scan_impl(G->shards[g], q, L, flags, accs.data() + (size_t)g * 5 * ng, G->peer); });
    if (rc) { if (G->peer) { const std::string keep = g_err; group_reset_exchange(gdb); g_err = keep; } return rc; }
    std::memcpy(acc, accs.data(), sizeof(uint64_t) * 5 * ng);   // peer mode: every shard already holds the table-level words
    if (!G->peer)
        for (int g = 1; g < G->active; ++g) sql_merge(acc, accs.data() + (size_t)g * 5 * ng, ng);
    return AQE_OK;
}

// ---- the C-ABI of sharded handles ------------------------------------------------------------------------------------------
extern "C" {

int aqe_create_sharded(const int* devices, int n_devices, aqe_db** out) {
    if (!out || n_devices < 0) return fail(AQE_ERR_INVALID, "bad argument");
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (n_devices == 0) { n_devices = visible; devices = nullptr; }
    if (n_devices < 1) return fail(AQE_ERR_CUDA, "no CUDA device");
    if (n_devices > kMaxRanks) return fail(AQE_ERR_INVALID, "at most 16 shards");
    std::vector<int> dev((size_t)n_devices);
    for (int g = 0; g < n_devices; ++g) {
        dev[g] = devices ? devices[g] : g;
        if (dev[g] < 0 || dev[g] >= visible) return fail(AQE_ERR_CUDA, "no such CUDA device " + std::to_string(dev[g]));
    }
    if (n_devices == 1) return aqe_create(dev[0], out);
    bool same = true, distinct = true;
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b) { same = same && dev[a] == dev[b]; distinct = distinct && dev[a] != dev[b]; }
    if (!same && !distinct) return fail(AQE_ERR_INVALID, "the shards' devices must be all different (peer mode) or all the same (colocated)");
    if (distinct)
        for (int a = 0; a < n_devices; ++a)
            for (int b = 0; b < n_devices; ++b) {
                int can = 0;
                if (a != b) { CU(cudaDeviceCanAccessPeer(&can, dev[a], dev[b])); if (!can) return fail(AQE_ERR_UNSUPPORTED, "devices " + std::to_string(dev[a]) + " and " + std::to_string(dev[b]) + " have no peer access"); }
            }
    aqe_db* gdb = new (std::nothrow) aqe_db();
    Group* G = new (std::nothrow) Group();
    if (!gdb || !G) { delete gdb; delete G; return fail(AQE_ERR_NOMEM, "out of host memory"); }
    gdb->group = G; gdb->device = dev[0];
    G->peer = distinct;
    auto bail = [&](int rc) { const std::string keep = g_err; aqe_close(gdb); g_err = keep; return rc; };
    for (int g = 0; g < n_devices; ++g) {   // plain handles; their GPUs are set up when a table is large enough to reach them (group_prepare)
        aqe_db* c = nullptr;
        const int rc = aqe_create(dev[g], &c);
        if (rc) return bail(rc);
        G->shards.push_back(c);
    }
    G->pool.reset(new Workers(n_devices));
    G->active = 1; G->first.assign(2, 0);   // an empty table on shard 0; no CUDA call until rows arrive or a query runs (like aqe_create)
    *out = gdb;
    return AQE_OK;
}

int aqe_open_sharded(const char* path, int n_gpus, aqe_db** out) {
    int rc = aqe_create_sharded(nullptr, n_gpus, out);
    if (rc) return rc;
    rc = aqe_load_file(*out, path, 0, UINT64_MAX);
    if (rc) { const std::string keep = g_err; aqe_close(*out); *out = nullptr; g_err = keep; }
    return rc;
}

int aqe_shard_count(const aqe_db* db) { return db ? (db->group ? db->group->active : 1) : 0; }
aqe_db* aqe_shard(aqe_db* db, int g) {
    if (!db) return nullptr;
    if (!db->group) return g == 0 ? db : nullptr;
    return g >= 0 && g < (int)db->group->shards.size() ? db->group->shards[g] : nullptr;
}
uint64_t aqe_shard_first_row(const aqe_db* db, int g) {
    if (!db) return 0;
    if (!db->group) return g <= 0 ? 0 : aqe_count(db);
    const Group* G = db->group;
    return g <= 0 ? 0 : (g >= G->active ? db->n : G->first[g]);
}
int aqe_shards_fused(const aqe_db* db) { return db && db->group && db->group->peer && db->group->active > 1 ? 1 : 0; }

}  // extern "C"
