// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A C-ABI wrapper that drives the UNMODIFIED reference `CustomBPlusDB` / `CustomApproximateScheduler`
// (compiled in place from /root/reference/src/aqe_backend/core/*.cpp by oracle/Makefile into
// oracle/_ref/libaqe_ref.so).  It is used to (1) validate the C restatement in oracle/aqe_oracle.c,
// (2) mint the golden vectors under tests/golden/, (3) serve as the `cpu_baseline.kind="reference"`
// leg of bench.py.  The reference's open_database()/load_from_file() self-deadlock
// (custom_bplus_db.cpp:689 + :165), so rows are loaded through the public insert_batch() (:196).
#include "custom_bplus_db.hpp"
#include "custom_scheduler.hpp"
#include "aqe_b200.h"

#include <chrono>
#include <cstring>
#include <vector>

static_assert(sizeof(Record) == sizeof(aqe_record), "Record must be the 32-byte row");

namespace {
int64_t emit(const std::vector<Record>& v, aqe_record* out, uint64_t cap) {
    uint64_t n = v.size() < cap ? v.size() : cap;
    if (out && n) std::memcpy(out, v.data(), n * sizeof(Record));
    return (int64_t)v.size();
}
}  // namespace

extern "C" {

void* ref_new() { return new CustomBPlusDB(); }
void ref_free(void* h) { delete static_cast<CustomBPlusDB*>(h); }

int ref_insert_batch(void* h, const aqe_record* rows, size_t n) {
    std::vector<Record> v(n);
    if (n) std::memcpy(v.data(), rows, n * sizeof(Record));
    return static_cast<CustomBPlusDB*>(h)->insert_batch(v) ? 0 : 1;
}
int ref_insert_record(void* h, const aqe_record* row) {
    Record r(row->id, row->amount, row->region, row->product_id, row->timestamp);
    return static_cast<CustomBPlusDB*>(h)->insert_record(r) ? 0 : 1;
}
int ref_save(void* h, const char* path) { return static_cast<CustomBPlusDB*>(h)->save_to_file(path) ? 0 : 1; }

uint64_t ref_total(void* h) { return static_cast<CustomBPlusDB*>(h)->get_total_records(); }
uint64_t ref_node_count(void* h) { return static_cast<CustomBPlusDB*>(h)->get_node_count(); }
uint64_t ref_tree_height(void* h) { return static_cast<CustomBPlusDB*>(h)->get_tree_height(); }
double ref_sum_amount(void* h) { return static_cast<CustomBPlusDB*>(h)->sum_amount(); }
double ref_avg_amount(void* h) { return static_cast<CustomBPlusDB*>(h)->avg_amount(); }
double ref_sum_amount_where(void* h, double lo, double hi) {
    return static_cast<CustomBPlusDB*>(h)->sum_amount_where(lo, hi);
}
double ref_fast_aggregated(void* h, double p, int threads) {
    return static_cast<CustomBPlusDB*>(h)->fast_aggregated_memory_stride_sum(p, threads);
}
double ref_parallel_sum_sample(void* h, double p, int threads) {
    return static_cast<CustomBPlusDB*>(h)->parallel_sum_sample(p, threads);
}

// Runs sampler `method` (ids = aqe_method) and copies up to `cap` rows to `out`; returns the number of
// rows the reference returned.
int64_t ref_sample(void* h, int method, const aqe_sample_params* p, aqe_record* out, uint64_t cap) {
    CustomBPlusDB& db = *static_cast<CustomBPlusDB*>(h);
    const double sp = p->sample_percent;
    switch (method) {
        case AQE_M_SLOW_POINTER: return emit(db.slow_pointer_sample(sp), out, cap);
        case AQE_M_FAST_POINTER: return emit(db.fast_pointer_sample(sp, (int)p->step_size), out, cap);
        case AQE_M_DUAL_POINTER: return emit(db.dual_pointer_sample(sp), out, cap);
        case AQE_M_PARALLEL_POINTER: return emit(db.parallel_pointer_sample(sp, (int)p->num_threads), out, cap);
        case AQE_M_RANDOM_POINTER: return emit(db.random_pointer_sample(sp, (unsigned)p->seed), out, cap);
        case AQE_M_MEMORY_STRIDE: return emit(db.memory_stride_sample(sp, (size_t)p->block_size), out, cap);
        case AQE_M_OPT_ADDRESS_ARITHMETIC: return emit(db.optimized_address_arithmetic_sample(sp), out, cap);
        case AQE_M_INDEX_BASED: return emit(db.index_based_sample(sp), out, cap);
        case AQE_M_BYTE_OFFSET: return emit(db.byte_offset_sample(sp), out, cap);
        case AQE_M_OPTIMIZED_CLT:
            return emit(db.optimized_clt_sample(sp, p->confidence_level, (int)p->check_interval,
                                                (int)p->num_threads, p->max_error_percent), out, cap);
        case AQE_M_BLOCK: return emit(db.block_sample(sp, (size_t)p->block_size), out, cap);
        case AQE_M_PAGE: return emit(db.page_sample(sp, (size_t)p->block_size), out, cap);
        case AQE_M_PARALLEL_BLOCK:
            return emit(db.parallel_block_sample(sp, (size_t)p->block_size, (int)p->num_threads), out, cap);
        case AQE_M_NODE_SKIP: return emit(db.node_skip_sample(sp, (int)p->step_size), out, cap);
        case AQE_M_BALANCED_TREE: return emit(db.balanced_tree_sample(sp), out, cap);
        case AQE_M_DIRECT_ACCESS: return emit(db.direct_access_sample(sp), out, cap);
        case AQE_M_ADAPTIVE_BLOCK:
            return emit(db.adaptive_block_sample(sp, (size_t)p->block_size, (size_t)p->block_size_max), out, cap);
        case AQE_M_STRATIFIED_BLOCK:
            return emit(db.stratified_block_sample(sp, (size_t)p->block_size, (int)p->block_size_max), out, cap);
        case AQE_M_SAMPLE_RECORDS: return emit(db.sample_records(sp), out, cap);
        case AQE_M_OPTIMIZED_SEQUENTIAL: return emit(db.optimized_sequential_sample(sp), out, cap);
        case AQE_M_RANDOM_START_NTH: return emit(db.random_start_nth_sample(sp, (int)p->step_size), out, cap);
        case AQE_M_ADDRESS_ARITHMETIC: return emit(db.address_arithmetic_sample(sp), out, cap);
        case AQE_M_RANDOM_START_MEMORY_STRIDE:
            return emit(db.random_start_memory_stride_sample(sp, (size_t)p->block_size), out, cap);
        case AQE_M_MULTITHREADED_MEMORY_STRIDE:
            return emit(db.multithreaded_memory_stride_sample(sp, (int)p->num_threads), out, cap);
        case AQE_M_CLT_VALIDATED_DUAL_POINTER:
            return emit(db.clt_validated_dual_pointer_sample(sp, p->confidence_level, (int)p->check_interval,
                                                             (int)p->num_threads, p->max_error_percent), out, cap);
        case AQE_M_SIGNAL_BASED_CLT: return emit(db.signal_based_clt_sample(sp, (int)p->check_interval), out, cap);
        default: return -1;
    }
}

// Timing helper for the CPU baseline: median-free, returns seconds for `reps` back-to-back calls.
// what: 0 sum_amount, 1 sum_amount_where(lo,hi), 2 memory_stride_sample(p), 3 clt_validated(p,0.95,10,4,err),
//       4 block_sample(p,1000)
double ref_time(void* h, int what, int reps, double a, double b, double* last_value) {
    CustomBPlusDB& db = *static_cast<CustomBPlusDB*>(h);
    double v = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        switch (what) {
            case 0: v = db.sum_amount(); break;
            case 1: v = db.sum_amount_where(a, b); break;
            case 2: v = (double)db.memory_stride_sample(a, 0).size(); break;
            case 3: v = (double)db.clt_validated_dual_pointer_sample(a, 0.95, 10, 4, b).size(); break;
            case 4: v = (double)db.block_sample(a, 1000).size(); break;
            default: break;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (last_value) *last_value = v;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- CustomApproximateScheduler (custom_scheduler.cpp) -------------------------------------------------
struct ref_sched_result {
    double value; int status; double confidence_level; double error_margin; int samples_used; double ms;
};
void* ref_sched_new(double thr) { return new CustomApproximateScheduler(thr); }
void ref_sched_free(void* h) { delete static_cast<CustomApproximateScheduler*>(h); }
int ref_sched_insert_batch(void* h, const aqe_record* rows, size_t n) {
    std::vector<Record> v(n);
    if (n) std::memcpy(v.data(), rows, n * sizeof(Record));
    return static_cast<CustomApproximateScheduler*>(h)->insert_batch(v) ? 0 : 1;
}
static void fill(const CustomValidationResult& r, ref_sched_result* o) {
    o->value = r.value; o->status = (int)r.status; o->confidence_level = r.confidence_level;
    o->error_margin = r.error_margin; o->samples_used = r.samples_used; o->ms = (double)r.computation_time.count();
}
// what: 0 exact_sum 1 exact_avg 2 exact_count 3 sum_query 4 avg_query 5 count_query
int ref_sched_exec(void* h, int what, const char* query, double p, int threads, ref_sched_result* out) {
    auto& s = *static_cast<CustomApproximateScheduler*>(h);
    switch (what) {
        case 0: fill(s.execute_exact_sum(), out); return 0;
        case 1: fill(s.execute_exact_avg(), out); return 0;
        case 2: fill(s.execute_exact_count(), out); return 0;
        case 3: fill(s.execute_sum_query(query, p, threads), out); return 0;
        case 4: fill(s.execute_avg_query(query, p, threads), out); return 0;
        case 5: fill(s.execute_count_query(query, p, threads), out); return 0;
        default: return 1;
    }
}
double ref_sched_size_mb(void* h) { return static_cast<CustomApproximateScheduler*>(h)->get_database_size_mb(); }

}  // extern "C"
