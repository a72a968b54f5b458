// aqe_pybind.cpp -- the `aqe_backend` CPython module: a thin pybind11 shim over the C-ABI of
// libaqe_b200.so with the Python-level surface of the reference's module
// (reference: src/aqe_backend/bindings/bindings.cpp:10-137 -- same class names, method names, argument
// names and defaults, return types and the GIL-held, synchronous call model).
//
// Differences a caller can observe, all deliberate (DESIGN.md "Boundary"):
//   * open_database()/load_from_file() work (the reference self-deadlocks, custom_bplus_db.cpp:689+165);
//   * arguments for which the reference divides by zero or never terminates raise ValueError;
//   * CUDA failures raise RuntimeError; there is no CPU fallback;
//   * additive methods: approx_sum/approx_avg/approx_count (fused persistent-kernel estimators with a
//     correct confidence interval), sum_column, sum_where, sample_stats, *_array numpy returns,
//     attach_columns / column_ptr (torch hand-off), generate_synthetic.
#include <pybind11/chrono.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "aqe_b200.h"

#include <structmember.h>

namespace py = pybind11;

struct Record {  // custom_bplus_db.hpp:17-27
    int64_t id = 0;
    double amount = 0.0;
    int32_t region = 0;
    int32_t product_id = 0;
    int64_t timestamp = 0;
};
static_assert(sizeof(Record) == sizeof(aqe_record), "Record is the 32-byte row");

// `Record` as a plain CPython type (five C members, default constructor, read/write attributes -- bindings.cpp:14-20) instead of a
// pybind11 class: every sampler returns list[Record] BY VALUE and the reference's callers loop over the objects in Python
// (enhanced_aqe_cli.py:188-200), so creating them is the cost of a sampled query (round 1: 100 k records 23.7 ms, 99 % of it pybind11
// instance bookkeeping at ~230 ns per object).  A type without a __dict__, GC header or holder costs one small allocation.
struct RecordObject {
    PyObject_HEAD
    Record r;
};
static PyTypeObject* g_record_type = nullptr;

static PyObject* record_new(PyTypeObject* type, PyObject*, PyObject*) {
    RecordObject* o = reinterpret_cast<RecordObject*>(type->tp_alloc(type, 0));
    if (o) o->r = Record{};
    return reinterpret_cast<PyObject*>(o);
}
static int record_init(PyObject*, PyObject* args, PyObject* kwargs) {
    if (PyTuple_GET_SIZE(args) != 0 || (kwargs && PyDict_Size(kwargs) != 0)) {
        PyErr_SetString(PyExc_TypeError, "Record() takes no arguments");   // bindings.cpp:15: py::init<>()
        return -1;
    }
    return 0;
}
static PyObject* record_repr(PyObject* self) {
    const Record& r = reinterpret_cast<RecordObject*>(self)->r;
    const std::string t = "Record(id=" + std::to_string(r.id) + ", amount=" + std::to_string(r.amount) + ", region=" + std::to_string(r.region) +
                          ", product_id=" + std::to_string(r.product_id) + ", timestamp=" + std::to_string(r.timestamp) + ")";
    return PyUnicode_FromStringAndSize(t.data(), (Py_ssize_t)t.size());
}
static PyMemberDef record_members[] = {
    {"id", T_LONGLONG, offsetof(RecordObject, r) + offsetof(Record, id), 0, nullptr},
    {"amount", T_DOUBLE, offsetof(RecordObject, r) + offsetof(Record, amount), 0, nullptr},
    {"region", T_INT, offsetof(RecordObject, r) + offsetof(Record, region), 0, nullptr},
    {"product_id", T_INT, offsetof(RecordObject, r) + offsetof(Record, product_id), 0, nullptr},
    {"timestamp", T_LONGLONG, offsetof(RecordObject, r) + offsetof(Record, timestamp), 0, nullptr},
    {nullptr, 0, 0, 0, nullptr}};
static PyType_Slot record_slots[] = {{Py_tp_new, reinterpret_cast<void*>(record_new)},   {Py_tp_init, reinterpret_cast<void*>(record_init)},
                                     {Py_tp_repr, reinterpret_cast<void*>(record_repr)}, {Py_tp_members, record_members},
                                     {Py_tp_doc, const_cast<char*>("fixed-width row {id, amount, region, product_id, timestamp} (custom_bplus_db.hpp:17-27)")},
                                     {0, nullptr}};
static PyType_Spec record_spec = {"aqe_backend.Record", (int)sizeof(RecordObject), 0, Py_TPFLAGS_DEFAULT, record_slots};

namespace pybind11 { namespace detail {
template <> struct type_caster<Record> {
    PYBIND11_TYPE_CASTER(Record, const_name("Record"));
    bool load(handle src, bool) {
        if (!src || !g_record_type || !PyObject_TypeCheck(src.ptr(), g_record_type)) return false;
        value = reinterpret_cast<RecordObject*>(src.ptr())->r;
        return true;
    }
    static handle cast(const Record& r, return_value_policy, handle) {
        RecordObject* o = reinterpret_cast<RecordObject*>(g_record_type->tp_alloc(g_record_type, 0));
        if (!o) throw error_already_set();
        o->r = r;
        return reinterpret_cast<PyObject*>(o);
    }
};
}}  // namespace pybind11::detail

namespace {

enum class CustomApproximationStatus { STABLE, DRIFTING, INSUFFICIENT_DATA, ERROR };  // custom_scheduler.hpp:8-13

struct CustomValidationResult {  // custom_scheduler.hpp:15-22
    double value = 0.0;
    CustomApproximationStatus status = CustomApproximationStatus::ERROR;
    double confidence_level = 0.0;
    double error_margin = 100.0;
    int samples_used = 0;
    std::chrono::milliseconds computation_time{0};
};

struct QueryResult { double value, ci_lower, ci_upper; };  // executor.h:8-12

// Result of the fused estimators: the reference's result fields plus the interval.
struct ApproxResult {
    double value = 0.0, ci_lower = 0.0, ci_upper = 0.0;
    CustomApproximationStatus status = CustomApproximationStatus::ERROR;
    double confidence_level = 0.0, error_margin = 0.0;
    uint64_t samples_used = 0, population = 0;
    uint32_t rounds = 0;
    double kernel_us = 0.0;
    std::chrono::microseconds computation_time{0};
};

struct BenchmarkResults {  // custom_scheduler.hpp:68-77 (unbound in the reference: calling it raised TypeError)
    double exact_value, approximate_value, exact_time_ms, approximate_time_ms, speedup, error_percentage;
    int threads_used;
    double sample_percentage;
};

[[noreturn]] void raise_status(int rc) {
    const std::string msg = aqe_last_error();
    if (rc == AQE_ERR_INVALID) throw py::value_error(msg);
    throw std::runtime_error(msg);
}
inline void check(int rc) { if (rc != AQE_OK) raise_status(rc); }

// Which GPUs a CustomBPlusDB() uses.  AQE_DEVICES = "all" | a count | a comma-separated list chooses explicitly; a process that
// is one rank of a one-process-per-GPU job (AQE_DEVICE / LOCAL_RANK set, e.g. under torchrun) keeps to its own GPU; otherwise
// the table is range-sharded over every visible GPU (aqe_create_sharded) -- the reference's callers are single-process
// (enhanced_aqe_cli.py:165-186), this is how they get the whole box.
std::vector<int> default_devices() {
    auto visible = [] { int n = 0; return aqe_device_count(&n) == AQE_OK ? n : 0; };
    if (const char* v = std::getenv("AQE_DEVICES"); v && *v) {
        std::vector<int> out;
        const std::string s(v);
        if (s == "all") { for (int g = 0; g < visible(); ++g) out.push_back(g); }
        else if (s.find(',') == std::string::npos) { for (int g = 0; g < std::atoi(v); ++g) out.push_back(g); }
        else {
            size_t i = 0;
            while (i < s.size()) { size_t j = s.find(',', i); if (j == std::string::npos) j = s.size(); if (j > i) out.push_back(std::atoi(s.substr(i, j - i).c_str())); i = j + 1; }
        }
        if (!out.empty()) return out;
    }
    for (const char* name : {"AQE_DEVICE", "LOCAL_RANK"}) {
        const char* v = std::getenv(name);
        if (v && *v) return {std::atoi(v)};
    }
    std::vector<int> out;
    for (int g = 0; g < std::min(visible(), 16); ++g) out.push_back(g);
    if (out.empty()) out.push_back(0);   // no driver: the handle is created lazily and the first data-path call raises
    return out;
}

int column_id(const std::string& name) {
    static const std::map<std::string, int> m = {{"id", AQE_COL_ID}, {"amount", AQE_COL_AMOUNT}, {"region", AQE_COL_REGION},
                                                 {"product_id", AQE_COL_PRODUCT_ID}, {"timestamp", AQE_COL_TIMESTAMP}};
    auto it = m.find(name);
    if (it == m.end()) throw py::value_error("unknown column '" + name + "'");
    return it->second;
}

class CustomBPlusDB {
public:
    CustomBPlusDB() : devices_(default_devices()) { open_handle(); }
    explicit CustomBPlusDB(int device) : devices_{device} { open_handle(); }
    explicit CustomBPlusDB(std::vector<int> devices) : devices_(std::move(devices)) {
        if (devices_.empty()) throw py::value_error("devices: give at least one device id");
        open_handle();
    }
    ~CustomBPlusDB() {
        try { close_database(); } catch (...) {}
        aqe_close(h_);
    }
    CustomBPlusDB(const CustomBPlusDB&) = delete;
    CustomBPlusDB& operator=(const CustomBPlusDB&) = delete;

    // ---- lifecycle (custom_bplus_db.cpp:135-162, 665-711) ----
    bool create_database(const std::string& path) {
        reset();
        path_ = path;
        return true;
    }
    bool open_database(const std::string& path) { return load_from_file(path); }
    void close_database() {  // rewrites the file iff create_database() named one (cbd:157-162)
        if (!path_.empty()) save_to_file(path_);
    }
    bool save_to_file(const std::string& path) {
        const int rc = aqe_save_file(h_, path.c_str());
        if (rc == AQE_ERR_IO) return false;
        check(rc);
        return true;
    }
    bool load_from_file(const std::string& path) {
        const int rc = aqe_load_file(h_, path.c_str(), 0, UINT64_MAX);
        if (rc == AQE_ERR_IO) return false;
        check(rc);
        return true;
    }
    bool load_shard(const std::string& path, uint64_t first_row, uint64_t n_rows) {
        const int rc = aqe_load_file(h_, path.c_str(), first_row, n_rows);
        if (rc == AQE_ERR_IO) return false;
        check(rc);
        return true;
    }
    bool insert_record(const Record& r) {
        check(aqe_append_records(h_, reinterpret_cast<const aqe_record*>(&r), 1));
        return true;
    }
    bool insert_batch(const std::vector<Record>& rows) {
        check(aqe_append_records(h_, reinterpret_cast<const aqe_record*>(rows.data()), rows.size()));
        return true;
    }
    void insert_array(py::array rows) {
        py::buffer_info b = rows.request();
        if (b.itemsize != 32 || b.ndim != 1 || (b.strides[0] != 32 && b.shape[0] > 1)) throw py::value_error("need a contiguous 1-d array of 32-byte records");
        check(aqe_append_records(h_, static_cast<const aqe_record*>(b.ptr), (size_t)b.shape[0]));
    }
    void from_array(py::array rows) {
        py::buffer_info b = rows.request();
        if (b.itemsize != 32 || b.ndim != 1 || (b.strides[0] != 32 && b.shape[0] > 1)) throw py::value_error("need a contiguous 1-d array of 32-byte records");
        check(aqe_from_host_records(h_, static_cast<const aqe_record*>(b.ptr), (size_t)b.shape[0]));
    }
    void generate_synthetic(uint64_t n_rows, uint64_t seed, uint64_t first_row, int dist, uint32_t columns_mask) {
        check(aqe_generate_synthetic(h_, seed, first_row, n_rows, dist, columns_mask));
    }
    // borrowed columns live on ONE GPU: a handle that was going to shard over several becomes a handle on `device`
    void single_device(int device) {
        if (devices_.size() == 1 && devices_[0] == device) return;
        devices_.assign(1, device);
        reset();
    }
    void attach_columns(uintptr_t id, uintptr_t amount, uintptr_t region, uintptr_t product_id, uintptr_t timestamp, uint64_t n) {
        if (devices_.size() > 1) single_device(devices_[0]);
        check(aqe_attach_device_columns(h_, reinterpret_cast<const int64_t*>(id), reinterpret_cast<const double*>(amount),
                                        reinterpret_cast<const int32_t*>(region), reinterpret_cast<const int32_t*>(product_id),
                                        reinterpret_cast<const int64_t*>(timestamp), n));
    }
    // Torch hand-off without a torch dependency: any object with data_ptr() / numel() / dtype / is_cuda (CUDA tensors).
    void from_torch(py::object id, py::object amount, py::object region, py::object product_id, py::object timestamp) {
        uint64_t n = 0;
        bool have_n = false;
        int tensor_device = 0;
        auto ptr_of = [&](py::object t, const char* name, const char* want) -> uintptr_t {
            if (t.is_none()) return 0;
            if (!py::hasattr(t, "data_ptr")) throw py::value_error(std::string(name) + ": expected a CUDA tensor");
            if (!t.attr("is_cuda").cast<bool>()) throw py::value_error(std::string(name) + ": tensor must live on the GPU");
            if (!t.attr("is_contiguous")().cast<bool>()) throw py::value_error(std::string(name) + ": tensor must be contiguous");
            const std::string dt = py::str(t.attr("dtype")).cast<std::string>();
            if (dt != want) throw py::value_error(std::string(name) + ": dtype must be " + want + ", got " + dt);
            const uint64_t m = t.attr("numel")().cast<uint64_t>();
            if (have_n && m != n) throw py::value_error("all columns must have the same length");
            const py::object index = t.attr("device").attr("index");
            const int dev = index.is_none() ? 0 : index.cast<int>();
            if (have_n && dev != tensor_device) throw py::value_error("all columns must live on the same GPU");
            tensor_device = dev;
            n = m; have_n = true;
            return t.attr("data_ptr")().cast<uintptr_t>();
        };
        const uintptr_t a = ptr_of(id, "id", "torch.int64"), b = ptr_of(amount, "amount", "torch.float64"), c = ptr_of(region, "region", "torch.int32"),
                        d = ptr_of(product_id, "product_id", "torch.int32"), e = ptr_of(timestamp, "timestamp", "torch.int64");
        if (!have_n) throw py::value_error("from_torch: give at least one column");
        single_device(tensor_device);
        attach_columns(a, b, c, d, e, n);
        keep_alive_ = py::make_tuple(id, amount, region, product_id, timestamp);  // the engine borrows the memory
    }
    uintptr_t column_ptr(const std::string& col) { return reinterpret_cast<uintptr_t>(aqe_column_device_ptr(h_, column_id(col))); }

    // ---- exact (cbd:242-274, 646-658) ----
    double sum_amount() { double v = 0; check(aqe_sum_f64(h_, AQE_COL_AMOUNT, &v)); return v; }
    double sum_amount_where(double lo, double hi) { double v = 0; check(aqe_sum_where_f64(h_, AQE_COL_AMOUNT, lo, hi, &v, nullptr)); return v; }
    double avg_amount() { const uint64_t n = aqe_count(h_); return n ? sum_amount() / (double)n : 0.0; }
    size_t count_records() const { return aqe_count(h_); }
    size_t get_total_records() const { return aqe_count(h_); }
    size_t get_node_count() const { return aqe_node_count(h_); }
    size_t get_tree_height() const { return aqe_tree_height(h_); }
    py::object sum_column(const std::string& col) {  // integer columns: exact Python int (int128)
        const int c = column_id(col);
        if (c == AQE_COL_AMOUNT) return py::float_(sum_amount());
        uint64_t lo = 0; int64_t hi = 0;
        check(aqe_sum_i128(h_, c, &lo, &hi));
        py::int_ r = py::int_(hi);
        return (r * py::int_(1).attr("__lshift__")(64)) + py::int_(lo);
    }
    py::dict scan(const std::string& agg_col, py::object pred_col, double lo, double hi) {
        aqe_scan_spec sp{column_id(agg_col), pred_col.is_none() ? AQE_COL_NONE : column_id(pred_col.cast<std::string>()), lo, hi};
        aqe_partial p;
        std::memset(&p, 0, sizeof(p));
        if (aqe_count(h_)) check(aqe_scan(h_, &sp, &p));
        py::dict d;
        d["count"] = p.count; d["sum"] = p.sum; d["comp"] = p.comp; d["sumsq"] = p.sumsq; d["min"] = p.minv; d["max"] = p.maxv;
        d["isum"] = (py::int_(p.isum_hi) * py::int_(1).attr("__lshift__")(64)) + py::int_(p.isum_lo);
        return d;
    }

    // ---- samplers (bindings.cpp:49-101): list[Record] by value ----
    std::vector<Record> sample(int method, const aqe_sample_params& p) {
        aqe_plan* pl = nullptr;
        check(aqe_plan_build(h_, aqe_count(h_), method, &p, &pl));
        std::vector<Record> out(aqe_plan_count(pl));
        const int rc = out.empty() ? AQE_OK : aqe_gather_plan(h_, pl, reinterpret_cast<aqe_record*>(out.data()), out.size());
        aqe_plan_free(pl);
        check(rc);
        return out;
    }
    py::array sample_array(int method, const aqe_sample_params& p) {
        aqe_plan* pl = nullptr;
        check(aqe_plan_build(h_, aqe_count(h_), method, &p, &pl));
        const size_t n = aqe_plan_count(pl);
        const py::dtype dt = record_dtype();
        py::array out(dt, std::vector<py::ssize_t>{(py::ssize_t)n}, std::vector<py::ssize_t>{32});
        if (out.itemsize() != 32 || out.nbytes() != (py::ssize_t)(n * 32)) throw std::runtime_error("record array allocation failed");
        const int rc = n ? aqe_gather_plan(h_, pl, static_cast<aqe_record*>(out.mutable_data()), n) : AQE_OK;
        aqe_plan_free(pl);
        check(rc);
        return out;
    }
    py::dict sample_stats(int method, const aqe_sample_params& p, const std::string& col) {
        aqe_plan* pl = nullptr;
        check(aqe_plan_build(h_, aqe_count(h_), method, &p, &pl));
        aqe_stats s{};
        const int rc = aqe_stats_from_plan(h_, pl, column_id(col), &s);
        aqe_plan_free(pl);
        check(rc);
        py::dict d;
        d["n"] = s.n; d["mean"] = s.mean; d["m2"] = s.m2; d["sum"] = s.sum;
        return d;
    }
    static py::dtype record_dtype() {
        py::list names, formats, offsets;
        for (const char* n : {"id", "amount", "region", "product_id", "timestamp"}) names.append(n);
        for (const char* f : {"<i8", "<f8", "<i4", "<i4", "<i8"}) formats.append(f);
        for (int o : {0, 8, 16, 20, 24}) offsets.append(o);
        return py::dtype(names, formats, offsets, 32);
    }
    static aqe_sample_params params(int method, double pct) {
        aqe_sample_params p;
        aqe_sample_params_default(&p, method);
        p.sample_percent = pct;
        return p;
    }
    uint64_t session_seed() {  // stands in for std::random_device: differs per call unless seeded
        if (fixed_seed_) return seed_;
        return seed_ = seed_ * 6364136223846793005ull + 1442695040888963407ull;
    }
    void set_seed(py::object s) {
        if (s.is_none()) { fixed_seed_ = false; return; }
        fixed_seed_ = true; seed_ = s.cast<uint64_t>();
    }

    double fast_aggregated_memory_stride_sum(double pct, int threads) {
        aqe_sample_params p = params(AQE_M_MULTITHREADED_MEMORY_STRIDE, pct);
        p.num_threads = threads; p.seed = session_seed();
        double s = 0; uint64_t n = 0;
        check(aqe_fast_aggregated_sum(h_, &p, &s, &n));
        return s;
    }

    // ---- fused estimators (K4) ----
    ApproxResult approx(int agg, double error_percent, double confidence, const std::string& design, py::object seed,
                        py::object where, const std::string& where_col, uint32_t block_size, uint64_t min_samples, uint64_t max_samples) {
        const auto t0 = std::chrono::steady_clock::now();
        aqe_approx_spec sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.agg = agg;
        if (design == "srs") sp.design = AQE_DESIGN_SRS; else if (design == "block") sp.design = AQE_DESIGN_BLOCK; else throw py::value_error("design: 'srs' or 'block'");
        sp.agg_col = AQE_COL_AMOUNT;
        sp.pred_col = AQE_COL_NONE;
        if (!where.is_none()) {
            auto t = where.cast<std::pair<double, double>>();
            sp.pred_col = column_id(where_col); sp.lo = t.first; sp.hi = t.second;
        }
        sp.error_percent = error_percent; sp.confidence_level = confidence;
        sp.seed = seed.is_none() ? session_seed() : seed.cast<uint64_t>();
        sp.block_size = block_size; sp.min_samples = min_samples; sp.max_samples = max_samples;
        aqe_approx_result r;
        check(aqe_approx(h_, &sp, &r));
        ApproxResult o;
        o.value = r.estimate; o.ci_lower = r.ci_lower; o.ci_upper = r.ci_upper;
        o.status = static_cast<CustomApproximationStatus>(r.status);
        o.confidence_level = r.confidence_level; o.error_margin = r.error_margin;
        o.samples_used = r.n_samples; o.population = r.population; o.rounds = r.rounds; o.kernel_us = r.elapsed_us;
        o.computation_time = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
        return o;
    }

    aqe_db* handle() { return h_; }
    int shard_count() const { return aqe_shard_count(h_); }
    bool shards_fused() const { return aqe_shards_fused(h_) != 0; }
    std::vector<int> devices() const { return devices_; }

    // ---- SQL-string path on this table (the four run_query* functions of bindings.cpp:126-136 without the file) ----
    // grouped = the *_groupby forms; ci: 0 value, 1 the reference's interval (executor.cpp:177-338), 2 corrected interval
    std::vector<aqe_sql_row> sql(const std::string& query, int sample_percent, bool grouped, int ci) {
        // every failure of the reference's SQL path is a std::runtime_error (parser.cpp:32, :64, :72; core/db.cpp:40)
        auto check = [](int rc) { if (rc != AQE_OK) throw std::runtime_error(aqe_last_error()); };
        aqe_sql_query q;
        check(aqe_sql_parse(query.c_str(), sample_percent, &q));
        if (grouped) {
            if (q.group_col == AQE_COL_NONE) throw std::runtime_error("No GROUP BY column found");  // executor.cpp:63
        } else {
            q.group_col = AQE_COL_NONE;  // execute_query never looks at q.group_by (executor.cpp:28-58)
        }
        static thread_local std::vector<aqe_sql_row> buf(AQE_SQL_MAX_GROUPS);  // reused: zero-filling 320 KiB per call costs more than a 10 M-row query
        uint32_t n = 0;
        check(aqe_sql_execute(h_, &q, ci, buf.data(), grouped ? (uint32_t)buf.size() : 1u, &n));
        std::vector<aqe_sql_row> rows(buf.begin(), buf.begin() + std::min<size_t>(n, grouped ? buf.size() : 1));
        for (const aqe_sql_row& r : rows)
            if (r.is_null) throw std::invalid_argument("stod");  // std::stod("NULL"), executor.cpp:46 -> ValueError
        return rows;
    }
    double query(const std::string& q, int p) { return sql(q, p, false, AQE_SQL_VALUE)[0].value; }
    QueryResult query_with_ci(const std::string& q, int p, bool correct) {
        const aqe_sql_row r = sql(q, p, false, correct ? AQE_SQL_CI_CORRECT : AQE_SQL_CI_REFERENCE)[0];
        return QueryResult{r.value, r.ci_lower, r.ci_upper};
    }
    std::map<std::string, double> query_groupby(const std::string& q, int p) {
        std::map<std::string, double> out;  // std::map<std::string, ...> as in executor.h:5: keys sort as text
        for (const aqe_sql_row& r : sql(q, p, true, AQE_SQL_VALUE)) out[std::to_string(r.key)] = r.value;
        return out;
    }
    std::map<std::string, QueryResult> query_groupby_with_ci(const std::string& q, int p, bool correct) {
        std::map<std::string, QueryResult> out;
        for (const aqe_sql_row& r : sql(q, p, true, correct ? AQE_SQL_CI_CORRECT : AQE_SQL_CI_REFERENCE))
            out[std::to_string(r.key)] = QueryResult{r.value, r.ci_lower, r.ci_upper};
        return out;
    }

private:
    void open_handle() {
        if (devices_.size() == 1) check(aqe_create(devices_[0], &h_));
        else check(aqe_create_sharded(devices_.data(), (int)devices_.size(), &h_));
    }
    void reset() {
        aqe_close(h_);
        h_ = nullptr;
        open_handle();
    }
    std::vector<int> devices_;
    aqe_db* h_ = nullptr;
    py::object keep_alive_;
    std::string path_;
    uint64_t seed_ = 0x9E3779B97F4A7C15ull ^ (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
    bool fixed_seed_ = false;
};

// ---- WHERE clause forms accepted by the reference scheduler (custom_scheduler.cpp:277-294) ----------
// `amount BETWEEN a AND b` | `amount >= a AND amount <= b` | `amount > a` (upper bound 99999.99); numbers are
// unsigned decimals; keywords are matched case-sensitively as in the reference.  (-1,-1) = no condition.
struct Scanner {
    const std::string& s; size_t i;
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool ws1() { const size_t b = i; ws(); return i > b; }
    bool lit(const char* w) { const size_t n = std::strlen(w); if (s.compare(i, n, w) == 0) { i += n; return true; } return false; }
    bool num(double& v) {
        const size_t b = i;
        while (i < s.size() && std::isdigit((unsigned char)s[i])) ++i;
        if (i == b) return false;
        if (i + 1 < s.size() && s[i] == '.' && std::isdigit((unsigned char)s[i + 1])) { ++i; while (i < s.size() && std::isdigit((unsigned char)s[i])) ++i; }
        v = std::stod(s.substr(b, i - b));
        return true;
    }
};
std::pair<double, double> where_conditions(const std::string& q) {
    auto each_amount = [&](auto&& f) { for (size_t p = q.find("amount"); p != std::string::npos; p = q.find("amount", p + 1)) { Scanner sc{q, p + 6}; if (f(sc)) return true; } return false; };
    double a = 0, b = 0;
    if (each_amount([&](Scanner& sc) { return sc.ws1() && sc.lit("BETWEEN") && sc.ws1() && sc.num(a) && sc.ws1() && sc.lit("AND") && sc.ws1() && sc.num(b); })) return {a, b};
    if (each_amount([&](Scanner& sc) { sc.ws(); if (!sc.lit(">=")) return false; sc.ws(); if (!sc.num(a)) return false; return sc.ws1() && sc.lit("AND") && sc.ws1() && sc.lit("amount") && (sc.ws(), sc.lit("<=")) && (sc.ws(), sc.num(b)); })) return {a, b};
    if (each_amount([&](Scanner& sc) { sc.ws(); if (!sc.lit(">")) return false; sc.ws(); return sc.num(a); })) return {a, 99999.99};
    return {-1, -1};
}

class CustomApproximateScheduler {  // custom_scheduler.cpp
public:
    explicit CustomApproximateScheduler(double error_threshold) : error_threshold_(error_threshold) {}
    bool create_database(const std::string& p) { return db_.create_database(p); }
    bool open_database(const std::string& p) { return db_.open_database(p); }
    void close_database() { db_.close_database(); }
    bool insert_record(int64_t id, double amount, int32_t region, int32_t product_id, int64_t ts) {
        Record r; r.id = id; r.amount = amount; r.region = region; r.product_id = product_id; r.timestamp = ts;
        return db_.insert_record(r);
    }
    bool insert_batch(const std::vector<Record>& rows) { return db_.insert_batch(rows); }

    // parallel_{sum,avg,count}_sample (custom_bplus_db.cpp:276-343): SRSWOR of floor(N p/100) rows, scaled 100/p.
    // kind: 0 sum, 1 avg, 2 count
    CustomValidationResult sampled(int kind, const std::string& query, double pct, int threads) {
        const auto t0 = std::chrono::high_resolution_clock::now();
        CustomValidationResult r;
        if (kind != 0) { r.confidence_level = 0.0; r.error_margin = 0.0; r.samples_used = 0; }  // reference leaves them unset
        try {
            const uint64_t N = db_.get_total_records();
            aqe_sample_params p = CustomBPlusDB::params(AQE_M_SAMPLE_RECORDS, pct);
            p.num_threads = threads; p.seed = db_.session_seed();
            aqe_plan* pl = nullptr;
            check(aqe_plan_build(db_.handle(), N, AQE_M_SAMPLE_RECORDS, &p, &pl));
            double value = 0.0;
            const uint64_t n = aqe_plan_count(pl);
            if (kind == 2) {
                value = (double)(size_t)((double)n * (100.0 / pct));  // cbd:312-315
            } else if (n) {
                aqe_stats s{};
                int rc;
                const auto w = kind == 0 ? where_conditions(query) : std::pair<double, double>{-1, -1};
                if (w.first != -1 && w.second != -1) rc = aqe_stats_from_plan_where(db_.handle(), pl, AQE_COL_AMOUNT, AQE_COL_AMOUNT, w.first, w.second, &s);
                else rc = aqe_stats_from_plan(db_.handle(), pl, AQE_COL_AMOUNT, &s);
                if (rc) { aqe_plan_free(pl); check(rc); }
                value = s.sum * (100.0 / pct);                          // cbd:303
                if (kind == 1) value = N ? value / (double)N : 0.0;     // cbd:306-310
            }
            aqe_plan_free(pl);
            r.value = value;
            r.status = CustomApproximationStatus::STABLE;
            const double ss = (double)N * pct / 100.0;                  // custom_scheduler.cpp:296-305
            r.confidence_level = ss >= 1000 ? 0.95 : ss >= 500 ? 0.90 : ss >= 100 ? 0.85 : ss >= 50 ? 0.80 : 0.70;
            r.error_margin = pct / 100.0;
            r.samples_used = (int)((double)N * pct / 100.0);
        } catch (const std::exception&) {
            r.value = 0.0; r.status = CustomApproximationStatus::ERROR;
        }
        r.computation_time = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0);
        return r;
    }
    CustomValidationResult exact(int kind) {  // custom_scheduler.cpp:141-205
        const auto t0 = std::chrono::high_resolution_clock::now();
        CustomValidationResult r;
        r.status = CustomApproximationStatus::STABLE; r.confidence_level = 1.0; r.error_margin = 0.0;
        r.samples_used = (int)db_.get_total_records();
        try {
            r.value = kind == 0 ? db_.sum_amount() : kind == 1 ? db_.avg_amount() : (double)db_.count_records();
        } catch (const std::exception&) { r.value = 0.0; r.status = CustomApproximationStatus::ERROR; }
        r.computation_time = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - t0);
        return r;
    }
    BenchmarkResults benchmark_query(const std::string& type, double pct, int threads) {  // custom_scheduler.cpp:207-246
        const int kind = type == "AVG" ? 1 : type == "COUNT" ? 2 : 0;
        const char* q = kind == 1 ? "SELECT AVG(amount)" : kind == 2 ? "SELECT COUNT(*)" : "SELECT SUM(amount)";
        const auto t0 = std::chrono::steady_clock::now();
        const CustomValidationResult e = exact(kind);
        const auto t1 = std::chrono::steady_clock::now();
        const CustomValidationResult a = sampled(kind, q, pct, threads);
        const auto t2 = std::chrono::steady_clock::now();
        BenchmarkResults b;
        b.sample_percentage = pct; b.threads_used = threads;
        b.exact_value = e.value; b.approximate_value = a.value;
        b.exact_time_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        b.approximate_time_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        b.speedup = b.exact_time_ms / b.approximate_time_ms;
        b.error_percentage = e.value != 0 ? std::fabs(e.value - a.value) / std::fabs(e.value) * 100.0 : 0.0;
        return b;
    }
    size_t get_total_records() const { return db_.get_total_records(); }
    size_t get_tree_height() const { return db_.get_tree_height(); }
    double get_database_size_mb() const { return db_.get_total_records() * sizeof(Record) / (1024.0 * 1024.0); }
    CustomBPlusDB& db() { return db_; }

private:
    CustomBPlusDB db_;
    double error_threshold_;
};

// ---- run_query*(sql, db_path, ...): tables opened by path ------------------------------------------------------
// The reference opens the SQLite file anew for every call (core/db.cpp:18-24).  Here db_path names a record file
// (custom_bplus_db.cpp:665-683); loading it into HBM per call would dominate, so the last few tables stay resident,
// keyed by path and invalidated when the file's size or mtime changes.
struct TableCache {
    struct Entry { std::string path; off_t size; int64_t mtime_ns; std::unique_ptr<CustomBPlusDB> db; uint64_t used; };
    std::vector<Entry> entries;
    uint64_t tick = 0;
    static constexpr size_t kMax = 4;
    CustomBPlusDB& get(const std::string& path) {
        struct stat st;
        if (::stat(path.c_str(), &st) != 0) throw std::runtime_error("Cannot open database: unable to open database file");  // core/db.cpp:19-23
        const int64_t mt = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
        for (auto& e : entries)
            if (e.path == path) {
                if (e.size == st.st_size && e.mtime_ns == mt) { e.used = ++tick; return *e.db; }
                e.db.reset();
            }
        entries.erase(std::remove_if(entries.begin(), entries.end(), [](const Entry& e) { return !e.db; }), entries.end());
        char magic[16] = {0};
        if (FILE* f = std::fopen(path.c_str(), "rb")) { const size_t got = std::fread(magic, 1, 16, f); (void)got; std::fclose(f); }
        if (std::memcmp(magic, "SQLite format 3", 15) == 0)
            throw std::runtime_error("run_query: '" + path + "' is a SQLite file; this engine runs the SQL path on the record file that "
                                     "CustomBPlusDB.save_to_file writes (INTEGRATION.md, section 7)");
        auto db = std::make_unique<CustomBPlusDB>();
        if (!db->load_from_file(path)) throw std::runtime_error("Cannot open database: not a record file: " + path);
        if (entries.size() >= kMax) {
            auto lru = std::min_element(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) { return a.used < b.used; });
            entries.erase(lru);
        }
        entries.push_back(Entry{path, st.st_size, mt, std::move(db), ++tick});
        return *entries.back().db;
    }
    void clear() { entries.clear(); }
};
TableCache& table_cache() { static TableCache* c = new TableCache(); return *c; }  // leaked on purpose: no CUDA calls at exit
// parse_query runs before the database is opened (executor.cpp:29-30): a malformed query fails whatever db_path is
CustomBPlusDB& table_for(const std::string& sql, int sample_percent, const std::string& path) {
    aqe_sql_query q;
    if (aqe_sql_parse(sql.c_str(), sample_percent, &q) != AQE_OK) throw std::runtime_error(aqe_last_error());
    return table_cache().get(path);
}

}  // namespace

PYBIND11_MODULE(aqe_backend, m) {
    m.doc() = "ApproximateQueryEngine backend on B200: HBM-resident columnar record table, CUDA scans and samplers";
    m.attr("__engine__") = "aqe_b200";
    m.attr("abi_version") = aqe_abi_version();

    {   // class Record (bindings.cpp:14-20) as a plain CPython type: see RecordObject above
        PyObject* t = PyType_FromSpec(&record_spec);
        if (!t) throw py::error_already_set();
        g_record_type = reinterpret_cast<PyTypeObject*>(t);
        m.add_object("Record", py::reinterpret_steal<py::object>(t));
        Py_INCREF(t);   // g_record_type keeps its own reference for the life of the process
    }

    py::enum_<CustomApproximationStatus>(m, "CustomApproximationStatus")
        .value("STABLE", CustomApproximationStatus::STABLE)
        .value("DRIFTING", CustomApproximationStatus::DRIFTING)
        .value("INSUFFICIENT_DATA", CustomApproximationStatus::INSUFFICIENT_DATA)
        .value("ERROR", CustomApproximationStatus::ERROR);

    py::class_<CustomValidationResult>(m, "CustomValidationResult")
        .def_readonly("value", &CustomValidationResult::value)
        .def_readonly("status", &CustomValidationResult::status)
        .def_readonly("confidence_level", &CustomValidationResult::confidence_level)
        .def_readonly("error_margin", &CustomValidationResult::error_margin)
        .def_readonly("samples_used", &CustomValidationResult::samples_used)
        .def_readonly("computation_time", &CustomValidationResult::computation_time);

    py::class_<QueryResult>(m, "QueryResult")
        .def_readonly("value", &QueryResult::value)
        .def_readonly("ci_lower", &QueryResult::ci_lower)
        .def_readonly("ci_upper", &QueryResult::ci_upper);

    py::class_<ApproxResult>(m, "ApproxResult")
        .def_readonly("value", &ApproxResult::value)
        .def_readonly("estimate", &ApproxResult::value)
        .def_readonly("ci_lower", &ApproxResult::ci_lower)
        .def_readonly("ci_upper", &ApproxResult::ci_upper)
        .def_readonly("status", &ApproxResult::status)
        .def_readonly("confidence_level", &ApproxResult::confidence_level)
        .def_readonly("error_margin", &ApproxResult::error_margin)
        .def_readonly("samples_used", &ApproxResult::samples_used)
        .def_readonly("population", &ApproxResult::population)
        .def_readonly("rounds", &ApproxResult::rounds)
        .def_readonly("kernel_us", &ApproxResult::kernel_us)
        .def_readonly("computation_time", &ApproxResult::computation_time);

    py::class_<BenchmarkResults>(m, "BenchmarkResults")
        .def_readonly("exact_value", &BenchmarkResults::exact_value)
        .def_readonly("approximate_value", &BenchmarkResults::approximate_value)
        .def_readonly("exact_time_ms", &BenchmarkResults::exact_time_ms)
        .def_readonly("approximate_time_ms", &BenchmarkResults::approximate_time_ms)
        .def_readonly("speedup", &BenchmarkResults::speedup)
        .def_readonly("error_percentage", &BenchmarkResults::error_percentage)
        .def_readonly("threads_used", &BenchmarkResults::threads_used)
        .def_readonly("sample_percentage", &BenchmarkResults::sample_percentage);

    using DB = CustomBPlusDB;
    auto P = [](int method, double pct) { return DB::params(method, pct); };
    py::class_<DB> db(m, "CustomBPlusDB");
    db.def(py::init<>())
        .def(py::init<int>(), py::arg("device"))
        .def(py::init<std::vector<int>>(), py::arg("devices"), "the table range-sharded over these GPUs of this process")
        .def_property_readonly("shard_count", &DB::shard_count, "GPUs holding rows of the current table")
        .def_property_readonly("shards_fused", &DB::shards_fused, "shard partials are exchanged inside the kernels (peer-mapped mailboxes)")
        .def_property_readonly("devices", &DB::devices)
        .def("create_database", &DB::create_database)
        .def("open_database", &DB::open_database)
        .def("close_database", &DB::close_database)
        .def("insert_record", &DB::insert_record)
        .def("sum_amount", &DB::sum_amount)
        .def("sum_amount_where", &DB::sum_amount_where)
        .def("get_total_records", &DB::get_total_records)
        .def("get_node_count", &DB::get_node_count)
        .def("save_to_file", &DB::save_to_file)
        .def("load_from_file", &DB::load_from_file)
        // --- list[Record] samplers, same names / defaults as bindings.cpp:49-101 ---
        .def("sample_records", [P](DB& d, double p) { auto a = P(AQE_M_SAMPLE_RECORDS, p); a.seed = d.session_seed(); return d.sample(AQE_M_SAMPLE_RECORDS, a); })
        .def("optimized_sequential_sample", [P](DB& d, double p) { auto a = P(AQE_M_OPTIMIZED_SEQUENTIAL, p); a.seed = d.session_seed(); return d.sample(AQE_M_OPTIMIZED_SEQUENTIAL, a); })
        .def("fast_pointer_sample", [P](DB& d, double p, int step) { auto a = P(AQE_M_FAST_POINTER, p); a.step_size = step; return d.sample(AQE_M_FAST_POINTER, a); },
             py::arg("sample_percent"), py::arg("step_size") = 2)
        .def("slow_pointer_sample", [P](DB& d, double p) { return d.sample(AQE_M_SLOW_POINTER, P(AQE_M_SLOW_POINTER, p)); })
        .def("dual_pointer_sample", [P](DB& d, double p) { return d.sample(AQE_M_DUAL_POINTER, P(AQE_M_DUAL_POINTER, p)); })
        .def("parallel_pointer_sample", [P](DB& d, double p, int th) { auto a = P(AQE_M_PARALLEL_POINTER, p); a.num_threads = th; return d.sample(AQE_M_PARALLEL_POINTER, a); },
             py::arg("sample_percent"), py::arg("num_threads") = 4)
        .def("random_pointer_sample", [P](DB& d, double p, unsigned int seed) { auto a = P(AQE_M_RANDOM_POINTER, p); a.seed = seed; return d.sample(AQE_M_RANDOM_POINTER, a); },
             py::arg("sample_percent"), py::arg("seed") = 42)
        .def("clt_validated_dual_pointer_sample",
             [P](DB& d, double p, double conf, int ci, int th, double maxerr) {
                 auto a = P(AQE_M_CLT_VALIDATED_DUAL_POINTER, p);
                 a.confidence_level = conf; a.check_interval = ci; a.num_threads = th; a.max_error_percent = maxerr;
                 return d.sample(AQE_M_CLT_VALIDATED_DUAL_POINTER, a);
             },
             py::arg("sample_percent"), py::arg("confidence_level") = 0.95, py::arg("check_interval") = 10, py::arg("num_threads") = 4,
             py::arg("max_error_percent") = 2.0)
        .def("optimized_clt_sample",
             [P](DB& d, double p, double conf, int ci, int th, double maxerr) {
                 auto a = P(AQE_M_OPTIMIZED_CLT, p);
                 a.confidence_level = conf; a.check_interval = ci; a.num_threads = th; a.max_error_percent = maxerr;
                 return d.sample(AQE_M_OPTIMIZED_CLT, a);
             },
             py::arg("sample_percent"), py::arg("confidence_level") = 0.95, py::arg("check_interval") = 20, py::arg("num_threads") = 4,
             py::arg("max_error_percent") = 2.0)
        .def("block_sample", [P](DB& d, double p, size_t bs) { auto a = P(AQE_M_BLOCK, p); a.block_size = (int64_t)bs; return d.sample(AQE_M_BLOCK, a); },
             py::arg("sample_percent"), py::arg("block_size") = 1000)
        .def("page_sample", [P](DB& d, double p, size_t ps) { auto a = P(AQE_M_PAGE, p); a.block_size = (int64_t)ps; return d.sample(AQE_M_PAGE, a); },
             py::arg("sample_percent"), py::arg("page_size") = 4096)
        .def("parallel_block_sample",
             [P](DB& d, double p, size_t bs, int th) { auto a = P(AQE_M_PARALLEL_BLOCK, p); a.block_size = (int64_t)bs; a.num_threads = th; return d.sample(AQE_M_PARALLEL_BLOCK, a); },
             py::arg("sample_percent"), py::arg("block_size") = 1000, py::arg("num_threads") = 4)
        .def("adaptive_block_sample",
             [P](DB& d, double p, size_t mn, size_t mx) { auto a = P(AQE_M_ADAPTIVE_BLOCK, p); a.block_size = (int64_t)mn; a.block_size_max = (int64_t)mx; return d.sample(AQE_M_ADAPTIVE_BLOCK, a); },
             py::arg("sample_percent"), py::arg("min_block_size") = 500, py::arg("max_block_size") = 2000)
        .def("stratified_block_sample",
             [P](DB& d, double p, size_t bs, int k) { auto a = P(AQE_M_STRATIFIED_BLOCK, p); a.block_size = (int64_t)bs; a.block_size_max = k; return d.sample(AQE_M_STRATIFIED_BLOCK, a); },
             py::arg("sample_percent"), py::arg("block_size") = 1000, py::arg("strata_count") = 4)
        .def("index_based_sample", [P](DB& d, double p) { return d.sample(AQE_M_INDEX_BASED, P(AQE_M_INDEX_BASED, p)); })
        .def("node_skip_sample", [P](DB& d, double p, int skip) { auto a = P(AQE_M_NODE_SKIP, p); a.step_size = skip; return d.sample(AQE_M_NODE_SKIP, a); },
             py::arg("sample_percent"), py::arg("skip_factor") = 2)
        .def("balanced_tree_sample", [P](DB& d, double p) { return d.sample(AQE_M_BALANCED_TREE, P(AQE_M_BALANCED_TREE, p)); })
        .def("direct_access_sample", [P](DB& d, double p) { return d.sample(AQE_M_DIRECT_ACCESS, P(AQE_M_DIRECT_ACCESS, p)); })
        .def("byte_offset_sample", [P](DB& d, double p) { return d.sample(AQE_M_BYTE_OFFSET, P(AQE_M_BYTE_OFFSET, p)); })
        .def("random_start_nth_sample", [P](DB& d, double p, int nth) { auto a = P(AQE_M_RANDOM_START_NTH, p); a.step_size = nth; a.seed = d.session_seed(); return d.sample(AQE_M_RANDOM_START_NTH, a); },
             py::arg("sample_percent"), py::arg("nth") = 10)
        .def("memory_stride_sample", [P](DB& d, double p, size_t sb) { auto a = P(AQE_M_MEMORY_STRIDE, p); a.block_size = (int64_t)sb; return d.sample(AQE_M_MEMORY_STRIDE, a); },
             py::arg("sample_percent"), py::arg("stride_bytes") = 0)
        .def("address_arithmetic_sample", [P](DB& d, double p) { auto a = P(AQE_M_ADDRESS_ARITHMETIC, p); a.seed = d.session_seed(); return d.sample(AQE_M_ADDRESS_ARITHMETIC, a); })
        .def("optimized_address_arithmetic_sample", [P](DB& d, double p) { return d.sample(AQE_M_OPT_ADDRESS_ARITHMETIC, P(AQE_M_OPT_ADDRESS_ARITHMETIC, p)); })
        .def("random_start_memory_stride_sample",
             [P](DB& d, double p, size_t sb) { auto a = P(AQE_M_RANDOM_START_MEMORY_STRIDE, p); a.block_size = (int64_t)sb; a.seed = d.session_seed(); return d.sample(AQE_M_RANDOM_START_MEMORY_STRIDE, a); },
             py::arg("sample_percent"), py::arg("stride_bytes") = 0)
        .def("multithreaded_memory_stride_sample",
             [P](DB& d, double p, int th) { auto a = P(AQE_M_MULTITHREADED_MEMORY_STRIDE, p); a.num_threads = th; a.seed = d.session_seed(); return d.sample(AQE_M_MULTITHREADED_MEMORY_STRIDE, a); },
             py::arg("sample_percent"), py::arg("num_threads") = 4)
        .def("fast_aggregated_memory_stride_sum", &DB::fast_aggregated_memory_stride_sum, py::arg("sample_percent"), py::arg("num_threads") = 4)
        .def("signal_based_clt_sample", [P](DB& d, double p, int ci) { auto a = P(AQE_M_SIGNAL_BASED_CLT, p); a.check_interval = ci; return d.sample(AQE_M_SIGNAL_BASED_CLT, a); },
             py::arg("sample_percent"), py::arg("check_interval") = 10)
        // --- additive surface ---
        .def("avg_amount", &DB::avg_amount)
        .def("count_records", &DB::count_records)
        .def("get_tree_height", &DB::get_tree_height)
        .def("insert_batch", &DB::insert_batch)
        .def("insert_array", &DB::insert_array, "append a numpy structured array of 32-byte records")
        .def("from_array", &DB::from_array, "replace the table with a numpy structured array of 32-byte records")
        .def("load_shard", &DB::load_shard, py::arg("path"), py::arg("first_row"), py::arg("n_rows"))
        .def("generate_synthetic", &DB::generate_synthetic, py::arg("n_rows"), py::arg("seed") = 7, py::arg("first_row") = 0, py::arg("dist") = 0,
             py::arg("columns_mask") = 0x1f)
        .def("attach_columns", &DB::attach_columns, py::arg("id"), py::arg("amount"), py::arg("region"), py::arg("product_id"), py::arg("timestamp"), py::arg("n"))
        .def("from_torch", &DB::from_torch, py::arg("id") = py::none(), py::arg("amount") = py::none(), py::arg("region") = py::none(),
             py::arg("product_id") = py::none(), py::arg("timestamp") = py::none(),
             "borrow CUDA tensors as the table's columns (zero copy); columns that are never queried may be omitted")
        .def("column_ptr", &DB::column_ptr)
        .def("sum_column", &DB::sum_column)
        .def("scan", &DB::scan, py::arg("agg_col") = "amount", py::arg("pred_col") = py::none(), py::arg("lo") = 0.0, py::arg("hi") = 0.0)
        .def("query", &DB::query, py::arg("sql_query"), py::arg("sample_percent") = 0, "run_query on this table")
        .def("query_groupby", &DB::query_groupby, py::arg("sql_query"), py::arg("sample_percent") = 0, "run_query_groupby on this table")
        .def("query_with_ci", &DB::query_with_ci, py::arg("sql_query"), py::arg("sample_percent") = 0, py::arg("correct_ci") = false,
             "run_query_with_ci on this table; correct_ci=True scales the SUM interval as a total")
        .def("query_groupby_with_ci", &DB::query_groupby_with_ci, py::arg("sql_query"), py::arg("sample_percent") = 0, py::arg("correct_ci") = false)
        .def("set_seed", &DB::set_seed, py::arg("seed") = py::none(), "fix the seed that replaces std::random_device (None = fresh per call)")
        .def("sample_array",
             [](DB& d, const std::string& method, double pct, py::kwargs kw) {
                 static const std::map<std::string, int> ids = {
                     {"slow_pointer", 0}, {"fast_pointer", 1}, {"dual_pointer", 2}, {"parallel_pointer", 3}, {"random_pointer", 4}, {"memory_stride", 5},
                     {"optimized_address_arithmetic", 6}, {"index_based", 7}, {"byte_offset", 8}, {"optimized_clt", 9}, {"block", 10}, {"page", 11},
                     {"parallel_block", 12}, {"node_skip", 13}, {"balanced_tree", 14}, {"direct_access", 15}, {"adaptive_block", 16},
                     {"stratified_block", 17}, {"sample_records", 18}, {"optimized_sequential", 19}, {"random_start_nth", 20}, {"address_arithmetic", 21},
                     {"random_start_memory_stride", 22}, {"multithreaded_memory_stride", 23}, {"clt_validated_dual_pointer", 24}, {"signal_based_clt", 25}};
                 auto it = ids.find(method);
                 if (it == ids.end()) throw py::value_error("unknown sampler '" + method + "'");
                 aqe_sample_params a = DB::params(it->second, pct);
                 bool seeded = false, stats = false;
                 std::string col = "amount";
                 for (auto kv : kw) {
                     const std::string k = kv.first.cast<std::string>();
                     if (k == "step_size" || k == "skip_factor" || k == "nth") a.step_size = kv.second.cast<int64_t>();
                     else if (k == "num_threads") a.num_threads = kv.second.cast<int64_t>();
                     else if (k == "block_size" || k == "page_size" || k == "stride_bytes" || k == "min_block_size") a.block_size = kv.second.cast<int64_t>();
                     else if (k == "block_size_max" || k == "max_block_size" || k == "strata_count") a.block_size_max = kv.second.cast<int64_t>();
                     else if (k == "check_interval") a.check_interval = kv.second.cast<int64_t>();
                     else if (k == "confidence_level") a.confidence_level = kv.second.cast<double>();
                     else if (k == "max_error_percent") a.max_error_percent = kv.second.cast<double>();
                     else if (k == "seed") { a.seed = kv.second.cast<uint64_t>(); seeded = true; }
                     else if (k == "stats") stats = kv.second.cast<bool>();
                     else if (k == "column") col = kv.second.cast<std::string>();
                     else throw py::value_error("unknown sampler argument '" + k + "'");
                 }
                 if (!seeded && it->second >= AQE_M_SAMPLE_RECORDS && it->second <= AQE_M_MULTITHREADED_MEMORY_STRIDE) a.seed = d.session_seed();
                 if (stats) return py::object(d.sample_stats(it->second, a, col));
                 return py::object(d.sample_array(it->second, a));
             },
             py::arg("method"), py::arg("sample_percent"),
             "numpy structured array of the sampled rows (or, with stats=True, {n, mean, m2, sum} computed on the device without returning rows)")
        .def("approx_sum", [](DB& d, double e, double c, const std::string& design, py::object seed, py::object where, const std::string& wc, uint32_t bs, uint64_t mn, uint64_t mx) { return d.approx(AQE_AGG_SUM, e, c, design, seed, where, wc, bs, mn, mx); },
             py::arg("error_percent") = 1.0, py::arg("confidence_level") = 0.95, py::arg("design") = "srs", py::arg("seed") = py::none(),
             py::arg("where") = py::none(), py::arg("where_column") = "amount", py::arg("block_size") = 1000, py::arg("min_samples") = 0, py::arg("max_samples") = 0)
        .def("approx_avg", [](DB& d, double e, double c, const std::string& design, py::object seed, py::object where, const std::string& wc, uint32_t bs, uint64_t mn, uint64_t mx) { return d.approx(AQE_AGG_AVG, e, c, design, seed, where, wc, bs, mn, mx); },
             py::arg("error_percent") = 1.0, py::arg("confidence_level") = 0.95, py::arg("design") = "srs", py::arg("seed") = py::none(),
             py::arg("where") = py::none(), py::arg("where_column") = "amount", py::arg("block_size") = 1000, py::arg("min_samples") = 0, py::arg("max_samples") = 0)
        .def("approx_count", [](DB& d, double e, double c, const std::string& design, py::object seed, py::object where, const std::string& wc, uint32_t bs, uint64_t mn, uint64_t mx) { return d.approx(AQE_AGG_COUNT, e, c, design, seed, where, wc, bs, mn, mx); },
             py::arg("error_percent") = 1.0, py::arg("confidence_level") = 0.95, py::arg("design") = "srs", py::arg("seed") = py::none(),
             py::arg("where") = py::none(), py::arg("where_column") = "amount", py::arg("block_size") = 1000, py::arg("min_samples") = 0, py::arg("max_samples") = 0);
    db.attr("record_dtype") = DB::record_dtype();

    using S = CustomApproximateScheduler;
    py::class_<S>(m, "CustomApproximateScheduler")
        .def(py::init<double>(), py::arg("error_threshold") = 0.05)
        .def("create_database", &S::create_database)
        .def("open_database", &S::open_database)
        .def("close_database", &S::close_database)
        .def("insert_record", &S::insert_record)
        .def("insert_batch", &S::insert_batch)
        .def("execute_sum_query", [](S& s, const std::string& q, double p, int t) { return s.sampled(0, q, p, t); }, py::arg("query"), py::arg("sample_percent") = 10.0, py::arg("num_threads") = 4)
        .def("execute_avg_query", [](S& s, const std::string& q, double p, int t) { return s.sampled(1, q, p, t); }, py::arg("query"), py::arg("sample_percent") = 10.0, py::arg("num_threads") = 4)
        .def("execute_count_query", [](S& s, const std::string& q, double p, int t) { return s.sampled(2, q, p, t); }, py::arg("query"), py::arg("sample_percent") = 10.0, py::arg("num_threads") = 4)
        .def("execute_exact_sum", [](S& s) { return s.exact(0); })
        .def("execute_exact_avg", [](S& s) { return s.exact(1); })
        .def("execute_exact_count", [](S& s) { return s.exact(2); })
        .def("benchmark_query", &S::benchmark_query, py::arg("query_type"), py::arg("sample_percent") = 10.0, py::arg("num_threads") = 4)
        .def("get_total_records", &S::get_total_records)
        .def("get_tree_height", &S::get_tree_height)
        .def("get_database_size_mb", &S::get_database_size_mb)
        .def_property_readonly("db", &S::db, py::return_value_policy::reference_internal)
        .def_static("_where_conditions", &where_conditions);

    // bindings.cpp:126-136.  db_path names a record file (save_to_file format), not a SQLite file; num_threads is accepted
    // and ignored (one grouped-scan kernel serves every group).
    m.def("run_query", [](const std::string& q, const std::string& path, int p) { return table_for(q, p, path).query(q, p); },
          "Execute SQL query with sampling and automatic scaling", py::arg("sql_query"), py::arg("db_path"), py::arg("sample_percent") = 0);
    m.def("run_query_groupby", [](const std::string& q, const std::string& path, int p, int) { return table_for(q, p, path).query_groupby(q, p); },
          "Execute GROUP BY query with sampling", py::arg("sql_query"), py::arg("db_path"), py::arg("sample_percent") = 0, py::arg("num_threads") = 4);
    m.def("run_query_with_ci", [](const std::string& q, const std::string& path, int p) { return table_for(q, p, path).query_with_ci(q, p, false); },
          "Execute query with confidence intervals", py::arg("sql_query"), py::arg("db_path"), py::arg("sample_percent") = 0);
    m.def("run_query_groupby_with_ci", [](const std::string& q, const std::string& path, int p, int) { return table_for(q, p, path).query_groupby_with_ci(q, p, false); },
          "Execute GROUP BY query with confidence intervals", py::arg("sql_query"), py::arg("db_path"), py::arg("sample_percent") = 0, py::arg("num_threads") = 4);
    m.def("close_cached_tables", [] { table_cache().clear(); }, "drop the tables run_query* keeps resident");
    m.def("device_count", [] { int n = 0; aqe_device_count(&n); return n; });
    m.def("launch_count", [] { return aqe_launch_count(); });
    m.def("z_score", &aqe_z_score, py::arg("confidence_level"), py::arg("exact") = 1);
}
