/*
 * oracle/aqe_oracle.c -- TEST INFRASTRUCTURE (the checker), never the product.
 *
 * A plain-C, single-threaded CPU restatement of the reference's hot path: exact and sampled
 * SUM/AVG/COUNT over the fixed-width record table of CustomBPlusDB.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's library.  The product
 * (libaqe_b200.so) never links, loads or calls it and has no CPU fallback.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path.
 * This restatement is therefore pinned against the reference ITSELF: oracle/_ref/libaqe_ref.so is the
 * unmodified reference core compiled from /root/reference (oracle/Makefile) and tests/test_oracle_vs_ref.py
 * checks every function below against it; tests/golden/ holds vectors minted from it
 * (tests/golden/make_golden.py) so the pin travels to machines without /root/reference.
 *
 * Citations are <file>:<line> relative to /root/reference/; "cbd" = src/aqe_backend/core/custom_bplus_db.cpp,
 * "cli" = enhanced_aqe_cli.py.
 *
 * Compile with -ffp-contract=off: every floating-point expression is evaluated as written.
 */
#include "aqe_b200.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ================================================================================================
 * 1. Philox4x32-10 and the synthetic "sales-shaped" table.  The reference has no generator
 *    for its record format (tools/create_db.py creates nothing); this one is ours and is
 *    restated here so host files and device-generated shards hold identical bits.
 * ============================================================================================== */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ORC_API void orc_philox(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
    philox4x32_10((uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32),
                  (uint32_t)key, (uint32_t)(key >> 32), out);
}

#define SYNTH_STREAM 0x41514544u /* "AQED": data stream id in counter word 2 */

static inline double u53(uint32_t hi, uint32_t lo) {
    uint64_t u = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)u * (1.0 / 9007199254740992.0);
}

ORC_API void orc_synth_row(uint64_t seed, uint64_t row, int dist, aqe_record* r) {
    uint32_t o[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), SYNTH_STREAM, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    double x = u53(o[0], o[1]);
    r->id = (int64_t)row + 1;
    if (dist == AQE_SYNTH_LOGNORMAL) {
        /* Box-Muller on a second Philox block; libm here vs CUDA libdevice on the device, so this
         * distribution is NOT bit-identical across host/device (tests read the device column back). */
        uint32_t q[4];
        philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), SYNTH_STREAM, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), q);
        double u1 = u53(q[0], q[1]);
        double u2 = u53(q[2], q[3]);
        double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
        r->amount = exp(4.0 + 1.5 * z);
    } else {
        double t = 999.0 * x; /* two roundings, no FMA: the device uses __dmul_rn/__dadd_rn */
        r->amount = 1.0 + t;
    }
    r->region = (int32_t)(o[2] & 7u);
    r->product_id = (int32_t)(o[3] % 1000u);
    r->timestamp = 1700000000LL + (int64_t)row;
}

ORC_API void orc_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, int dist, aqe_record* rows) {
    for (uint64_t i = 0; i < n; ++i) orc_synth_row(seed, first_row + i, dist, &rows[i]);
}

/* ================================================================================================
 * 2. File format -- cbd:665-683 (save_to_file) / cbd:685-711 (load_from_file)
 *    u64 total_records | u64 tree_height | u64 record_count | record_count x 32-byte Record
 * ============================================================================================== */
static uint64_t leaves_for(uint64_t n);
ORC_API uint64_t orc_tree_height(uint64_t n);

ORC_API int orc_save_file(const char* path, const aqe_record* rows, uint64_t n) {
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    uint64_t hdr[3] = {n, orc_tree_height(n), n};
    int ok = fwrite(hdr, 8, 3, f) == 3;
    if (ok && n) ok = fwrite(rows, sizeof(aqe_record), n, f) == n;
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : 1;
}

ORC_API int64_t orc_file_count(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    uint64_t hdr[3];
    int ok = fread(hdr, 8, 3, f) == 3;
    fclose(f);
    return ok ? (int64_t)hdr[2] : -1;
}

static int cmp_id(const void* a, const void* b) {
    int64_t x = ((const aqe_record*)a)->id, y = ((const aqe_record*)b)->id;
    return (x > y) - (x < y);
}

/* Reads rows and orders them by id as load_from_file's insert_batch does (cbd:198-200).  Ties keep file
 * order here (merge sort would; qsort may not -- files with duplicate ids are outside the golden set). */
ORC_API int orc_load_file(const char* path, aqe_record* rows, uint64_t cap, uint64_t* n_out) {
    FILE* f = fopen(path, "rb");
    if (!f) return 1;
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) { fclose(f); return 1; }
    uint64_t n = hdr[2];
    *n_out = n;
    if (n > cap) { fclose(f); return 2; }
    if (n && fread(rows, sizeof(aqe_record), n, f) != n) { fclose(f); return 1; }
    fclose(f);
    int sorted = 1;
    for (uint64_t i = 1; i < n; ++i) if (rows[i].id < rows[i - 1].id) { sorted = 0; break; }
    if (!sorted) qsort(rows, n, sizeof(aqe_record), cmp_id);
    return 0;
}

/* ================================================================================================
 * 3. Exact scans -- cbd:242-251 sum_amount (strict left-to-right in id order), cbd:263-274
 *    sum_amount_where (closed interval on amount itself), cbd:259-261/646-648 count.
 * ============================================================================================== */
ORC_API double orc_sum_amount(const aqe_record* rows, uint64_t n) {
    double sum = 0.0;
    for (uint64_t i = 0; i < n; ++i) sum += rows[i].amount; /* cbd:247-249 */
    return sum;
}

ORC_API double orc_sum_amount_where(const aqe_record* rows, uint64_t n, double lo, double hi, uint64_t* count) {
    double sum = 0.0;
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (rows[i].amount >= lo && rows[i].amount <= hi) { sum += rows[i].amount; ++c; } /* cbd:269-271 */
    }
    if (count) *count = c;
    return sum;
}

ORC_API double orc_avg_amount(const aqe_record* rows, uint64_t n) { /* cbd:253-257 */
    return n ? orc_sum_amount(rows, n) / (double)n : 0.0;
}

static inline double col_as_double(const aqe_record* r, int col) {
    switch (col) {
        case AQE_COL_ID: return (double)r->id;
        case AQE_COL_AMOUNT: return r->amount;
        case AQE_COL_REGION: return (double)r->region;
        case AQE_COL_PRODUCT_ID: return (double)r->product_id;
        case AQE_COL_TIMESTAMP: return (double)r->timestamp;
        default: return 0.0;
    }
}
static inline int64_t col_as_i64(const aqe_record* r, int col) {
    switch (col) {
        case AQE_COL_ID: return r->id;
        case AQE_COL_REGION: return r->region;
        case AQE_COL_PRODUCT_ID: return r->product_id;
        case AQE_COL_TIMESTAMP: return r->timestamp;
        default: return 0;
    }
}

/* Generic scan: same loop shape as cbd:263-274 with the predicate on any column (evaluated in double,
 * closed interval) and the aggregate on any column.  Integer aggregates (new capability,
 * oracle = int128 sum of the reference-loaded rows) are exact. */
ORC_API void orc_scan(const aqe_record* rows, uint64_t n, int agg_col, int pred_col, double lo, double hi,
                      aqe_partial* out) {
    memset(out, 0, sizeof(*out));
    __int128 isum = 0;
    double sum = 0.0, sumsq = 0.0, mn = INFINITY, mx = -INFINITY;
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (pred_col != AQE_COL_NONE) {
            double pv = col_as_double(&rows[i], pred_col);
            if (!(pv >= lo && pv <= hi)) continue;
        }
        ++c;
        if (agg_col == AQE_COL_AMOUNT) {
            double v = rows[i].amount;
            sum += v;
            sumsq += v * v;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        } else {
            isum += (__int128)col_as_i64(&rows[i], agg_col);
        }
    }
    out->count = c;
    out->sum = sum;
    out->sumsq = sumsq;
    out->minv = mn;
    out->maxv = mx;
    out->isum_lo = (uint64_t)isum;
    out->isum_hi = (int64_t)(isum >> 64);
    if (agg_col != AQE_COL_AMOUNT) out->sum = (double)isum;
}

/* Exactly rounded sum (Shewchuk-style expansion kept small: Neumaier in long double is enough for the
 * 1e-12 gate; tests also use Python's math.fsum). */
ORC_API double orc_sum_amount_ld(const aqe_record* rows, uint64_t n) {
    long double s = 0.0L, c = 0.0L;
    for (uint64_t i = 0; i < n; ++i) {
        long double x = rows[i].amount, t = s + x;
        if (fabsl(s) >= fabsl(x)) c += (s - t) + x; else c += (x - t) + s;
        s = t;
    }
    return (double)(s + c);
}

/* ================================================================================================
 * 4. Bulk-load tree shape.  load_from_file inserts rows in ascending id (cbd:196-208); a leaf splits
 *    when key_count reaches MAX_KEYS=255 (cbd:215) into 127 | 128 (cbd:45-56), an internal node when
 *    key_count reaches 255 (256 children, cbd:235) into 128 | 128 children (cbd:63-70).  All inserts
 *    go to the rightmost path, so the shape is a closed form of N.
 * ============================================================================================== */
static uint64_t leaves_for(uint64_t n) { return n < 255 ? 1 : (n - 255) / 127 + 2; }
static uint64_t parents_for(uint64_t children) { return children <= 255 ? 1 : (children - 256) / 128 + 2; }

ORC_API uint64_t orc_leaf_count(uint64_t n) { return leaves_for(n); }
ORC_API uint64_t orc_tree_height(uint64_t n) { /* cbd:650; height 1 = a single leaf */
    uint64_t h = 1, c = leaves_for(n);
    while (c > 1) { c = parents_for(c); ++h; }
    return h;
}
ORC_API uint64_t orc_node_count(uint64_t n) { return n / 255 + 1; } /* cbd:654-658 (a formula, not a count) */

static inline uint64_t leaf_size(uint64_t n, uint64_t L, uint64_t j) { /* rows in leaf j (0-based) */
    return j + 1 < L ? 127 : n - 127 * (L - 1);
}

/* ================================================================================================
 * 5. Sampler index sets.  Each returns positions into R (rows in ascending id,
 *    collect_leaf_records cbd:715-735).  T = static_cast<int>(N*p/100.0) (e.g. cbd:745) is kept in 64 bits.
 * ============================================================================================== */
typedef struct { int64_t* v; uint64_t n, cap; } ivec;
static void push(ivec* o, int64_t x) { if (o->n < o->cap) o->v[o->n] = x; o->n++; }

static int64_t target_count(uint64_t n, double p) { return (int64_t)((double)n * p / 100.0); }
static int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

static void gen_slow_pointer(uint64_t N, double p, int64_t mult, ivec* o) { /* cbd:759-778; fast: cbd:737-757 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return;
    int64_t step = imax(1, (int64_t)(N / (uint64_t)T)) * mult;
    for (uint64_t i = 0; i < N && (int64_t)o->n < T; i += (uint64_t)step) push(o, (int64_t)i);
}

static int gen_dual_pointer(uint64_t N, double p, ivec* o) { /* cbd:780-812 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return 0;
    int64_t Tf = T / 3, Ts = T - Tf;
    if (Tf == 0) return AQE_ERR_INVALID; /* reference divides by zero (cbd:796) */
    int64_t fs = imax(1, (int64_t)(N / (uint64_t)Tf)) * 3;
    for (uint64_t i = 0; i < N && (int64_t)o->n < Tf; i += (uint64_t)fs) push(o, (int64_t)i);
    int64_t ss = imax(1, (int64_t)(N / (uint64_t)Ts));
    for (uint64_t i = (uint64_t)(fs / 2); i < N && (int64_t)o->n < T; i += (uint64_t)ss) push(o, (int64_t)i);
    return 0;
}

static int gen_parallel_pointer(uint64_t N, double p, int64_t Th, ivec* o) { /* cbd:814-854 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return 0;
    if (Th <= 0) return AQE_ERR_INVALID;
    int64_t spt = T / Th;
    int64_t step = imax(1, (int64_t)(N / (uint64_t)T));
    for (int64_t t = 0; t < Th; ++t) {
        uint64_t start = (N / (uint64_t)Th) * (uint64_t)t;
        int64_t c = 0;
        for (uint64_t i = start; i < N && c < spt; i += (uint64_t)step) { push(o, (int64_t)i); ++c; }
    }
    return 0;
}

/* --- std::mt19937 + libstdc++ uniform_int_distribution<size_t> (GCC 13 <bits/uniform_int_dist.h>):
 * a 32-bit generator and a range < 2^32 take the Lemire multiply-shift path `_S_nd<uint64_t>`.  This is
 * implementation-defined; it is what the reference build uses on this toolchain. */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(mt19937_t* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* uniform_int_distribution<uint64_t>(0, range_max) driven by a 32-bit URBG. */
static uint64_t mt_uniform(mt19937_t* g, uint64_t range_max) {
    const uint64_t urngrange = 0xffffffffull;
    if (urngrange > range_max) {
        uint32_t erange = (uint32_t)(range_max + 1);
        uint64_t product = (uint64_t)mt_next(g) * (uint64_t)erange;
        uint32_t low = (uint32_t)product;
        if (low < erange) {
            uint32_t threshold = (uint32_t)(-erange) % erange;
            while (low < threshold) {
                product = (uint64_t)mt_next(g) * (uint64_t)erange;
                low = (uint32_t)product;
            }
        }
        return product >> 32;
    } else if (urngrange < range_max) {
        /* upscaling branch: tmp = uerngrange * dist(0, urange/uerngrange) ; ret = tmp + g() */
        uint64_t ret, tmp;
        const uint64_t uerngrange = urngrange + 1;
        do {
            tmp = uerngrange * mt_uniform(g, range_max / uerngrange);
            ret = tmp + (uint64_t)mt_next(g);
        } while (ret > range_max || ret < tmp);
        return ret;
    }
    return (uint64_t)mt_next(g);
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

static int gen_random_pointer(uint64_t N, double p, uint32_t seed, ivec* o) { /* cbd:856-882 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return 0;
    uint64_t want = (uint64_t)T < N ? (uint64_t)T : N;
    /* std::set<size_t> until `want` distinct: bitmap for membership, sorted at the end */
    uint8_t* seen = (uint8_t*)calloc((N + 7) / 8, 1);
    int64_t* tmp = (int64_t*)malloc(want * sizeof(int64_t));
    if (!seen || !tmp) { free(seen); free(tmp); return AQE_ERR_NOMEM; }
    mt19937_t g;
    mt_seed(&g, seed);
    uint64_t got = 0;
    while (got < want) {
        uint64_t x = mt_uniform(&g, N - 1);
        if (!(seen[x >> 3] & (1u << (x & 7)))) { seen[x >> 3] |= (uint8_t)(1u << (x & 7)); tmp[got++] = (int64_t)x; }
    }
    qsort(tmp, want, sizeof(int64_t), cmp_i64);
    for (uint64_t i = 0; i < want; ++i) push(o, tmp[i]);
    free(seen); free(tmp);
    return 0;
}

/* |cached_records_| after a load of N rows: insert_record refreshes the "mmap" cache only when
 * total_records % 1000 == 0 (cbd:188-191), so methods that read the cache see M = 1000*floor(N/1000) rows
 * once N >= 1000.  Below 1000 rows the cache is never built and the reference's behaviour depends on
 * which sampler ran before (root->subtree_record_count is stale); the engine uses M = N there. */
static uint64_t cache_rows(uint64_t N) { return N >= 1000 ? N - N % 1000 : N; }
ORC_API uint64_t orc_cache_rows(uint64_t N) { return cache_rows(N); }

/* cbd:1526-1566 (cached path), M = |cached_records_|. */
static void gen_memory_stride(uint64_t M, double p, int64_t stride_bytes, uint64_t start, ivec* o) {
    int64_t T = target_count(M, p);
    if (M == 0 || T <= 0) return;
    uint64_t stride = stride_bytes == 0 ? (uint64_t)imax(1, (int64_t)(M / (uint64_t)T))
                                        : (uint64_t)imax(1, stride_bytes / 32);
    for (uint64_t off = start; (int64_t)o->n < T && off < M; off += stride) push(o, (int64_t)off);
}

static void gen_opt_address_arithmetic(uint64_t M, double p, ivec* o) { /* cbd:1667-1703 */
    int64_t T = target_count(M, p);
    if (M == 0 || T <= 0) return;
    uint64_t stride = M / (uint64_t)T;
    if (stride == 0) stride = 1;
    for (int64_t i = 0; i < T; ++i) {
        uint64_t off = (uint64_t)i * stride;
        if (off < M) push(o, (int64_t)off);
    }
}

static void gen_all(uint64_t N, ivec* o) { for (uint64_t i = 0; i < N; ++i) push(o, (int64_t)i); }

static void gen_index_based(uint64_t N, double p, ivec* o) { /* cbd:444-487 */
    if (N == 0 || p <= 0.0) return;
    if (p >= 100.0) { gen_all(N, o); return; }
    uint64_t T = (uint64_t)((double)N * p / 100.0);
    if (T == 0) return;
    double step = (double)N / (double)T;
    /* walk: take row `cur` when cur >= (size_t)(|S| * step) (cbd:470) */
    uint64_t taken = 0;
    for (uint64_t cur = 0; cur < N && taken < T; ++cur) {
        if (cur >= (uint64_t)((double)taken * step)) { push(o, (int64_t)cur); ++taken; }
    }
}

static void gen_byte_offset(uint64_t N, double p, ivec* o) { /* cbd:1461-1481 */
    if (N == 0) return;
    if (target_count(N, p) <= 0) return;
    gen_index_based(N, p, o);
}

static void gen_optimized_clt(uint64_t N, double p, int64_t Th, ivec* o) { /* cbd:1046-1147 */
    if (N == 0) return;
    uint64_t T = (uint64_t)((double)N * p / 100.0);
    if (p <= 0.0 || T == 0) return;
    int64_t opt = imin(Th, imax(1, (int64_t)(int)(T / 100)));
    if (N < 5000 || T < 200 || opt == 1) {
        uint64_t step = N / T; if (step < 1) step = 1;
        uint64_t c = 0;
        for (uint64_t i = 0; i < N && c < T; i += step) { push(o, (int64_t)i); ++c; }
        return;
    }
    uint64_t spt = T / (uint64_t)opt;
    uint64_t rpt = N / (uint64_t)opt;
    for (int64_t t = 0; t < opt; ++t) {
        uint64_t a = (uint64_t)t * rpt, b = (t == opt - 1) ? N : (uint64_t)(t + 1) * rpt;
        uint64_t lt = (t == opt - 1) ? (T - (uint64_t)(opt - 1) * spt) : spt;
        if (lt == 0) continue;
        uint64_t stride = (b - a) / lt; if (stride < 1) stride = 1;
        uint64_t c = 0;
        for (uint64_t i = a; i < b && c < lt; i += stride) { push(o, (int64_t)i); ++c; }
    }
}

static void gen_block(uint64_t N, double p, uint64_t B, ivec* o) { /* cbd:1151-1181; page: cbd:1183-1216 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0 || B == 0) return;
    uint64_t nb = (N + B - 1) / B;
    uint64_t k = (uint64_t)((double)nb * p / 100.0); if (k < 1) k = 1;
    uint64_t iv = nb / k; if (iv == 0) iv = 1;
    for (uint64_t b = 0; b < nb && (int64_t)o->n < T; b += iv) {
        uint64_t s = b * B, e = s + B < N ? s + B : N;
        for (uint64_t i = s; i < e && (int64_t)o->n < T; ++i) push(o, (int64_t)i);
    }
}

static int gen_parallel_block(uint64_t N, double p, uint64_t B, int64_t Th, ivec* o) { /* cbd:1218-1271 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return 0;
    if (Th <= 0 || B == 0) return AQE_ERR_INVALID;
    uint64_t nb = (N + B - 1) / B;
    uint64_t k = (uint64_t)((double)nb * p / 100.0); if (k < 1) k = 1;
    uint64_t bpt = k / (uint64_t)Th; if (bpt == 0) bpt = 1;
    uint64_t iv = nb / k; if (iv == 0) iv = 1;
    uint64_t tt = (uint64_t)(T / Th);
    for (int64_t t = 0; t < Th; ++t) {
        uint64_t sb = (uint64_t)t * bpt, eb = sb + bpt < k ? sb + bpt : k, c = 0;
        for (uint64_t j = sb; j < eb && c < tt; ++j) {
            uint64_t s = j * iv * B, e = s + B < N ? s + B : N;
            for (uint64_t i = s; i < e && c < tt; ++i) { push(o, (int64_t)i); ++c; }
        }
    }
    return 0;
}

static int gen_node_skip(uint64_t N, double p, int64_t skip, ivec* o) { /* cbd:489-532 */
    if (N == 0 || p <= 0.0) return 0;
    if (p >= 100.0) { gen_all(N, o); return 0; }
    if (skip <= 0) return AQE_ERR_INVALID;
    uint64_t T = (uint64_t)((double)N * p / 100.0);
    uint64_t L = leaves_for(N);
    for (uint64_t j = 0; j < L && o->n < T; ++j) {
        if (((int64_t)(j + 1) % skip) != 0) continue; /* node_counter is 1-based (cbd:510-512) */
        uint64_t kc = leaf_size(N, L, j);
        uint64_t take = T - o->n < kc ? T - o->n : kc;
        for (uint64_t i = 0; i < take; ++i) push(o, (int64_t)(127 * j + i));
    }
    return 0;
}

static void gen_direct_access(uint64_t N, double p, ivec* o) { /* cbd:584-644 */
    if (N == 0 || p <= 0.0) return;
    if (p >= 100.0) { gen_all(N, o); return; }
    uint64_t T = (uint64_t)((double)N * p / 100.0);
    uint64_t L = leaves_for(N);
    uint64_t nodes = T / 10 > 1 ? T / 10 : 1;
    double node_step = (double)L / (double)nodes;
    for (uint64_t i = 0; i < nodes && o->n < T; ++i) {
        uint64_t ni = (uint64_t)((double)i * node_step);
        if (ni >= L) continue;
        int kc = (int)leaf_size(N, L, ni);
        int rpn = (int)(T / nodes) > 1 ? (int)(T / nodes) : 1;
        if (rpn > kc) rpn = kc;
        double rs = (double)kc / rpn;
        for (int j = 0; j < rpn && o->n < T; ++j) {
            int ri = (int)(j * rs);
            if (ri < kc) push(o, (int64_t)(127 * ni + (uint64_t)ri));
        }
    }
}

/* balanced_tree_sample cbd:534-582: proportional allocation down the bulk-load shape.
 * Level l node g covers children [128 g, ...) of level l-1 (last node takes the rest). */
typedef struct { uint64_t N, L, T; uint64_t cnt[16]; int levels; ivec* o; } bt_ctx;
static uint64_t bt_rows_of(const bt_ctx* c, int level, uint64_t g, uint64_t* first_row) {
    /* rows spanned by node g of `level` (level 0 = leaves) */
    uint64_t lo = g, hi = g + 1; /* child range at level `level` */
    for (int l = level; l > 0; --l) {
        uint64_t nl = c->cnt[l - 1];
        uint64_t last = c->cnt[l] - 1;
        lo = 128 * lo;
        hi = (hi - 1 == last) ? nl : 128 * hi;
        (void)last;
    }
    *first_row = 127 * lo;
    return (hi == c->L ? c->N : 127 * hi) - 127 * lo;
}
static void bt_rec(bt_ctx* c, int level, uint64_t g, uint64_t want) {
    if (c->o->n >= c->T || want == 0) return;
    if (level == 0) {
        uint64_t kc = leaf_size(c->N, c->L, g);
        int take = (int)want < (int)kc ? (int)want : (int)kc;
        double step = (double)(int)kc / take;
        for (int i = 0; i < take && c->o->n < c->T; ++i) {
            int idx = (int)(i * step);
            if (idx < (int)kc) push(c->o, (int64_t)(127 * g + (uint64_t)idx));
        }
        return;
    }
    uint64_t fr, node_rows = bt_rows_of(c, level, g, &fr);
    uint64_t nchild_level = c->cnt[level - 1];
    uint64_t c0 = 128 * g, c1 = (g == c->cnt[level] - 1) ? nchild_level : 128 * (g + 1);
    for (uint64_t ch = c0; ch < c1 && c->o->n < c->T; ++ch) {
        uint64_t cfr, crow = bt_rows_of(c, level - 1, ch, &cfr);
        if (crow > 0) bt_rec(c, level - 1, ch, (want * crow) / node_rows);
    }
}
static void gen_balanced_tree(uint64_t N, double p, ivec* o) {
    if (N == 0 || p <= 0.0) return;
    if (p >= 100.0) { gen_all(N, o); return; }
    bt_ctx c; memset(&c, 0, sizeof(c));
    c.N = N; c.L = leaves_for(N); c.T = (uint64_t)((double)N * p / 100.0); c.o = o;
    c.cnt[0] = c.L; c.levels = 1;
    while (c.cnt[c.levels - 1] > 1) { c.cnt[c.levels] = parents_for(c.cnt[c.levels - 1]); c.levels++; }
    bt_rec(&c, c.levels - 1, 0, c.T);
}

/* adaptive_block_sample cbd:1273-1329 (data dependent: 10 zone variances) */
static void gen_adaptive_block(const aqe_record* R, uint64_t N, double p, uint64_t mn, uint64_t mx, ivec* o) {
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return;
    const uint64_t zones = 10;
    uint64_t zs = N / zones;
    double var[10], maxvar = -INFINITY;
    for (uint64_t z = 0; z < zones; ++z) {
        uint64_t s = z * zs, e = s + zs < N ? s + zs : N;
        double sum = 0.0, sq = 0.0;
        for (uint64_t i = s; i < e; ++i) { sum += R[i].amount; sq += R[i].amount * R[i].amount; }
        uint64_t cnt = e - s;
        double mean = sum / (double)cnt;
        var[z] = (sq / (double)cnt) - (mean * mean);
        if (var[z] > maxvar) maxvar = var[z];
    }
    for (uint64_t z = 0; z < zones && (int64_t)o->n < T; ++z) {
        uint64_t s = z * zs, e = s + zs < N ? s + zs : N;
        double ratio = var[z] / maxvar;
        uint64_t bs = mn + (uint64_t)((double)(mx - mn) * (1.0 - ratio));
        if (bs == 0) return; /* reference would spin forever */
        for (uint64_t i = s; i < e && (int64_t)o->n < T; i += bs) {
            uint64_t be = i + bs < e ? i + bs : e;
            uint64_t bc = (uint64_t)((double)(be - i) * p / 100.0); if (bc < 1) bc = 1;
            for (uint64_t j = 0; j < bc && i + j < be && (int64_t)o->n < T; ++j) push(o, (int64_t)(i + j));
        }
    }
}

/* stratified_block_sample cbd:1331-1379: positions are into the amount-sorted order. */
static int gen_stratified_block(uint64_t N, double p, uint64_t B, int64_t K, ivec* o) {
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0) return 0;
    if (K <= 0 || B == 0) return AQE_ERR_INVALID;
    uint64_t ssz = N / (uint64_t)K, sps = (uint64_t)(T / K);
    for (int64_t s = 0; s < K && (int64_t)o->n < T; ++s) {
        uint64_t a = (uint64_t)s * ssz, b = (s == K - 1) ? N : a + ssz;
        uint64_t recs = b - a, nb = (recs + B - 1) / B;
        uint64_t k = (uint64_t)((double)nb * p / 100.0); if (k < 1) k = 1;
        uint64_t iv = nb / k; if (iv == 0) iv = 1;
        for (uint64_t bi = 0; bi < nb && (int64_t)o->n < T; bi += iv) {
            uint64_t bs = a + bi * B, be = bs + B < b ? bs + B : b;
            uint64_t rem = sps < (uint64_t)T - o->n ? sps : (uint64_t)T - o->n;
            uint64_t take = rem < be - bs ? rem : be - bs;
            for (uint64_t i = 0; i < take; ++i) push(o, (int64_t)(bs + i));
        }
    }
    return 0;
}

/* ---- seeded stand-ins for the std::random_device methods.  The reference draws its seed from
 * random_device, so only the *shape* is pinned (same-index-list and in-distribution tests); the engine
 * replaces the one random draw by Philox(key=seed, ctr=(draw,0,method,0)). ------------------------ */
static uint64_t seeded_u64(uint64_t seed, uint32_t method, uint64_t draw) {
    uint32_t r[4];
    philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), 0x53454544u /* "SEED" */, method, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    return ((uint64_t)r[1] << 32) | r[0];
}
static uint64_t seeded_below(uint64_t seed, uint32_t method, uint64_t draw, uint64_t bound) { /* [0,bound) */
    return (uint64_t)(((unsigned __int128)seeded_u64(seed, method, draw) * bound) >> 64);
}

static void gen_random_start_nth(uint64_t N, double p, int64_t nth, uint64_t seed, ivec* o) { /* cbd:1483-1524 */
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0 || nth <= 0) return;
    uint64_t start = seeded_below(seed, AQE_M_RANDOM_START_NTH, 0, N);
    for (uint64_t i = start; (int64_t)o->n < T && i < N; i += (uint64_t)nth) push(o, (int64_t)i);
    if ((int64_t)o->n < T)
        for (uint64_t i = 0; i < start && (int64_t)o->n < T; i += (uint64_t)nth) push(o, (int64_t)i);
}

static void gen_address_arithmetic(uint64_t N, double p, uint64_t seed, ivec* o) { /* cbd:1605-1665 */
    int64_t T = target_count(N, p); /* from total_records (cbd:1619) */
    if (N == 0 || T <= 0) return;
    uint64_t M = cache_rows(N);     /* stride and wrap use the cache (cbd:1653-1659) */
    uint64_t stride = M / (uint64_t)T; if (stride == 0) stride = 1;
    for (int64_t i = 0; i < T; ++i) {
        uint64_t off = seeded_below(seed, AQE_M_ADDRESS_ARITHMETIC, (uint64_t)i, stride / 2 + 1);
        push(o, (int64_t)(((uint64_t)i * stride + off) % M));
    }
}

static void gen_random_start_memory_stride(uint64_t M, double p, int64_t stride_bytes, uint64_t seed, ivec* o) {
    int64_t T = target_count(M, p); /* cbd:1838-1878 */
    if (M == 0 || T <= 0) return;
    uint64_t stride = stride_bytes == 0 ? (uint64_t)imax(1, (int64_t)(M / (uint64_t)T))
                                        : (uint64_t)imax(1, stride_bytes / 32);
    uint64_t start = seeded_below(seed, AQE_M_RANDOM_START_MEMORY_STRIDE, 0, stride);
    for (uint64_t off = start; (int64_t)o->n < T && off < M; off += stride) push(o, (int64_t)off);
}

static int gen_multithreaded_memory_stride(uint64_t M, double p, int64_t Th, uint64_t seed, ivec* o) {
    if (M == 0) return 0; /* cbd:1880-1960; same index sets as fast_aggregated cbd:1962-2048 */
    if (Th <= 0) return AQE_ERR_INVALID;
    double pp = p / (double)Th;
    uint64_t rs = M / (uint64_t)Th, rem = M % (uint64_t)Th;
    for (int64_t t = 0; t < Th; ++t) {
        uint64_t a = (uint64_t)t * rs, len = rs + ((uint64_t)t < rem ? 1 : 0), b = a + len;
        if (a >= M) continue;
        if (b > M) b = M;
        uint64_t rt = b - a;
        uint64_t tt = (uint64_t)((double)rt * pp / 100.0);
        if (tt == 0) continue;
        uint64_t span = rt / 10 < 100 ? rt / 10 : 100;
        uint64_t start = a + seeded_below(seed, AQE_M_MULTITHREADED_MEMORY_STRIDE, (uint64_t)t, span + 1);
        uint64_t stride = rt / tt; if (stride == 0) stride = 1;
        uint64_t c = 0;
        for (uint64_t off = start; off < b && c < tt; off += stride) { push(o, (int64_t)off); ++c; }
    }
    return 0;
}

static void gen_optimized_sequential(uint64_t N, double p, uint64_t seed, ivec* o) { /* cbd:366-428 */
    if (p >= 100.0) { gen_all(N, o); return; }
    if (p <= 0.0) return;
    uint64_t T = (uint64_t)((double)N * p / 100.0);
    if (T == 0) return;
    double step = 100.0 / p;
    double start = step * ((double)(seeded_u64(seed, AQE_M_OPTIMIZED_SEQUENTIAL, 0) >> 11) * (1.0 / 9007199254740992.0));
    double next = start;
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < N && o->n < T; ++i) {
        ++cnt;
        if ((double)cnt >= next) { push(o, (int64_t)i); next += step; }
    }
}

/* sample_records cbd:345-363: SRSWOR of floor(N*p/100) rows.  The reference shuffles a copy of the table with a
 * random_device-seeded mt19937, so only the design is pinned (k distinct rows, each k-subset equally likely up to
 * the quality of the permutation).  Engine design restated here: position j = perm(j), perm = 4-round balanced Feistel
 * network on the smallest even bit width covering N, splitmix64 round function, cycle-walked into [0, N). */
static uint64_t feistel_mix(uint64_t r, uint32_t round, uint64_t seed) {
    uint64_t z = (r + 0x9E3779B97F4A7C15ull * (uint64_t)(round + 1)) ^ seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint64_t feistel_index(uint64_t k, uint64_t n, uint64_t seed) {
    uint32_t half = 1;
    while (half < 32 && (1ull << (2 * half)) < n) ++half;
    const uint64_t mask = (1ull << half) - 1;
    uint64_t x = k;
    do {
        uint64_t l = x >> half, r = x & mask;
        for (uint32_t i = 0; i < 4; ++i) { uint64_t t = l ^ (feistel_mix(r, i, seed) & mask); l = r; r = t; }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}
static int gen_sample_records(uint64_t N, double p, uint64_t seed, ivec* o) {
    if (N == 0) return 0;
    if (p >= 100.0) { gen_all(N, o); return 0; }
    if (p <= 0.0) return 0;
    uint64_t k = (uint64_t)((double)N * p / 100.0);
    if (k > N) k = N;
    for (uint64_t i = 0; i < k; ++i) push(o, (int64_t)feistel_index(i, N, seed));
    return 0;
}

static double z_table(double conf) { return conf >= 0.99 ? 2.576 : (conf >= 0.95 ? 1.96 : 1.645); } /* cbd:911-912 */

/* clt_validated_dual_pointer_sample cbd:885-1043 under the LOCK-STEP schedule: every live thread takes
 * its k-th sample at step k, fast threads act before slow threads within a step and in thread order.
 * This is one legal interleaving of the reference's racy std::async threads; the reference's output is a
 * set of prefixes of the same per-thread stride sequences. */
static int gen_clt_validated(const aqe_record* R, uint64_t N, const aqe_sample_params* P, ivec* o) {
    int64_t T = target_count(N, P->sample_percent);
    if (N == 0 || T <= 0) return 0;
    int64_t Th = P->num_threads, ci = P->check_interval;
    int64_t F = Th / 2, S = Th - F;
    if (F <= 0 || S <= 0 || ci < 2 || T / F == 0 || T / S == 0) return AQE_ERR_INVALID; /* reference: div by 0 */
    double z = z_table(P->confidence_level), maxerr = P->max_error_percent;
    int64_t nt = F + S;
    uint64_t *st = calloc(nt, 8), *first = calloc(nt, 8), *len = calloc(nt, 8);
    for (int64_t q = 0; q < nt; ++q) {
        int fast = q < F;
        uint64_t t = (uint64_t)(fast ? q : q - F), G = (uint64_t)(fast ? F : S);
        uint64_t a = (N * t) / G, b = (N * (t + 1)) / G;                       /* cbd:925-926 / 979-980 */
        int64_t s = (int64_t)(int)((b - a) / (uint64_t)(T / (int64_t)G));
        st[q] = (uint64_t)(fast ? imax(3, s) : imax(1, s));                     /* cbd:927 / 981 */
        first[q] = fast ? a : a + st[q] / 2;                                    /* cbd:984 */
        len[q] = first[q] < b ? (b - first[q] + st[q] - 1) / st[q] : 0;
    }
    uint64_t maxlen = 0;
    for (int64_t q = 0; q < nt; ++q) if (len[q] > maxlen) maxlen = len[q];
    double current_mean = 0.0;
    uint64_t sample_count = 0, kstop = 0;
    int64_t stopper = -1;
    for (uint64_t k = 1; k <= maxlen && stopper < 0; ++k) {
        for (int64_t q = 0; q < nt && stopper < 0; ++q) {
            if (k > len[q]) continue;
            int fast = q < F;
            uint64_t every = (uint64_t)(fast ? ci : ci / 2), minn = fast ? 30 : 20;  /* cbd:936 / 993 */
            if (k % every != 0 || k < minn) continue;
            double mean = 0.0;
            for (uint64_t j = 0; j < k; ++j) mean += R[first[q] + j * st[q]].amount;
            mean /= (double)k;
            double var = 0.0;
            for (uint64_t j = 0; j < k; ++j) { double d = R[first[q] + j * st[q]].amount - mean; var += d * d; }
            var /= (double)(k - 1);
            if (fast) {
                current_mean = mean; sample_count = k;                             /* cbd:949-951 */
                double se = sqrt(var / (double)k), moe = z * se, ep = (moe / mean) * 100.0;
                if (ep <= maxerr && k >= 50) { stopper = q; kstop = k; }           /* cbd:958-961 */
            } else if (current_mean > 0) {
                double md = fabs(mean - current_mean) / current_mean;              /* cbd:1008 */
                if (md <= maxerr / 100.0 && sample_count >= (uint64_t)(T / 2)) { stopper = q; kstop = k; }
            }
        }
    }
    /* threads up to and including the stopper hold kstop samples, later ones saw the flag before
     * taking sample kstop (loop test `!should_stop.load()`, cbd:930/987) */
    for (int64_t q = 0; q < nt; ++q) {
        uint64_t take = len[q];
        if (stopper >= 0) {
            uint64_t lim = q <= stopper ? kstop : kstop - 1;
            if (take > lim) take = lim;
        }
        for (uint64_t j = 0; j < take; ++j) push(o, (int64_t)(first[q] + j * st[q]));
    }
    if ((int64_t)o->n < T / 4) { /* cbd:1032-1040 top-up */
        int64_t add = T / 4;
        uint64_t step = (uint64_t)imax(1, (int64_t)(int)(N / (uint64_t)add));
        for (uint64_t i = 0; i < N && (int64_t)o->n < T; i += step) push(o, (int64_t)i);
    }
    free(st); free(first); free(len);
    return 0;
}

/* signal_based_clt_sample cbd:1705-1818, lock-step: the fast thread alone decides the stop (shared counter
 * counts only fast samples, cbd:1738); the slow thread has taken min(kstop, T/4) rows by then. */
static void gen_signal_based(uint64_t N, double p, int64_t ci, ivec* o) {
    int64_t T = target_count(N, p);
    if (N == 0 || T <= 0 || ci <= 0) return;
    uint64_t fs = N / (uint64_t)(T * 2); if (fs < 2) fs = 2;
    uint64_t nf = 0;
    for (uint64_t i = 0; i < N && (int64_t)nf < T; i += fs) {
        push(o, (int64_t)i); ++nf;
        if (nf % (uint64_t)ci == 0 && (int64_t)nf >= T / 2) break;
    }
    uint64_t ns = 0;
    for (uint64_t i = 0; i < N && ns < nf && (int64_t)ns < T / 4; ++i) { push(o, (int64_t)i); ++ns; }
    if ((int64_t)o->n > T) o->n = (uint64_t)T;
}

/* Dispatcher: writes up to cap positions, returns the full count (or -status). */
ORC_API int64_t orc_indices(const aqe_record* R, uint64_t N, int method, const aqe_sample_params* P, int64_t* out,
                            uint64_t cap) {
    ivec o = {out, 0, cap};
    double p = P->sample_percent;
    int rc = 0;
    switch (method) {
        case AQE_M_SLOW_POINTER: gen_slow_pointer(N, p, 1, &o); break;
        case AQE_M_FAST_POINTER:
            if (P->step_size <= 0) return -AQE_ERR_INVALID;
            gen_slow_pointer(N, p, P->step_size, &o); break;
        case AQE_M_DUAL_POINTER: rc = gen_dual_pointer(N, p, &o); break;
        case AQE_M_PARALLEL_POINTER: rc = gen_parallel_pointer(N, p, P->num_threads, &o); break;
        case AQE_M_RANDOM_POINTER: rc = gen_random_pointer(N, p, (uint32_t)P->seed, &o); break;
        case AQE_M_MEMORY_STRIDE: gen_memory_stride(cache_rows(N), p, P->block_size, 0, &o); break;
        case AQE_M_OPT_ADDRESS_ARITHMETIC: gen_opt_address_arithmetic(cache_rows(N), p, &o); break;
        case AQE_M_INDEX_BASED: gen_index_based(N, p, &o); break;
        case AQE_M_BYTE_OFFSET: gen_byte_offset(N, p, &o); break;
        case AQE_M_OPTIMIZED_CLT: gen_optimized_clt(N, p, P->num_threads, &o); break;
        case AQE_M_BLOCK: gen_block(N, p, (uint64_t)P->block_size, &o); break;
        case AQE_M_PAGE: { uint64_t rpp = (uint64_t)P->block_size / 32; if (rpp == 0) rpp = 1; gen_block(N, p, rpp, &o); break; }
        case AQE_M_PARALLEL_BLOCK: rc = gen_parallel_block(N, p, (uint64_t)P->block_size, P->num_threads, &o); break;
        case AQE_M_NODE_SKIP: rc = gen_node_skip(N, p, P->step_size, &o); break;
        case AQE_M_BALANCED_TREE: gen_balanced_tree(N, p, &o); break;
        case AQE_M_DIRECT_ACCESS: gen_direct_access(N, p, &o); break;
        case AQE_M_ADAPTIVE_BLOCK: gen_adaptive_block(R, N, p, (uint64_t)P->block_size, (uint64_t)P->block_size_max, &o); break;
        case AQE_M_STRATIFIED_BLOCK: rc = gen_stratified_block(N, p, (uint64_t)P->block_size, P->block_size_max, &o); break;
        case AQE_M_SAMPLE_RECORDS: rc = gen_sample_records(N, p, P->seed, &o); break;
        case AQE_M_OPTIMIZED_SEQUENTIAL: gen_optimized_sequential(N, p, P->seed, &o); break;
        case AQE_M_RANDOM_START_NTH: gen_random_start_nth(N, p, P->step_size, P->seed, &o); break;
        case AQE_M_ADDRESS_ARITHMETIC: gen_address_arithmetic(N, p, P->seed, &o); break;
        case AQE_M_RANDOM_START_MEMORY_STRIDE: gen_random_start_memory_stride(cache_rows(N), p, P->block_size, P->seed, &o); break;
        case AQE_M_MULTITHREADED_MEMORY_STRIDE: rc = gen_multithreaded_memory_stride(cache_rows(N), p, P->num_threads, P->seed, &o); break;
        case AQE_M_CLT_VALIDATED_DUAL_POINTER: rc = gen_clt_validated(R, N, P, &o); break;
        case AQE_M_SIGNAL_BASED_CLT: gen_signal_based(cache_rows(N), p, P->check_interval, &o); break;
        default: return -AQE_ERR_UNSUPPORTED;
    }
    if (rc) return -rc;
    return (int64_t)o.n;
}

ORC_API void orc_params_default(aqe_sample_params* p, int method) { /* bindings.cpp:56-101 */
    memset(p, 0, sizeof(*p));
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

/* ================================================================================================
 * 6. Estimators over a sample (the Python loops of the CLI) -- cli:188-200 (random) and cli:257-291 (clt)
 * ============================================================================================== */
/* Python's built-in sum() over floats: since CPython 3.12 (Python/bltinmodule.c, builtin_sum) this is
 * Neumaier compensated summation with the correction added once at the end; before 3.12 it was a plain
 * left-to-right loop.  The CLI's estimators are `sum(...)` generator expressions (cli:190, 262, 277, 279),
 * so their bits depend on the interpreter; this image runs 3.12 and the golden vectors were minted on it.
 * py312 != 0 restates the 3.12 algorithm, 0 the plain loop. */
typedef struct { double total, c; int py312; } pysum;
static inline void pysum_add(pysum* s, double x) {
    if (!s->py312) { s->total += x; return; }
    double t = s->total + x;
    if (fabs(s->total) >= fabs(x)) s->c += (s->total - t) + x; else s->c += (x - t) + s->total;
    s->total = t;
}
static inline double pysum_get(const pysum* s) {
    if (s->py312 && s->c != 0.0 && isfinite(s->c)) return s->total + s->c;
    return s->total;
}

static int g_py312 = 1;
ORC_API void orc_set_python_sum(int py312) { g_py312 = py312; }

ORC_API void orc_stats(const aqe_record* R, const int64_t* idx, uint64_t n, int col, aqe_stats* s) {
    pysum a = {0.0, 0.0, g_py312};
    for (uint64_t i = 0; i < n; ++i) pysum_add(&a, col_as_double(&R[idx[i]], col)); /* cli:190/262 */
    double sum = pysum_get(&a);
    double mean = n ? sum / (double)n : 0.0;                                         /* cli:278 */
    pysum b = {0.0, 0.0, g_py312};
    for (uint64_t i = 0; i < n; ++i) { double d = col_as_double(&R[idx[i]], col) - mean; pysum_add(&b, d * d); } /* cli:279 */
    s->n = n; s->sum = sum; s->mean = mean; s->m2 = pysum_get(&b);
}

/* agg: SUM -> sum*(N/n) (cli:192-193), AVG -> sum/n (cli:195), COUNT -> N (cli:197).
 * CI: MoE = z*s/sqrt(n), s^2 = m2/(n-1) (cli:278-281).  legacy_ci: SUM margin = MoE*(N/n) (cli:285, the
 * reference's too-narrow interval); otherwise the correct MoE*N. */
ORC_API void orc_estimate(const aqe_stats* s, uint64_t N, int agg, double z, int legacy_ci, double* est, double* lo,
                          double* hi) {
    double n = (double)s->n;
    double e;
    if (agg == AQE_AGG_SUM) e = s->sum * ((double)N / n);
    else if (agg == AQE_AGG_COUNT) e = (double)N;
    else e = s->sum / n;
    double var = s->n > 1 ? s->m2 / (n - 1.0) : 0.0;
    double sd = pow(var, 0.5);
    double moe = z * sd / pow(n, 0.5);
    double m;
    if (agg == AQE_AGG_SUM) m = legacy_ci ? moe * ((double)N / n) : moe * (double)N;
    else if (agg == AQE_AGG_COUNT) m = 0.0;
    else m = moe;
    *est = e; *lo = e - m; *hi = e + m;
}

/* fast_aggregated_memory_stride_sum cbd:1962-2048 given the same per-thread start draws: raw sample sum,
 * per-thread serial sums added in thread order (the reference's CAS order is racy). */
ORC_API double orc_fast_aggregated(const aqe_record* R, uint64_t N, const aqe_sample_params* P, uint64_t* n_out) {
    uint64_t M = cache_rows(N);
    int64_t Th = P->num_threads;
    double total = 0.0; uint64_t cnt = 0;
    if (M == 0 || Th <= 0) { if (n_out) *n_out = 0; return 0.0; }
    double pp = P->sample_percent / (double)Th;
    uint64_t rs = M / (uint64_t)Th, rem = M % (uint64_t)Th;
    for (int64_t t = 0; t < Th; ++t) {
        uint64_t a = (uint64_t)t * rs, len = rs + ((uint64_t)t < rem ? 1 : 0), b = a + len;
        if (a >= M) continue;
        if (b > M) b = M;
        uint64_t rt = b - a, tt = (uint64_t)((double)rt * pp / 100.0);
        if (tt == 0) continue;
        uint64_t span = rt / 10 < 100 ? rt / 10 : 100;
        uint64_t start = a + seeded_below(P->seed, AQE_M_MULTITHREADED_MEMORY_STRIDE, (uint64_t)t, span + 1);
        uint64_t stride = rt / tt; if (stride == 0) stride = 1;
        double ts = 0.0; uint64_t c = 0;
        for (uint64_t off = start; off < b && c < tt; off += stride) { ts += R[off].amount; ++c; }
        total += ts; cnt += c;
    }
    if (n_out) *n_out = cnt;
    return cnt ? total : 0.0;
}

/* ================================================================================================
 * 7. Persistent CLT estimator (K4) restated.  Not a reference algorithm -- the reference's racy thread
 *    pool (cbd:885-1043) is REPLACED by this design (north_star item 3); restated here so the device
 *    kernel has a bit-for-bit sample-set checker.  Sample j of the stream uses Philox(key=seed,
 *    ctr=(j>>1, 0x53525330 "SRS0" | design, 0)) and the (j&1) 64-bit lane; position = mulhi64(u64, units).
 *    Looks happen at cumulative sizes n_0 = min_samples, then n_{r+1} = clamp(ceil(1.1 * n_req),
 *    n_r + n_r/4 + 1, 8 n_r), n_req = (z s 100 / (eps |mean|))^2, capped by max_samples.
 * ============================================================================================== */
ORC_API double orc_z_score(double conf, int exact) {
    if (!exact) return z_table(conf);
    /* Acklam's inverse normal CDF, p = 1 - (1-conf)/2 */
    double p = 1.0 - (1.0 - conf) / 2.0;
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double q, r;
    if (p < 0.02425) {
        q = sqrt(-2 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5; r = q * q;
        return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
               (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    }
    q = sqrt(-2 * log(1 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
}

/* The interval of an aqe_ci_mode (include/aqe_b200.h): normal quantile at the (guarded) level, Student-t through Fisher's
 * expansion, and -- Stein type -- the variance of the PREVIOUS look at its degrees of freedom. */
static double ci_z(double conf, uint32_t ci_mode, int* stein) {
    uint32_t m = ci_mode == AQE_CI_DEFAULT ? (uint32_t)AQE_CI_STEIN_GUARDED : ci_mode;
    *stein = m != AQE_CI_PLAIN;
    return orc_z_score(1.0 - (m == AQE_CI_STEIN_GUARDED ? AQE_CI_GUARD : 1.0) * (1.0 - conf), 1);
}
static double t_from_z(double z, double df) {
    if (!(df > 4.0)) df = 4.0;
    double z2 = z * z;
    return z + z * (z2 + 1.0) / (4.0 * df) + z * ((5.0 * z2 + 16.0) * z2 + 3.0) / (96.0 * df * df);
}

static inline uint64_t draw_position(uint64_t seed, uint32_t design, uint64_t j, uint64_t units) {
    uint32_t r[4];
    uint64_t c = j >> 1;
    philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), 0x53525330u | design, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    uint64_t u = (j & 1) ? (((uint64_t)r[3] << 32) | r[2]) : (((uint64_t)r[1] << 32) | r[0]);
    return (uint64_t)(((unsigned __int128)u * units) >> 64);
}

ORC_API uint64_t orc_draw_position(uint64_t seed, uint32_t design, uint64_t j, uint64_t units) {
    return draw_position(seed, design, j, units);
}

/* value of sampling unit `u` : SRS -> y(row u); BLOCK -> sum over the tile's rows of y(row).
 * y = x (SUM/AVG, no predicate), x*1[pred] (SUM with predicate), 1[pred] (COUNT with predicate).
 * AVG with predicate is a ratio estimator: numerator y = x*1[pred], denominator c = 1[pred]. */
static inline void unit_value(const aqe_record* R, uint64_t N, const aqe_approx_spec* S, uint64_t u, double* y, double* c) {
    uint64_t a = u, b = u + 1;
    if (S->design == AQE_DESIGN_BLOCK) {
        uint64_t B = S->block_size ? S->block_size : 1000;
        a = u * B; b = a + B < N ? a + B : N;
    }
    double ys = 0.0, cs = 0.0;
    for (uint64_t i = a; i < b; ++i) {
        int pass = 1;
        if (S->pred_col != AQE_COL_NONE) { double pv = col_as_double(&R[i], S->pred_col); pass = (pv >= S->lo && pv <= S->hi); }
        if (!pass) continue;
        cs += 1.0;
        ys += (S->agg == AQE_AGG_COUNT) ? 1.0 : col_as_double(&R[i], S->agg_col);
    }
    *y = ys; *c = cs;
}

ORC_API int orc_approx(const aqe_record* R, uint64_t N, const aqe_approx_spec* S, aqe_approx_result* out) {
    memset(out, 0, sizeof(*out));
    out->population = N; out->confidence_level = S->confidence_level;
    if (N == 0) { out->status = AQE_INSUFFICIENT_DATA; return 0; }
    uint64_t B = S->design == AQE_DESIGN_BLOCK ? (S->block_size ? S->block_size : 1000) : 1;
    uint64_t units = (N + B - 1) / B;
    int stein; double z = ci_z(S->confidence_level, S->ci_mode, &stein);
    if (S->agg == AQE_AGG_COUNT && S->pred_col == AQE_COL_NONE) { /* cli:197: COUNT is always exact */
        out->estimate = out->ci_lower = out->ci_upper = (double)N; out->status = AQE_STABLE; out->pass_fraction = 1.0; return 0;
    }
    uint64_t n0 = S->min_samples ? S->min_samples : (S->design == AQE_DESIGN_BLOCK ? 1024 : 16384);
    uint64_t nmax = S->max_samples ? S->max_samples : units;
    if (n0 > nmax) n0 = nmax;
    if (units <= n0) { /* the first look would draw as many units as the table has: exact scan instead */
        aqe_partial part;
        orc_scan(R, N, S->agg == AQE_AGG_COUNT ? S->pred_col : S->agg_col, S->pred_col, S->lo, S->hi, &part);
        double v = S->agg == AQE_AGG_COUNT ? (double)part.count : (S->agg == AQE_AGG_SUM ? part.sum : (part.count ? part.sum / (double)part.count : 0.0));
        out->estimate = out->ci_lower = out->ci_upper = v; out->n_samples = N; out->n_units = units; out->status = AQE_STABLE;
        out->pass_fraction = (double)part.count / (double)N;
        return 0;
    }
    int ratio = (S->agg == AQE_AGG_AVG && S->pred_col != AQE_COL_NONE);
    /* running sums in long double: the checker is allowed to be more exact than the device */
    long double sy = 0, syy = 0, sc = 0, rows = 0;
    uint64_t n = 0, target = n0; uint32_t rounds = 0;
    double est = 0, half = 0, rel = INFINITY, rel_next = INFINITY, mean = 0, m2 = 0;
    double var_prev = 0; uint64_t nvar_prev = 0; int have_prev = 0;
    for (;;) {
        for (; n < target; ++n) {
            uint64_t u = draw_position(S->seed, (uint32_t)S->design, n, units);
            double y, c; unit_value(R, N, S, u, &y, &c);
            sy += y; syy += (long double)y * y; sc += c;
            if (S->design == AQE_DESIGN_BLOCK) { uint64_t a = u * B; rows += (long double)((a + B < N ? a + B : N) - a); }
            else rows += 1;
        }
        ++rounds;
        long double nn = (long double)n;
        long double mu = sy / nn;
        long double ss = syy - sy * mu; if (ss < 0) ss = 0;
        mean = (double)mu; m2 = (double)ss;
        double var, scale;
        if (ratio) {
            /* R = sy/sc ; linearised residual sum of squares = syy - R*sy (c^2=c, y*c=y) */
            if (sc <= 0) { est = 0; rel = INFINITY; var = INFINITY; scale = 1; }
            else {
                long double Rr = sy / sc;
                long double rss = syy - Rr * sy; if (rss < 0) rss = 0;
                long double cbar = sc / nn;
                var = (double)(rss / (nn - 1) / (cbar * cbar));
                est = (double)Rr; scale = 1;
            }
        } else {
            var = n > 1 ? (double)(ss / (nn - 1)) : INFINITY;
            if (S->agg == AQE_AGG_AVG) { est = (double)(mu * (long double)units / (long double)N); scale = (double)units / (double)N; }
            else { est = (double)(mu * (long double)units); scale = (double)units; }
        }
        int use_prev = stein && have_prev && nvar_prev > 1 && var_prev > var;   /* the larger of the two variances */
        double v_use = use_prev ? var_prev : var;
        double df_use = use_prev ? (double)(nvar_prev - 1) : (double)n - 1.0;
        half = t_from_z(z, df_use) * sqrt(v_use / (double)n) * scale;
        double half_next = t_from_z(z, (double)n - 1.0) * sqrt(var / (double)n) * scale;
        rel = est != 0 ? half / fabs(est) * 100.0 : INFINITY;
        rel_next = est != 0 ? half_next / fabs(est) * 100.0 : INFINITY;
        var_prev = var; nvar_prev = n; have_prev = 1;
        if (rel <= S->error_percent) { out->status = AQE_STABLE; break; }
        if (n >= nmax) { out->status = AQE_DRIFTING; break; }
        double ratio_n = rel_next / S->error_percent;
        double nreq = (double)n * ratio_n * ratio_n;
        double want = ceil(1.1 * nreq);
        uint64_t lo_n = n + n / 4 + 1, hi_n = n * 8;
        uint64_t t2 = want >= (double)hi_n ? hi_n : (want <= (double)lo_n ? lo_n : (uint64_t)want);
        if (t2 > nmax) t2 = nmax;
        target = t2;
    }
    out->estimate = est; out->ci_lower = est - half; out->ci_upper = est + half;
    out->error_margin = rel / 100.0; out->n_units = n; out->n_samples = (uint64_t)rows; out->rounds = rounds;
    out->mean = mean; out->m2 = m2;
    out->pass_fraction = rows > 0 ? (double)(sc / rows) : 0.0;
    return 0;
}

/* Multi-GPU form of section 7 restated on one CPU: `world` contiguous shards [N g/G, N (g+1)/G) are strata with
 * proportional allocation; rank g draws from Philox key seed + g*0x9E3779B97F4A7C15; after every look the GLOBAL
 * stratified estimate  T = sum_g U_g mean_g,  Var = sum_g U_g^2 s_g^2 / n_g  decides (same growth rule on the global
 * cumulative size).  Checker for aqe_approx_exchange (tests/multi_gpu_check.py). */
static uint64_t share_of(uint64_t T, uint64_t ug, uint64_t utot) {
    if (ug == 0) return 0;
    double x = ceil((double)T * ((double)ug / (double)utot));
    uint64_t v = x >= (double)ug ? ug : (uint64_t)x;
    return v < 1 ? 1 : v;
}
ORC_API int orc_approx_sharded(const aqe_record* R, uint64_t N, int world, const aqe_approx_spec* S, aqe_approx_result* out) {
    memset(out, 0, sizeof(*out));
    out->population = N; out->confidence_level = S->confidence_level;
    if (N == 0) { out->status = AQE_INSUFFICIENT_DATA; return 0; }
    if (world < 1 || world > 16) return 1;
    uint64_t B = S->design == AQE_DESIGN_BLOCK ? (S->block_size ? S->block_size : 1000) : 1;
    int stein; double z = ci_z(S->confidence_level, S->ci_mode, &stein);
    if (S->agg == AQE_AGG_COUNT && S->pred_col == AQE_COL_NONE) {
        out->estimate = out->ci_lower = out->ci_upper = (double)N; out->status = AQE_STABLE; out->pass_fraction = 1.0; return 0;
    }
    uint64_t first[16], rows[16], units[16], n[16], target[16], utot = 0;
    long double sy[16], syy[16], sc[16], nrows[16];
    for (int g = 0; g < world; ++g) {
        first[g] = (uint64_t)(((unsigned __int128)N * g) / world);
        rows[g] = (uint64_t)(((unsigned __int128)N * (g + 1)) / world) - first[g];
        units[g] = (rows[g] + B - 1) / B; utot += units[g];
        n[g] = 0; sy[g] = syy[g] = sc[g] = nrows[g] = 0;
    }
    uint64_t n0 = S->min_samples ? S->min_samples : (S->design == AQE_DESIGN_BLOCK ? 1024 : 16384);
    uint64_t nmax = S->max_samples ? S->max_samples : utot;
    if (n0 > nmax) n0 = nmax;
    if (utot <= n0) {
        aqe_partial part;
        orc_scan(R, N, S->agg == AQE_AGG_COUNT ? S->pred_col : S->agg_col, S->pred_col, S->lo, S->hi, &part);
        double v = S->agg == AQE_AGG_COUNT ? (double)part.count : (S->agg == AQE_AGG_SUM ? part.sum : (part.count ? part.sum / (double)part.count : 0.0));
        out->estimate = out->ci_lower = out->ci_upper = v; out->n_samples = N; out->n_units = utot; out->status = AQE_STABLE;
        return 0;
    }
    int ratio = (S->agg == AQE_AGG_AVG && S->pred_col != AQE_COL_NONE);
    uint64_t Tg = n0; uint32_t rounds = 0;
    double est = 0, half = 0, rel = INFINITY, rel_next = INFINITY, passf = 0;
    uint64_t ntot = 0; long double rows_read = 0;
    long double var_prev[16]; uint64_t nvar_prev[16]; int have_prev = 0;
    for (int g = 0; g < 16; ++g) { var_prev[g] = 0; nvar_prev[g] = 0; }
    for (;;) {
        for (int g = 0; g < world; ++g) {
            uint64_t t = share_of(Tg, units[g], utot);
            target[g] = t > n[g] ? t : n[g];
            uint64_t seed = S->seed + 0x9E3779B97F4A7C15ull * (uint64_t)g;
            for (; n[g] < target[g]; ++n[g]) {
                uint64_t u = draw_position(seed, (uint32_t)S->design, n[g], units[g]);
                double y, c; unit_value(R + first[g], rows[g], S, u, &y, &c);
                sy[g] += y; syy[g] += (long double)y * y; sc[g] += c;
                if (S->design == AQE_DESIGN_BLOCK) { uint64_t a = u * B; nrows[g] += (long double)((a + B < rows[g] ? a + B : rows[g]) - a); }
                else nrows[g] += 1;
            }
        }
        ++rounds;
        long double T = 0, V = 0, Vcur = 0, num = 0, den = 0;
        double df_use = 0, df_cur = 0;
        int use_prev = stein && have_prev;
        ntot = 0; rows_read = 0;
        for (int g = 0; g < world; ++g) {
            ntot += n[g]; rows_read += nrows[g];
            if (!units[g] || !n[g]) continue;
            num += (long double)units[g] * (sy[g] / (long double)n[g]); den += (long double)units[g] * (sc[g] / (long double)n[g]);
        }
        long double Rr = den > 0 ? num / den : 0;
        for (int g = 0; g < world; ++g) {
            if (!units[g] || !n[g]) continue;
            long double nn = (long double)n[g], U = (long double)units[g], mu = sy[g] / nn, var;
            if (ratio) {
                long double se2 = syy[g] - 2 * Rr * sy[g] + Rr * Rr * sc[g];
                long double eb = (sy[g] - Rr * sc[g]) / nn;
                se2 -= nn * eb * eb; if (se2 < 0) se2 = 0;
                var = n[g] > 1 ? se2 / (nn - 1) : INFINITY;
            } else {
                long double ss = syy[g] - sy[g] * mu; if (ss < 0) ss = 0;
                var = n[g] > 1 ? ss / (nn - 1) : INFINITY;
            }
            int up = use_prev && nvar_prev[g] > 1 && var_prev[g] > var;
            long double v_use = up ? var_prev[g] : var;
            df_use += up ? (double)(nvar_prev[g] - 1) : (double)n[g] - 1.0;
            df_cur += (double)n[g] - 1.0;
            T += U * mu; V += U * U * v_use / nn; Vcur += U * U * var / nn;
            var_prev[g] = var; nvar_prev[g] = n[g];
        }
        have_prev = 1;
        double tq = t_from_z(z, df_use), tqn = t_from_z(z, df_cur), half_next;
        if (ratio) { est = (double)Rr; half = den > 0 ? (double)(tq * sqrtl(V) / den) : INFINITY; half_next = den > 0 ? (double)(tqn * sqrtl(Vcur) / den) : INFINITY; }
        else if (S->agg == AQE_AGG_AVG) { est = (double)(T / (long double)N); half = (double)(tq * sqrtl(V) / (long double)N); half_next = (double)(tqn * sqrtl(Vcur) / (long double)N); }
        else { est = (double)T; half = (double)(tq * sqrtl(V)); half_next = (double)(tqn * sqrtl(Vcur)); }
        passf = (double)(den / (long double)N);
        rel = (ratio && !(den > 0)) ? INFINITY : (est != 0 ? half / fabs(est) * 100.0 : INFINITY);
        rel_next = (ratio && !(den > 0)) ? INFINITY : (est != 0 ? half_next / fabs(est) * 100.0 : INFINITY);
        if (rel <= S->error_percent) { out->status = AQE_STABLE; break; }
        if (Tg >= nmax) { out->status = AQE_DRIFTING; break; }
        double rn = rel_next / S->error_percent;
        double want = ceil(1.1 * ((double)ntot * rn * rn));
        uint64_t lo_n = Tg + Tg / 4 + 1, hi_n = Tg * 8;
        uint64_t t2 = want >= (double)hi_n ? hi_n : (want <= (double)lo_n ? lo_n : (uint64_t)want);
        if (t2 > nmax) t2 = nmax;
        Tg = t2;
    }
    out->estimate = est; out->ci_lower = est - half; out->ci_upper = est + half; out->error_margin = rel / 100.0;
    out->n_units = ntot; out->n_samples = (uint64_t)rows_read; out->rounds = rounds; out->pass_fraction = passf;
    return 0;
}

/* ================================================================================================
 * 8. Restated multithreaded scan for the CPU baseline at sizes the reference cannot load (SURVEY D10,
 *    8d-ii): contiguous regions (the split of cbd:1984-2011 at 100 %), serial sum per region, ordered
 *    merge.  aos != 0 walks 32-byte rows (what the reference touches), else an 8-byte amount column.
 * ============================================================================================== */
typedef struct { const void* base; uint64_t a, b; int aos; int pred; double lo, hi; double sum; uint64_t cnt; } mt_job;
static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    double s = 0.0; uint64_t c = 0;
    if (j->aos) {
        const aqe_record* R = (const aqe_record*)j->base;
        for (uint64_t i = j->a; i < j->b; ++i) { double v = R[i].amount; if (!j->pred || (v >= j->lo && v <= j->hi)) { s += v; ++c; } }
    } else {
        const double* X = (const double*)j->base;
        for (uint64_t i = j->a; i < j->b; ++i) { double v = X[i]; if (!j->pred || (v >= j->lo && v <= j->hi)) { s += v; ++c; } }
    }
    j->sum = s; j->cnt = c;
    return NULL;
}
/* ---- checkers for the BASELINE sizes (10 M / 100 M / 1 B rows): the amount column of the synthetic table generated on all host
 * cores, the reference's strict left-to-right sum over it (cbd:242-251, :263-274 -- one thread, one double), and an effectively
 * exact sum (long double Neumaier accumulation per thread, 64 + 64 bits, folded in order) to hold the device to a few ulp. ---- */
typedef struct { uint64_t seed, first, a, b; int dist; double* out; } synth_job;
static void* synth_worker(void* arg) {
    synth_job* j = (synth_job*)arg;
    aqe_record r;
    for (uint64_t i = j->a; i < j->b; ++i) { orc_synth_row(j->seed, j->first + i, j->dist, &r); j->out[i] = r.amount; }
    return NULL;
}
ORC_API void orc_synth_amount_mt(uint64_t seed, uint64_t first_row, uint64_t n, int dist, double* out, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    synth_job* jobs = (synth_job*)malloc(sizeof(synth_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (synth_job){seed, first_row, n * (uint64_t)t / (uint64_t)threads, n * (uint64_t)(t + 1) / (uint64_t)threads, dist, out};
        pthread_create(&th[t], NULL, synth_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
/* sum_amount / sum_amount_where as the reference computes them: for (r : records) sum += r.amount  (one double, in id order) */
ORC_API double orc_sum_col_serial(const double* x, uint64_t n, int pred, double lo, double hi, uint64_t* count) {
    volatile double s = 0.0; uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) { double v = x[i]; if (!pred || (v >= lo && v <= hi)) { s = s + v; ++c; } }
    if (count) *count = c;
    return s;
}
typedef struct { const double* x; uint64_t a, b; int pred; double lo, hi; long double s, c; uint64_t cnt; } exact_job;
static void* exact_worker(void* arg) {
    exact_job* j = (exact_job*)arg;
    long double s = 0, c = 0; uint64_t cnt = 0;
    for (uint64_t i = j->a; i < j->b; ++i) {
        double v = j->x[i];
        if (j->pred && !(v >= j->lo && v <= j->hi)) continue;
        ++cnt;
        long double t = s + (long double)v;     /* Neumaier: the rounding error of every add is kept in c */
        if (fabsl(s) >= fabsl((long double)v)) c += (s - t) + (long double)v; else c += ((long double)v - t) + s;
        s = t;
    }
    j->s = s; j->c = c; j->cnt = cnt;
    return NULL;
}
ORC_API double orc_sum_col_exact_mt(const double* x, uint64_t n, int pred, double lo, double hi, int threads, uint64_t* count) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    exact_job* jobs = (exact_job*)malloc(sizeof(exact_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (exact_job){x, n * (uint64_t)t / (uint64_t)threads, n * (uint64_t)(t + 1) / (uint64_t)threads, pred, lo, hi, 0, 0, 0};
        pthread_create(&th[t], NULL, exact_worker, &jobs[t]);
    }
    long double s = 0, c = 0; uint64_t cnt = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        long double v = jobs[t].s, tt = s + v;
        if (fabsl(s) >= fabsl(v)) c += (s - tt) + v; else c += (v - tt) + s;
        s = tt; c += jobs[t].c; cnt += jobs[t].cnt;
    }
    free(th); free(jobs);
    if (count) *count = cnt;
    return (double)(s + c);
}

ORC_API double orc_scan_mt(const void* base, uint64_t n, int aos, int pred, double lo, double hi, int threads, uint64_t* count) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (mt_job){base, n * (uint64_t)t / (uint64_t)threads, n * (uint64_t)(t + 1) / (uint64_t)threads, aos, pred, lo, hi, 0.0, 0};
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    double s = 0.0; uint64_t c = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); s += jobs[t].sum; c += jobs[t].cnt; }
    free(th); free(jobs);
    if (count) *count = c;
    return s;
}
