// aqe_sql_kernels.cuh -- sm_100a kernels of the SQL-string path.
//
//   k_col_stats   min / max of a column as order-preserving 64-bit keys (+ "ids are first_id + row" check)
//   k_sql_ring    the grouped scan: SELECT agg(col) FROM t [WHERE conjunction] [GROUP BY g] with the reference's
//                 `rowid % step = 0` sampling, one launch, columns staged through a TMA ring
//   k_sql_agg     the same query over a strided sample or unaligned columns (register-staged)
//
// What this replaces: executor.cpp:28-338 issues one SQLite statement per group (plus a SELECT DISTINCT to
// find the groups), each a full table scan on the CPU; here every row is read once from HBM and lands in the
// accumulator of its group.
//
// Accumulation is 128-bit FIXED POINT: x -> round(x * 2^shift) (shift chosen from the column's max |x| so that
// the value fits 63 bits; integer columns use the value itself), summed as integers.  Integer addition
// commutes, so per-thread / per-CTA / per-GPU partials can be combined in any order -- including with atomics
// -- and the result is bit-reproducible; for data whose magnitudes span less than 2^10 (U(1,1000) amounts) the
// fixed-point sum is the EXACT sum.  Three bin layouts, picked by the group count G:
//   G == 1        registers, warp shuffles, one global update per CTA
//   G <= 16       thread-private bins in shared memory ([bin][thread], conflict-free), no atomics in the loop
//   G <= 4096     CTA-shared bins, 32-bit shared atomics (64-bit shared atomic adds are CAS loops on sm_100: SASS
//                 ATOMS.CAST.SPIN.64).  Packed form (MODE 3 / 4, whenever the fixed-point values of the aggregate column span
//                 fewer than 2^62 steps): THREE atomics per row carry the row count and the 62-bit value; general form
//                 (MODE 2: COUNT-only queries, full-range int64 columns): a count word + four limbs with carry chains
#pragma once

#include "aqe_kernels.cuh"

namespace aqe {

// ---- order-preserving keys -------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long okey_i64(long long v) { return (unsigned long long)v ^ 0x8000000000000000ull; }
__host__ __device__ __forceinline__ long long okey_to_i64(unsigned long long k) { return (long long)(k ^ 0x8000000000000000ull); }
__host__ __device__ __forceinline__ unsigned long long okey_bits_f64(unsigned long long b) {
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ unsigned long long okey_to_bits_f64(unsigned long long k) {
    return (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
}

struct ColStatsArgs {
    const void* col;
    int kind;              // 0 f64, 1 i64, 2 i32
    uint64_t n;
    long long first_id;    // dense check: col[i] == first_id + i  (kind 1 only, when check_dense)
    int check_dense;
    unsigned long long* out;  // [0] min key, [1] max key, [2] not-dense flag; initialised by the host
};

__global__ void __launch_bounds__(256) k_col_stats(const ColStatsArgs a) {
    unsigned long long mn = ~0ull, mx = 0ull;
    unsigned int bad = 0;
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += G) {
        unsigned long long k;
        if (a.kind == 2) k = okey_i64((long long)__ldg(static_cast<const int32_t*>(a.col) + i));
        else {
            const unsigned long long raw = (unsigned long long)__ldg(static_cast<const long long*>(a.col) + i);
            k = a.kind == 0 ? okey_bits_f64(raw) : okey_i64((long long)raw);
            if (a.check_dense && (long long)raw != a.first_id + (long long)i) bad = 1;
        }
        mn = k < mn ? k : mn; mx = k > mx ? k : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o1 = __shfl_down_sync(0xffffffffu, mn, d), o2 = __shfl_down_sync(0xffffffffu, mx, d);
        mn = o1 < mn ? o1 : mn; mx = o2 > mx ? o2 : mx;
        bad |= __shfl_down_sync(0xffffffffu, bad, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mn <= mx) { atomicMin(a.out + 0, mn); atomicMax(a.out + 1, mx); }
        if (bad) atomicOr(a.out + 2, 1ull);
    }
}

// ---- the grouped scan --------------------------------------------------------------------------------------
constexpr int kSqlThreads = 256;
constexpr int kSqlMaxCols = 5;          // the table has five columns; each is loaded at most once per row
constexpr int kSqlPrivateMaxGroups = 16;
constexpr unsigned int kSqlPackedRows = 65535;   // rows a thread may add to one private bin between drains (16-bit packed counter)
constexpr unsigned int kSqlPackedRowsMoments = 4095;   // ... with squares: three words per bin, 12 bits of headroom per field

constexpr unsigned int kSqlSharedPackedLimit = 1024;   // MODE 3: a bin whose packed row counter reached this is emptied by the thread that saw it

constexpr int kSqlMaxAlt = AQE_SQL_MAX_ALT;   // OR-ed conjunctions of a WHERE clause

struct SqlPred {       // one conjunct on one column: closed interval [lo, hi] and optionally != ne (raw 64-bit: f64 bits or int64)
    int has_pred;      // 0 none | 1 interval | 2 membership bitmap of an integer column: bit (v - lo) of `hi`, v - lo < 64 (IN lists and
                       //   other multi-branch clauses on a column whose values span fewer than 64 keys, folded on the host into ONE pass)
                       // | 3 the same for up to AQE_SQL_MAX_GROUPS keys: bit (v - lo) of SqlArgs::member_bits, v - lo < hi
    int has_ne;
    long long lo, hi, ne;
};
struct SqlCol {
    const void* ptr;
    int kind;          // 0 f64, 1 i64, 2 i32
    int mod_step;      // > 0: also requires value % mod_step == 0  (rowid sampling when ids are not dense)
    SqlPred pred[kSqlMaxAlt];
};

// Cross-GPU exchange fused into the grouped scan (one process per GPU, peers mapped with CUDA IPC over NVLink; same mailbox
// allocation as the scan exchange of aqe_kernels.cuh, the SQL area sits behind its slots).  The LAST CTA of rank r stores its
// shard's accumulators into slot [r][seq & 1] of EVERY rank's mailbox (peer stores), raises that slot's sequence flag, waits
// until its own mailbox holds `seq` from all ranks and adds the shards' words with 128-bit carries -- integer adds commute, so
// every rank ends with the identical table-level accumulators and no host merge, NCCL launch or staging copy is left.
constexpr size_t kSqlxSlotWords = (size_t)AQE_SQL_MAX_GROUPS * 5;
constexpr size_t kSqlxFlagsOffset = sizeof(ExSlot) * kMaxRanks * 4;                       // behind the scan / approx slots
constexpr size_t kSqlxDataOffset = kSqlxFlagsOffset + sizeof(unsigned long long) * kMaxRanks * 2;
constexpr size_t kMailboxBytes = kSqlxDataOffset + sizeof(unsigned long long) * kSqlxSlotWords * kMaxRanks * 2;
struct SqlExchange {
    int world;                      // 0/1 = disabled
    int rank;
    unsigned long long seq;
    unsigned long long timeout_cycles;
    unsigned char* peers[kMaxRanks];  // peers[r] = base of rank r's mailbox allocation as mapped in this process
    unsigned int* status;             // set to 1 if a peer did not show up in time
    unsigned long long* local;        // [n_groups][5] scratch in this GPU's memory
};
__device__ __forceinline__ unsigned long long* sqlx_flag(unsigned char* base, int sender, int par) {
    return reinterpret_cast<unsigned long long*>(base + kSqlxFlagsOffset) + sender * 2 + par;
}
__device__ __forceinline__ unsigned long long* sqlx_data(unsigned char* base, int sender, int par) {
    return reinterpret_cast<unsigned long long*>(base + kSqlxDataOffset) + ((size_t)sender * 2 + par) * kSqlxSlotWords;
}

struct SqlArgs {
    SqlCol cols[kSqlMaxCols];
    int ncols;
    int n_alt;         // 0: no WHERE; else a row passes when, for some alt < n_alt, every column's pred[alt] holds
    int agg_slot;      // index into cols, -1: count only
    int group_slot;    // index into cols, -1: no GROUP BY
    int agg_kind;
    long long key_min;
    unsigned int n_groups;
    double sum_scale, sq_scale;   // 2^sum_shift, 2^sq_shift
    // rows visited: i = first + j * stride for j in [0, count)
    uint64_t first, stride, count;
    unsigned long long* global_acc;   // [n_groups][5] {count, sum_lo, sum_hi, sq_lo, sq_hi}, zero before launch, zeroed again by the last CTA
    unsigned long long* out;          // [n_groups][5] device-visible result
    unsigned int* ticket;
    unsigned int member_bits[AQE_SQL_MAX_GROUPS / 32];   // has_pred == 3 (at most one predicate of a query uses it)
    int member_used;
    unsigned int drain_rows;          // private bins (G <= 16) are drained before a thread has added this many rows to one (<= kSqlPackedRows)
    int pair_bins;                    // private bins: words 0 and 1 of a bin sit side by side and are updated with ONE 128-bit load / store
    long long fx_bias;                // MODE 3 / 4: min(0, smallest fixed-point value of the aggregate column); bins hold sums of u = fx - fx_bias >= 0
    SqlExchange ex;
};

__device__ __forceinline__ long long sql_load_raw(const SqlCol& c, uint64_t i) {
    if (c.kind == 2) {
        int v;
        asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(static_cast<const int32_t*>(c.ptr) + i));
        return (long long)v;
    }
    long long v;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(static_cast<const long long*>(c.ptr) + i));
    return v;
}
// membership bitmap of a folded predicate (has_pred == 3), copied from the kernel parameters by sql_member_load
__shared__ unsigned int g_sql_member[AQE_SQL_MAX_GROUPS / 32];
__device__ __forceinline__ void sql_member_load(const SqlArgs& a, int tid, int nthreads) {   // the caller's next __syncthreads() publishes it
    if (a.member_used)
        for (int i = tid; i < AQE_SQL_MAX_GROUPS / 32; i += nthreads) g_sql_member[i] = a.member_bits[i];
}
__device__ __forceinline__ bool sql_member(unsigned long long d, unsigned long long n) {
    return d < n && ((g_sql_member[(unsigned int)d >> 5] >> ((unsigned int)d & 31u)) & 1u) != 0u;
}

__device__ __forceinline__ bool sql_pass(const SqlCol& c, const SqlPred& p, long long raw) {
    if (!p.has_pred) return true;
    if (p.has_pred == 2) {
        const unsigned long long d = (unsigned long long)(raw - p.lo);
        return d < 64ull && (((unsigned long long)p.hi >> d) & 1ull) != 0ull;
    }
    if (p.has_pred == 3) return sql_member((unsigned long long)(raw - p.lo), (unsigned long long)p.hi);
    bool ok;
    if (c.kind == 0) {
        const double d = __longlong_as_double(raw);
        ok = d >= __longlong_as_double(p.lo) && d <= __longlong_as_double(p.hi);
        if (p.has_ne) ok = ok && d != __longlong_as_double(p.ne);
    } else {
        ok = raw >= p.lo && raw <= p.hi;
        if (p.has_ne) ok = ok && raw != p.ne;
    }
    return ok;
}

// 128-bit add into two 64-bit words of GLOBAL memory with atomics (order-independent)
__device__ __forceinline__ void global_add128(unsigned long long* lo_word, unsigned long long lo, unsigned long long hi) {
    if (lo == 0 && hi == 0) return;
    unsigned long long carry = 0;
    if (lo) { const unsigned long long old = atomicAdd(lo_word, lo); carry = (old + lo < old) ? 1ull : 0ull; }
    if (hi + carry) atomicAdd(lo_word + 1, hi + carry);
}
// (lo32-sum, hi32-sum) split accumulators -> true 128-bit two's complement:  value = hi * 2^32 + lo
__device__ __forceinline__ void split_to_128(unsigned long long slo, long long shi, unsigned long long& lo, unsigned long long& hi) {
    lo = slo + ((unsigned long long)shi << 32);
    hi = (unsigned long long)((shi >> 32) + (lo < slo ? 1 : 0));
}

// 32-bit shared-memory limb add with carry out
__device__ __forceinline__ unsigned int limb_add(unsigned int* limb, unsigned int x, unsigned int carry_in) {
    const unsigned int s = x + carry_in;
    unsigned int carry = (s < x) ? 1u : 0u;  // x = 0xffffffff and carry_in = 1
    if (s) { const unsigned int old = atomicAdd(limb, s); carry |= (old + s < old) ? 1u : 0u; }
    return carry;
}
// The four limbs of bin g live G words apart (limb-major: word l * G + g), so that the 32 lanes of a warp, each with its own
// group key, spread over all 32 banks.  (Group-major [g][4] put limb l of every bin into the 8 banks l, l + 4, ...: the lanes of a
// warp queued ~7 deep on them instead of ~3.5, and the ATOMS wavefronts are what bounds this kernel.)
__device__ __forceinline__ void shared_add128(unsigned int* limbs, unsigned int G, unsigned int g, long long v) {
    const unsigned int sign = v < 0 ? 0xffffffffu : 0u;
    unsigned int c = limb_add(limbs + g, (unsigned int)v, 0u);
    c = limb_add(limbs + G + g, (unsigned int)((unsigned long long)v >> 32), c);
    c = limb_add(limbs + 2 * G + g, sign, c);
    limb_add(limbs + 3 * G + g, sign, c);
}

// MODE 3 of SqlBins (packed CTA-shared bins): the contents of one bin, taken out of shared memory by the caller -> global accumulators
template <bool MOMENTS>
__device__ __forceinline__ void sql_packed_to_global(unsigned long long* ga, long long bias, const unsigned int (&x)[4], const unsigned int (&y)[4]) {
    // (no shortcut on "no rows": a bin emptied by two threads at once can hand the second one words of rows whose counter the
    // first one took, or the other way round; every word is added for what it holds)
    if ((x[0] | x[1] | x[2] | x[3] | y[0] | y[1] | y[2] | y[3]) == 0u) return;
    const unsigned long long n = x[2] >> 20;
    const unsigned __int128 u = (unsigned __int128)x[0] + ((unsigned __int128)x[1] << 32) + ((unsigned __int128)(x[2] & 0xfffffu) << 54) + ((unsigned __int128)x[3] << 64);
    const __int128 v = (__int128)u + (__int128)(long long)n * (__int128)bias;
    if (n) atomicAdd(ga + 0, n);
    global_add128(ga + 1, (unsigned long long)v, (unsigned long long)((unsigned __int128)v >> 64));
    if constexpr (MOMENTS) {
        const unsigned __int128 q = (unsigned __int128)y[0] + ((unsigned __int128)y[1] << 32) + ((unsigned __int128)y[2] << 54) + ((unsigned __int128)y[3] << 64);
        global_add128(ga + 3, (unsigned long long)q, (unsigned long long)(q >> 64));
    }
}
// A bin whose row counter reached kSqlSharedPackedLimit, emptied by the thread that saw it (rare; out of line and with scalar arguments
// only, so that the bins' bookkeeping stays in registers in the callers' loops).  See the layout note in SqlBins.
template <bool MOMENTS>
__device__ __noinline__ void sql_packed_spill(unsigned int* s_sum, unsigned int* s_sq, unsigned int G, unsigned int g, long long bias, unsigned long long* ga) {
    unsigned int x[4], y[4] = {0u, 0u, 0u, 0u};
    x[2] = atomicExch(s_sum + 2 * G + g, 0u);
    x[0] = atomicExch(s_sum + g, 0u); x[1] = atomicExch(s_sum + G + g, 0u); x[3] = atomicExch(s_sum + 3 * G + g, 0u);
    if constexpr (MOMENTS) {
#pragma unroll
        for (int l = 0; l < 4; ++l) y[l] = atomicExch(s_sq + l * G + g, 0u);
    }
    sql_packed_to_global<MOMENTS>(ga, bias, x, y);
}

// one 62-bit value into words 0, 1 (and 3) of a packed bin
__device__ __forceinline__ void sql_packed_add_value(unsigned int* w, unsigned int G, unsigned int g, unsigned long long u) {
    const unsigned int x0 = (unsigned int)u;
    const unsigned int o0 = atomicAdd(w + g, x0);
    const unsigned int a1 = ((unsigned int)(u >> 32) & 0x3fffffu) + ((o0 + x0 < o0) ? 1u : 0u);
    if (a1) {   // (always, for floating-point columns; integer columns below 2^32 skip it)
        const unsigned int o1 = atomicAdd(w + G + g, a1);
        if (o1 + a1 < o1) atomicAdd(w + 3 * G + g, 1u);
    }
}
// one row into packed bin g: three shared atomics (six with squares), see the layout note in SqlBins
template <bool MOMENTS>
__device__ __forceinline__ void sql_packed_add_row(unsigned int* s_sum, unsigned int* s_sq, unsigned int G, unsigned int g, long long bias,
                                                   unsigned long long* spill_acc, long long fx, long long fq) {
    const unsigned long long u = (unsigned long long)(fx - bias);
    const unsigned int o2 = atomicAdd(s_sum + 2 * G + g, (unsigned int)(u >> 54) + (1u << 20));
    sql_packed_add_value(s_sum, G, g, u);
    if constexpr (MOMENTS) {
        const unsigned long long q = (unsigned long long)fq;
        sql_packed_add_value(s_sq, G, g, q);
        const unsigned int q2 = (unsigned int)(q >> 54);
        if (q2) atomicAdd(s_sq + 2 * G + g, q2);
    }
    if (o2 >= (kSqlSharedPackedLimit << 20)) sql_packed_spill<MOMENTS>(s_sum, s_sq, G, g, bias, spill_acc + (size_t)g * 5);
}

// ---- bins: where a row's (count, value, value^2) lands ------------------------------------------------------
// MODE 0: no GROUP BY (registers) | 1: thread-private shared bins | 2: CTA-shared bins with atomics, general form |
// 3: CTA-shared bins with atomics, packed form | 4: the same + the sparse walk of the ring kernel (queries with a WHERE clause).
// T = threads that add rows (the private bins are laid out [bin][T]); every thread of the CTA must call flush().
template <int MODE, bool MOMENTS, int T> struct SqlBins {
    unsigned int G;
    // MODE 1: w0[G][T] u64 = rows << 48 | sum of the values' low 32 bits   (both below their field width while a thread adds
    //                        fewer than kSqlPackedRows rows per bin between drains: one read-modify-write chain instead of two)
    //         w1[G][T] u64 = sum of (value >> 32) + 2^31 (biased non-negative)
    //         with squares, three words instead of four (16 bytes of shared-memory traffic less per row), u = value + 2^63, q = square < 2^62:
    //         w0 = rows << 52 | sum of u[0:40)      w1 = sum of q[0:16) << 36 | sum of u[40:64)      w2 = sum of q[16:62)
    //         -- every field has 12 spare bits: fewer than kSqlPackedRowsMoments rows per bin between drains
    // MODE 2: cnt[G] u32 | sum limbs [4][G] u32 | (sq limbs [4][G])
    // MODE 3: sum words [4][G] u32 | (sq words [4][G]), u = fx - bias in [0, 2^62), q = square in [0, 2^62):
    //         w0 += u[0:32)             the return value tells the carry c0 out of the word
    //         w1 += u[32:54) + c0       22-bit pieces: the word wraps at most once in 1023 rows; the return value tells, w3 counts the wraps
    //         w2 += u[54:62) + 2^20     rows in the top 12 bits, the sum of the pieces below them: both stay inside their fields for 4095 rows
    //         The ATOMS issue rate bounds these kernels, so what counts is atomics per row: 3 here (count included) against
    //         1 + 3.15 in MODE 2 (count word, two limbs, and a third limb for the 15 % of the rows whose second limb carries -- some
    //         lane of nearly every warp).  A thread that reads >= kSqlSharedPackedLimit rows out of w2's return value empties that bin
    //         into the global accumulators on the spot, word by word with atomicExch: every word is a plain sum and every wrap is
    //         counted by the add that caused it, so emptying a word between two adds loses nothing.  At most one add per thread is in
    //         flight between seeing the limit and emptying, so a counter stays below 1024 + 288 whatever the key distribution.
    unsigned int* b_cnt;
    unsigned long long *p_slo, *p_shi, *p_qlo;
    unsigned int *s_sum, *s_sq;
    unsigned long long r_cnt, r_slo, r_qlo;
    long long r_shi, r_qhi;
    bool with_sums;   // the query aggregates a column (false: COUNT only); uniform over the launch, set by the kernel after init()
    bool paired;      // MODE 1: words 0 and 1 interleaved as [bin][thread] 16-byte pairs (one LDS.128 + one STS.128 per row instead of
                      // two 64-bit chains); set by the kernel after init(), uniform over the launch.  p_slo then addresses the pairs.
    long long bias;                 // MODE 3 / 4 (SqlArgs::fx_bias), set by the kernel after init()
    unsigned long long* spill_acc;  // MODE 3 / 4: the global accumulators a full bin is emptied into, set by the kernel after init()

    static size_t smem_bytes(unsigned int G) {
        if (MODE == 1) return (size_t)(G + 1) * T * (MOMENTS ? 24 : 16);   // + one bin per thread that is never drained: where the ring kernel sends keys outside the layout
        if (MODE == 2) return (size_t)G * (4 + 16 + (MOMENTS ? 16 : 0));
        if (MODE >= 3) return (size_t)G * (16 + (MOMENTS ? 16 : 0));
        return 0;
    }
    // all threads of the CTA call init (it contains a barrier); tid < T owns a private column of bins
    __device__ __forceinline__ void init(unsigned char* smem, unsigned int groups, int tid, int nthreads) {
        G = groups;
        with_sums = true; paired = false; bias = 0; spill_acc = nullptr;
        b_cnt = reinterpret_cast<unsigned int*>(smem);
        r_cnt = 0; r_slo = 0; r_qlo = 0; r_shi = 0; r_qhi = 0;
        if constexpr (MODE == 1) {
            p_slo = reinterpret_cast<unsigned long long*>(smem);
            p_shi = p_slo + (size_t)(G + 1) * T;
            p_qlo = p_shi + (size_t)(G + 1) * T;
            // every word of the bins, whichever way the kernel addresses them afterwards (`paired` is set after init)
            for (unsigned int i = tid; i < (G + 1) * T * (MOMENTS ? 3u : 2u); i += nthreads) p_slo[i] = 0;
        } else if constexpr (MODE == 2) {
            s_sum = b_cnt + G;
            s_sq = s_sum + (size_t)G * 4;
            for (unsigned int i = tid; i < G; i += nthreads) b_cnt[i] = 0;
            for (unsigned int i = tid; i < G * 4; i += nthreads) { s_sum[i] = 0; if constexpr (MOMENTS) s_sq[i] = 0; }
        } else if constexpr (MODE >= 3) {
            s_sum = b_cnt;
            s_sq = s_sum + (size_t)G * 4;
            for (unsigned int i = tid; i < G * 4; i += nthreads) { s_sum[i] = 0; if constexpr (MOMENTS) s_sq[i] = 0; }
        }
        __syncthreads();
    }
    __device__ __forceinline__ void add(unsigned int g, int tid, bool has_sum, long long fx, long long fq) {
        if constexpr (MODE == 0) {
            r_cnt += 1;
            r_slo += (unsigned long long)fx & 0xffffffffull; r_shi += fx >> 32;
            if constexpr (MOMENTS) { r_qlo += (unsigned long long)fq & 0xffffffffull; r_qhi += fq >> 32; }
        } else if constexpr (MODE == 1) {
            const unsigned int s = g * T + tid;
            if constexpr (MOMENTS) {
                const unsigned long long u = (unsigned long long)fx ^ (1ull << 63), q = (unsigned long long)fq;
                if (paired) {
                    ulonglong2* pp = reinterpret_cast<ulonglong2*>(p_slo) + s;
                    ulonglong2 v = *pp;
                    v.x += (1ull << 52) | (u & ((1ull << 40) - 1));
                    v.y += (u >> 40) | ((q & 0xffffull) << 36);
                    *pp = v;
                } else {
                    p_slo[s] += (1ull << 52) | (u & ((1ull << 40) - 1));
                    p_shi[s] += (u >> 40) | ((q & 0xffffull) << 36);
                }
                p_qlo[s] += q >> 16;
            } else if (paired) {   // (implies has_sum)
                ulonglong2* pp = reinterpret_cast<ulonglong2*>(p_slo) + s;
                ulonglong2 v = *pp;
                v.x += (1ull << 48) + ((unsigned long long)fx & 0xffffffffull);
                v.y += (unsigned long long)((fx >> 32) + 0x80000000ll);
                *pp = v;
            } else {
                p_slo[s] += (1ull << 48) + ((unsigned long long)fx & 0xffffffffull);
                if (has_sum) p_shi[s] += (unsigned long long)((fx >> 32) + 0x80000000ll);   // COUNT-only queries keep one chain
            }
        } else if constexpr (MODE == 2) {
            atomicAdd(b_cnt + g, 1u);
            if (has_sum) {
                shared_add128(s_sum, G, g, fx);
                if constexpr (MOMENTS) shared_add128(s_sq, G, g, fq);
            }
        } else {
            sql_packed_add_row<MOMENTS>(s_sum, s_sq, G, g, bias, spill_acc, fx, fq);
        }
    }
    __device__ __forceinline__ void zero_private(int tid) {
        for (unsigned int g = 0; g < G; ++g) {
            // a thread clears ITS OWN bins (after a drain the others are already adding to theirs)
            if (paired) reinterpret_cast<ulonglong2*>(p_slo)[g * T + tid] = make_ulonglong2(0ull, 0ull);
            else { p_slo[g * T + tid] = 0; p_shi[g * T + tid] = 0; }
            if constexpr (MOMENTS) p_qlo[g * T + tid] = 0;
        }
    }
    // MODE 1: the T bin-owning threads fold their private bins into the global accumulators and start over.  Called by exactly
    // those T threads (a named barrier keeps the producer warp of the ring kernel out of it), at the end of the kernel and -- so
    // that the packed row counters cannot overflow -- every time a thread may have added kSqlPackedRows rows to one bin.
    __device__ __forceinline__ void drain(unsigned long long* global_acc, int tid) {
        asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
        const int warp = tid >> 5, lane = tid & 31;
        for (unsigned int g = warp; g < G; g += T / 32) {
            unsigned long long c = 0, sl = 0, ql = 0;
            long long sh = 0, qh = 0;
#pragma unroll
            for (int k = 0; k < T / 32; ++k) {
                const unsigned int s = g * T + lane + 32 * k;
                if constexpr (MOMENTS) {   // value = lo + (hi << 32) for both sums, as split_to_128 takes them
                    const unsigned long long w0 = paired ? p_slo[2 * s] : p_slo[s], w1 = paired ? p_slo[2 * s + 1] : p_shi[s], w2 = p_qlo[s], n = w0 >> 52;
                    const unsigned long long a = w0 & ((1ull << 52) - 1), b = w1 & ((1ull << 36) - 1), cq = w1 >> 36;
                    c += n;
                    sl += a & 0xffffffffull;
                    sh += (long long)((a >> 32) + (b << 8)) - (long long)(n << 31);   // sum of u = a + (b << 40); the n biases of 2^63 leave as n << 31 here
                    ql += cq + ((w2 & 0xffffull) << 16);                              // sum of q = cq + (w2 << 16)
                    qh += (long long)(w2 >> 16);
                } else {
                    const unsigned long long w0 = paired ? p_slo[2 * s] : p_slo[s], n = w0 >> 48;
                    c += n; sl += w0 & 0xffffffffffffull;
                    if (with_sums) sh += (long long)(paired ? p_slo[2 * s + 1] : p_shi[s]) - (long long)(n << 31);   // remove the bias of the n rows
                }
            }
            c = warp_reduce_u64(c); sl = warp_reduce_u64(sl); sh = (long long)warp_reduce_u64((unsigned long long)sh);
            if constexpr (MOMENTS) { ql = warp_reduce_u64(ql); qh = (long long)warp_reduce_u64((unsigned long long)qh); }
            if (lane == 0 && c) {
                unsigned long long lo, hi;
                unsigned long long* ga = global_acc + (size_t)g * 5;
                atomicAdd(ga + 0, c);
                split_to_128(sl, sh, lo, hi); global_add128(ga + 1, lo, hi);
                if constexpr (MOMENTS) { split_to_128(ql, qh, lo, hi); global_add128(ga + 3, lo, hi); }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
        zero_private(tid);
    }
    // CTA totals -> global accumulators (integer atomics: order does not matter)
    __device__ __forceinline__ void flush(unsigned long long* global_acc, int tid, int nthreads) {
        if constexpr (MODE == 0) {
            __shared__ unsigned long long part[32][5];
            r_cnt = warp_reduce_u64(r_cnt);
            r_slo = warp_reduce_u64(r_slo); r_shi = (long long)warp_reduce_u64((unsigned long long)r_shi);
            if constexpr (MOMENTS) { r_qlo = warp_reduce_u64(r_qlo); r_qhi = (long long)warp_reduce_u64((unsigned long long)r_qhi); }
            const int warp = tid >> 5, lane = tid & 31, nw = (nthreads + 31) >> 5;
            if (lane == 0) { part[warp][0] = r_cnt; part[warp][1] = r_slo; part[warp][2] = (unsigned long long)r_shi; part[warp][3] = r_qlo; part[warp][4] = (unsigned long long)r_qhi; }
            __syncthreads();
            if (warp == 0) {
                unsigned long long v[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) v[i] = warp_reduce_u64(lane < nw ? part[lane][i] : 0ull);
                if (lane == 0 && v[0]) {
                    unsigned long long lo, hi;
                    atomicAdd(global_acc + 0, v[0]);
                    split_to_128(v[1], (long long)v[2], lo, hi); global_add128(global_acc + 1, lo, hi);
                    if constexpr (MOMENTS) { split_to_128(v[3], (long long)v[4], lo, hi); global_add128(global_acc + 3, lo, hi); }
                }
            }
        } else if constexpr (MODE == 1) {
            if (tid < T) drain(global_acc, tid);
        } else if constexpr (MODE >= 3) {
            __syncthreads();
            for (unsigned int g = tid; g < G; g += nthreads) {
                unsigned int x[4], y[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int l = 0; l < 4; ++l) { x[l] = s_sum[l * G + g]; if constexpr (MOMENTS) y[l] = s_sq[l * G + g]; }
                sql_packed_to_global<MOMENTS>(global_acc + (size_t)g * 5, bias, x, y);
            }
        } else {
            __syncthreads();
            for (unsigned int g = tid; g < G; g += nthreads) {
                const unsigned int c = b_cnt[g];
                if (!c) continue;
                unsigned long long* ga = global_acc + (size_t)g * 5;
                atomicAdd(ga + 0, (unsigned long long)c);
                const unsigned int* l = s_sum + g;
                global_add128(ga + 1, ((unsigned long long)l[G] << 32) | l[0], ((unsigned long long)l[3 * G] << 32) | l[2 * G]);
                if constexpr (MOMENTS) {
                    const unsigned int* m = s_sq + g;
                    global_add128(ga + 3, ((unsigned long long)m[G] << 32) | m[0], ((unsigned long long)m[3 * G] << 32) | m[2 * G]);
                }
            }
        }
    }
};

// One row, row-at-a-time form (register-staged kernel and ragged tails): predicate, fixed-point conversion, bin update.
template <int MODE, bool MOMENTS, int T>
__device__ __forceinline__ void sql_consume(const SqlArgs& a, SqlBins<MODE, MOMENTS, T>& bins, int tid, const long long (&raw)[kSqlMaxCols]) {
    bool pass = a.n_alt == 0;
    for (int alt = 0; alt < a.n_alt; ++alt) {
        bool all = true;
#pragma unroll
        for (int c = 0; c < kSqlMaxCols; ++c)
            if (c < a.ncols) all = all && sql_pass(a.cols[c], a.cols[c].pred[alt], raw[c]);
        pass = pass || all;
    }
#pragma unroll
    for (int c = 0; c < kSqlMaxCols; ++c)
        if (c < a.ncols && a.cols[c].mod_step > 0) pass = pass && (raw[c] % (long long)a.cols[c].mod_step == 0);
    if (!pass) return;
    long long fx = 0, fq = 0;
    if (a.agg_slot >= 0) {
        long long rv = 0;
#pragma unroll
        for (int c = 0; c < kSqlMaxCols; ++c) if (c == a.agg_slot) rv = raw[c];
        double d;
        if (a.agg_kind == 0) { d = __longlong_as_double(rv); fx = __double2ll_rn(__dmul_rn(d, a.sum_scale)); }
        else { d = (double)rv; fx = rv; }
        if constexpr (MOMENTS) fq = __double2ll_rn(__dmul_rn(__dmul_rn(d, d), a.sq_scale));
    }
    unsigned int g = 0;
    if constexpr (MODE != 0) {
        long long kv = 0;
#pragma unroll
        for (int c = 0; c < kSqlMaxCols; ++c) if (c == a.group_slot) kv = raw[c];
        g = (unsigned int)(kv - a.key_min);
        if (g >= bins.G) return;  // cannot happen when the layout came from this table's statistics
    }
    bins.add(g, tid, a.agg_slot >= 0, fx, fq);
}

// The exchange proper; called by every thread of ONE CTA with this shard's accumulators in ex.local.
__device__ __forceinline__ void sql_exchange(const SqlExchange& ex, unsigned int n_groups, unsigned long long* out, int tid, int nthreads) {
    const unsigned int words = n_groups * 5;
    const int par = (int)(ex.seq & 1ull);
    for (int r = 0; r < ex.world; ++r) {   // peer stores over NVLink, coalesced
        volatile unsigned long long* dst = sqlx_data(ex.peers[r], ex.rank, par);
        for (unsigned int i = tid; i < words; i += nthreads) dst[i] = __ldcg(ex.local + i);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < ex.world) st_release_sys(sqlx_flag(ex.peers[tid], ex.rank, par), ex.seq);
    if (tid < ex.world) {
        const unsigned long long* flag = sqlx_flag(ex.peers[ex.rank], tid, par);
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != ex.seq) {
            if ((unsigned long long)(clock64() - t0) > ex.timeout_cycles) { *(volatile unsigned int*)ex.status = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (unsigned int g = tid; g < n_groups; g += nthreads) {   // fold the shards in rank order (any order gives the same bits)
        unsigned long long cnt = 0, lo[2] = {0, 0}, hi[2] = {0, 0};
        for (int r = 0; r < ex.world; ++r) {
            const volatile unsigned long long* src = sqlx_data(ex.peers[ex.rank], r, par) + (size_t)g * 5;
            cnt += src[0];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const unsigned long long l = src[1 + 2 * k], h = src[2 + 2 * k];
                const unsigned long long nl = lo[k] + l;
                hi[k] += h + (nl < lo[k] ? 1ull : 0ull);
                lo[k] = nl;
            }
        }
        unsigned long long* o = out + (size_t)g * 5;
        o[0] = cnt; o[1] = lo[0]; o[2] = hi[0]; o[3] = lo[1]; o[4] = hi[1];
    }
}

// last CTA: publish the accumulators (through the cross-GPU exchange when one is connected) and re-arm them for the next launch
__device__ __forceinline__ void sql_publish(const SqlArgs& a, int tid, int nthreads) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // (replicated accumulators -- CTA b adding into copy b mod 8 to shorten the same-address atomic queues in L2 -- were
    // tried: no gain on the scan, and summing the copies here cost 0.1-0.2 ms at 1000 groups)
    const unsigned int words = a.n_groups * 5;
    unsigned long long* dst = a.ex.world > 1 ? a.ex.local : a.out;
    for (unsigned int i0 = tid; i0 < words; i0 += 4 * nthreads) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const unsigned int i = i0 + u * nthreads; v[u] = i < words ? __ldcg(a.global_acc + i) : 0ull; }  // 4 loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) { const unsigned int i = i0 + u * nthreads; if (i < words) { dst[i] = v[u]; a.global_acc[i] = 0ull; } }
    }
    if (tid == 0) *a.ticket = 0u;
    if (a.ex.world > 1) {
        __threadfence();
        __syncthreads();
        sql_exchange(a.ex, a.n_groups, a.out, tid, nthreads);
    }
}

// A shard with nothing to scan (no rows, an unsatisfiable WHERE, COUNT(*) answered from metadata) still takes part in the exchange:
// ex.local holds what the host computed.
__global__ void __launch_bounds__(256) k_sql_exchange_only(const SqlExchange ex, unsigned int n_groups, unsigned long long* out) {
    sql_exchange(ex, n_groups, out, threadIdx.x, blockDim.x);
}

// Register-staged visit: strided samples (rowid % step = 0 over dense ids becomes an arithmetic progression of row
// numbers) and columns that are not 16-byte aligned.  Same column-at-a-time shape as the ring consumer, the rows of a
// batch coming straight from global memory: a thread owns U rows per batch, keeps one pass bit per row, runs each
// predicate column as one pass of U independent loads, then the group column, then the aggregate column.
template <int U> __device__ __forceinline__ uint32_t sql_pred_pass_global(const SqlCol& col, const SqlPred& p, const uint64_t (&idx)[U], uint32_t mask) {
#pragma unroll
    for (int e = 0; e < U; ++e) {
        const long long raw = sql_load_raw(col, idx[e]);
        mask &= ~((sql_pass(col, p, raw) ? 0u : 1u) << e);
    }
    return mask;
}

template <int MODE, bool MOMENTS>
__global__ void __launch_bounds__(kSqlThreads) k_sql_agg(const SqlArgs a) {
    extern __shared__ __align__(16) unsigned char sql_smem[];
    const int tid = threadIdx.x;
    constexpr int T = kSqlThreads;
    constexpr int U = 8;
    SqlBins<MODE, MOMENTS, T> bins;
    sql_member_load(a, tid, T);
    bins.init(sql_smem, a.n_groups, tid, T);
    bins.with_sums = a.agg_slot >= 0;
    bins.paired = MODE == 1 && a.pair_bins != 0 && a.agg_slot >= 0;
    bins.bias = a.fx_bias; bins.spill_acc = a.global_acc;

    // per-query facts, read from the parameter bank once
    const int agg_slot = a.agg_slot, group_slot = a.group_slot, n_alt = a.n_alt;
    const int agg_kind = agg_slot >= 0 ? a.cols[agg_slot].kind : -1;
    uint32_t pred_cols[kSqlMaxAlt];
    int mod_slot = -1;
#pragma unroll
    for (int alt = 0; alt < kSqlMaxAlt; ++alt) {
        pred_cols[alt] = 0u;
        for (int k = 0; k < a.ncols; ++k)
            if (alt < n_alt && a.cols[k].pred[alt].has_pred) pred_cols[alt] |= 1u << k;
    }
    for (int k = 0; k < a.ncols; ++k) if (a.cols[k].mod_step > 0) mod_slot = k;

    const uint64_t gsz = (uint64_t)gridDim.x * T;
    const uint64_t last = a.count - 1;  // a.count > 0 (the host launches nothing otherwise)
    unsigned int rows_since_drain = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * T; base < a.count; base += U * gsz) {   // trip count uniform over the CTA (drain() has a barrier)
        const uint64_t j = base + tid;
        if constexpr (MODE == 1) {
            if (rows_since_drain + U > a.drain_rows) { bins.drain(a.global_acc, tid); rows_since_drain = 0; }
            rows_since_drain += U;
        }
        uint64_t idx[U];
        uint32_t mask = 0u;
#pragma unroll
        for (int e = 0; e < U; ++e) {
            const uint64_t je = j + (uint64_t)e * gsz;
            if (je < a.count) mask |= 1u << e;
            idx[e] = a.first + (je < a.count ? je : last) * a.stride;  // slots past the end re-read the last row and stay masked off
        }
        if (n_alt == 1) {
            for (uint32_t pc = pred_cols[0]; pc; pc &= pc - 1) {
                const int k = __ffs(pc) - 1;
                mask = sql_pred_pass_global<U>(a.cols[k], a.cols[k].pred[0], idx, mask);
            }
        } else if (n_alt > 1) {
            uint32_t any = 0u;
#pragma unroll
            for (int alt = 0; alt < kSqlMaxAlt; ++alt) {
                if (alt >= n_alt) break;
                uint32_t m = mask;
                for (uint32_t pc = pred_cols[alt]; pc; pc &= pc - 1) {
                    const int k = __ffs(pc) - 1;
                    m = sql_pred_pass_global<U>(a.cols[k], a.cols[k].pred[alt], idx, m);
                }
                any |= m;
            }
            mask = any;
        }
        if (mod_slot >= 0) {
#pragma unroll
            for (int e = 0; e < U; ++e)
                if ((mask >> e) & 1u) {
                    if (sql_load_raw(a.cols[mod_slot], idx[e]) % (long long)a.cols[mod_slot].mod_step != 0) mask &= ~(1u << e);
                }
        }
        unsigned int g[U];
#pragma unroll
        for (int e = 0; e < U; ++e) g[e] = 0;
        if constexpr (MODE != 0) {
#pragma unroll
            for (int e = 0; e < U; ++e) g[e] = (unsigned int)(sql_load_raw(a.cols[group_slot], idx[e]) - a.key_min);
#pragma unroll
            for (int e = 0; e < U; ++e) if (g[e] >= bins.G) mask &= ~(1u << e);  // cannot happen for live rows with this table's own layout
        }
        long long fx[U], fq[U];
#pragma unroll
        for (int e = 0; e < U; ++e) { fx[e] = 0; fq[e] = 0; }
        if (agg_kind >= 0) {
#pragma unroll
            for (int e = 0; e < U; ++e) {
                const long long rv = sql_load_raw(a.cols[agg_slot], idx[e]);
                double d;
                if (agg_kind == 0) { d = __longlong_as_double(rv); fx[e] = __double2ll_rn(__dmul_rn(d, a.sum_scale)); }
                else { d = (double)rv; fx[e] = rv; }
                if constexpr (MOMENTS) fq[e] = __double2ll_rn(__dmul_rn(__dmul_rn(d, d), a.sq_scale));
            }
        }
#pragma unroll
        for (int e = 0; e < U; ++e) if ((mask >> e) & 1u) bins.add(g[e], tid, agg_kind >= 0, fx[e], fq[e]);
    }
    bins.flush(a.global_acc, tid, T);
    sql_publish(a, tid, T);
}

// TMA-staged ring (the default for full scans of aligned columns): a producer warp streams, per tile of `tile_rows`
// rows, the matching slice of EVERY column the query reads into one shared-memory stage with 1-D bulk copies
// (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; 8 consumer warps work on the tile out of shared memory.
// Bytes in flight are set by the ring, not by registers.
//
// The consumers run COLUMN AT A TIME over a tile: thread t owns rows t, t+256, ... (K of them,
// conflict-free LDS.64 / LDS.32) and keeps one pass bit per row in a register.  Each predicate column is one tight,
// type-specialised, fully unrolled pass over those rows (the type switch sits outside the row loop), then one pass
// derives the group index, then one pass converts and accumulates the aggregate column.  A row-at-a-time interpreter
// of the same query costs ~110 warp instructions per 32 rows (measured, ncu) and is issue/latency bound at 1.0 TB/s.
// A tile holds exactly 256 K rows (K = 16 for a single int32 column, 8 for rows of <= 8 bytes, 8 / 6 / 4 for wider rows as the
// occupancy allows -- sql_launch_ring -- so a stage stays <= 32 KiB and every row slot of a full tile is live); the passes are
// unconditional over the K slots, and slots beyond the rows of the table's last tile read stale bytes of the same stage and
// stay masked off.  Every tile but the table's last is whole: its pass bits start as the constant (1 << K) - 1.
//
// The reference's `rowid % step = 0` over dense ids is a filter on the row number here (no id column read) -- used
// while step is small enough that every 32-byte sector is touched anyway; larger steps take the strided visit.
constexpr int kSqlMaxRowsPerThread = 16;  // K, rows per consumer thread and tile: 16 for 4-byte rows, 8 for rows of <= 8 bytes, 8 / 6 / 4 for wider rows (stage = 256 K rows)

struct SqlRingArgs {
    SqlArgs q;
    uint32_t tile_rows;           // 256 * K
    uint32_t col_off[kSqlMaxCols];  // byte offset of each column's slice inside a stage (multiples of 16)
    uint32_t stage_bytes;
    uint32_t ring_bytes;          // STAGES * stage_bytes
    uint32_t samp_step;           // 0/1: every row; else row i passes iff (i + samp_phase) % samp_step == 0
    uint32_t samp_phase;
};

template <typename T> __device__ __forceinline__ T lds_row(const unsigned char* base, uint32_t row) {
    return *reinterpret_cast<const T*>(base + (size_t)row * sizeof(T));
}

// f64 range test of one row into its pass bit: three chained setp and one predicated and (the `? 0u : 1u) << k` form of the
// same test compiles to a SEL per comparison: ~7.5 instructions per row against 4).  `ne` is NaN when the conjunct has no `!=`:
// setp.neu of anything with NaN holds.  A NaN value fails the first comparison, as in `v >= lo && v <= hi && !(v == ne)`.
__device__ __forceinline__ uint32_t sql_f64_range_bit(uint32_t mask, int bit, double v, double lo, double hi, double ne) {
    asm("{\n .reg .pred p;\n setp.ge.f64 p, %1, %2;\n setp.le.and.f64 p, %1, %3, p;\n setp.neu.and.f64 p, %1, %4, p;\n @!p and.b32 %0, %0, %5;\n}"
        : "+r"(mask) : "d"(v), "d"(lo), "d"(hi), "d"(ne), "r"(~(1u << bit)));   // (an immediate once the row loop is unrolled)
    return mask;
}

// the same for an int32 column (region, product_id): two (three with a `!=`) chained setp and one predicated and -- the 4-byte scans
// spend 16 warp instructions per row (ncu, profiles/r2_sql_count_ncu_full_raw.csv) and the SEL form of this test was 7 of them
__device__ __forceinline__ uint32_t sql_i32_range_bit(uint32_t mask, int bit, int v, int lo, int hi) {
    asm("{\n .reg .pred p;\n setp.ge.s32 p, %1, %2;\n setp.le.and.s32 p, %1, %3, p;\n @!p and.b32 %0, %0, %4;\n}"
        : "+r"(mask) : "r"(v), "r"(lo), "r"(hi), "r"(~(1u << bit)));
    return mask;
}
__device__ __forceinline__ uint32_t sql_i32_range_ne_bit(uint32_t mask, int bit, int v, int lo, int hi, int ne) {
    asm("{\n .reg .pred p;\n setp.ge.s32 p, %1, %2;\n setp.le.and.s32 p, %1, %3, p;\n setp.ne.and.s32 p, %1, %4, p;\n @!p and.b32 %0, %0, %5;\n}"
        : "+r"(mask) : "r"(v), "r"(lo), "r"(hi), "r"(ne), "r"(~(1u << bit)));
    return mask;
}
// one conjunct of one column over this thread's row slots of the tile: clears the pass bit of every row that fails
template <int T, int K> __device__ __forceinline__ uint32_t sql_pred_pass(const SqlCol& col, const SqlPred& p, const unsigned char* base, int tid, uint32_t mask) {
    if (col.kind == 0) {
        const double lo = __longlong_as_double(p.lo), hi = __longlong_as_double(p.hi);
        const double ne = p.has_ne != 0 ? __longlong_as_double(p.ne) : __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
        for (int k = 0; k < K; ++k) mask = sql_f64_range_bit(mask, k, lds_row<double>(base, tid + k * T), lo, hi, ne);
    } else if (p.has_pred == 3) {   // membership bitmap in shared memory over [lo, lo + hi)
        const long long first = p.lo;
        const unsigned long long n = (unsigned long long)p.hi;
        if (col.kind == 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) mask &= ~((sql_member((unsigned long long)(lds_row<long long>(base, tid + k * T) - first), n) ? 0u : 1u) << k);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) mask &= ~((sql_member((unsigned long long)((long long)lds_row<int>(base, tid + k * T) - first), n) ? 0u : 1u) << k);
        }
    } else if (p.has_pred == 2) {   // membership bitmap over [lo, lo + 64)
        const long long first = p.lo;
        const unsigned long long bits = (unsigned long long)p.hi;
        if (col.kind == 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned long long d = (unsigned long long)(lds_row<long long>(base, tid + k * T) - first);
                mask &= ~(((d < 64ull && ((bits >> d) & 1ull)) ? 0u : 1u) << k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned long long d = (unsigned long long)((long long)lds_row<int>(base, tid + k * T) - first);
                mask &= ~(((d < 64ull && ((bits >> d) & 1ull)) ? 0u : 1u) << k);
            }
        }
    } else if (col.kind == 1) {
        const bool has_ne = p.has_ne != 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long v = lds_row<long long>(base, tid + k * T);
            const bool ok = v >= p.lo && v <= p.hi && !(has_ne && v == p.ne);
            mask &= ~((ok ? 0u : 1u) << k);
        }
    } else {
        // int32 column: the int64 bounds clamp to the int32 range
        const long long lo64 = p.lo < -2147483648ll ? -2147483648ll : p.lo, hi64 = p.hi > 2147483647ll ? 2147483647ll : p.hi;
        if (lo64 > 2147483647ll || hi64 < -2147483648ll) return 0u;
        const int lo = (int)lo64, hi = (int)hi64;
        const bool has_ne = p.has_ne != 0 && p.ne >= -2147483648ll && p.ne <= 2147483647ll;
        const int ne = (int)p.ne;
        if (has_ne) {
#pragma unroll
            for (int k = 0; k < K; ++k) mask = sql_i32_range_ne_bit(mask, k, lds_row<int>(base, tid + k * T), lo, hi, ne);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) mask = sql_i32_range_bit(mask, k, lds_row<int>(base, tid + k * T), lo, hi);
        }
    }
    return mask;
}
// rowid % step = 0 on the id values themselves (ids with gaps)
template <int T, int K> __device__ __forceinline__ uint32_t sql_mod_pass(const SqlCol& col, const unsigned char* base, int tid, uint32_t mask) {
#pragma unroll
    for (int k = 0; k < K; ++k)
        if ((mask >> k) & 1u) {
            const long long v = lds_row<long long>(base, tid + k * T);
            if (v % (long long)col.mod_step != 0) mask &= ~(1u << k);
        }
    return mask;
}

// CTAs per SM the register allocation must leave room for (shared memory usually sets the real limit): 4 for the plain
// ungrouped and the shared-atomic kernels (<= 56 registers), 3 for private bins and for moments, 2 for private bins with moments.
// The packed shared bins always come with rows of >= 12 bytes (group column + aggregate column): shared memory holds three CTAs of them at K = 8 (with squares at K = 6).
constexpr int sql_ring_min_ctas(int mode, bool moments, int k = 8) {
    if (k == 16 && mode == 2) return 3;   // sixteen row slots per thread: under the 56 registers of four CTAs per SM they spill (0.83 -> 0.95 ms), under 72 they do not (0.78 -> 0.64)
    return mode == 1 ? (moments ? 2 : 3) : (moments || mode == 4 ? 3 : 4);
}

// Packed shared-atomic bins, tiles few of whose rows pass: the ATOMS of a predicated-off row still costs its issue slot, so when no
// thread of the warp kept more than half of its K rows the warp walks the set pass bits instead -- max over the lanes of popc(mask)
// rounds in place of K, each reading its row's group key and value out of the stage again (two LDS, far cheaper than the atomics
// saved).  This second code path costs registers (72 against 56): it exists in MODE 4 only, which the host picks for queries
// with a WHERE clause; queries without one run MODE 3, whose 56 registers leave room for a fourth CTA per SM.
struct SqlSparseConsts {
    unsigned int* s_sum; unsigned int* s_sq; unsigned long long* spill_acc;
    long long bias, key_min;
    double sum_scale, sq_scale;
    unsigned int G; int group_kind, agg_kind;
};
template <bool MOMENTS, int T>
__device__ __forceinline__ void sql_sparse_adds(const SqlSparseConsts& s, const unsigned char* gb, const unsigned char* ab, int tid, uint32_t mask, unsigned int rounds) {
    for (unsigned int i = 0; i < rounds; ++i) {
        if (mask) {
            const uint32_t row = (uint32_t)tid + (uint32_t)(__ffs(mask) - 1) * (uint32_t)T;
            mask &= mask - 1u;
            const unsigned int g = s.group_kind == 2 ? (unsigned int)(lds_row<int>(gb, row) - (int)s.key_min) : (unsigned int)(lds_row<long long>(gb, row) - s.key_min);
            long long fx, fq = 0;
            double d;
            if (s.agg_kind == 0) { d = lds_row<double>(ab, row); fx = __double2ll_rn(__dmul_rn(d, s.sum_scale)); }
            else { fx = s.agg_kind == 2 ? (long long)lds_row<int>(ab, row) : lds_row<long long>(ab, row); d = (double)fx; }
            if constexpr (MOMENTS) fq = __double2ll_rn(__dmul_rn(__dmul_rn(d, d), s.sq_scale));
            sql_packed_add_row<MOMENTS>(s.s_sum, s.s_sq, s.G, g, s.bias, s.spill_acc, fx, fq);
        }
    }
}

// MODE 1 with squares over paired bins (the reference's run_query_groupby_with_ci over a handful of groups), one row, no branch:
// the bin is read whether or not the row passes (the address is always a bin of this thread) and only the two stores carry the
// pass bit.  The kernel is bound by its instruction count: this form costs 19 instructions per row against 25 for
// `if (bit) bins.add(...)` (BSSY / BRA / BSYNC around every row, the runtime `paired` test, the row index).  Same words as
// SqlBins<1, true, T>::add: w0 += 1 << 52 | u[0:40), w1 += q[0:16) << 36 | u[40:64), w2 += q >> 16 with u = fx ^ 2^63.
// Loads and stores are volatile asm without a memory clobber, so that the compiler keeps them in order among themselves and
// stays free to batch the stage reads of the K rows ahead of them; the mbarrier wait at the top of every tile and the
// bar.sync at the start of a drain (both with memory clobbers) fence them against the C++ stores that empty the bins.
__device__ __forceinline__ void sql_private_moments_row(uint32_t pair_addr, uint32_t q_addr, long long fx, long long fq, uint32_t pass_bit) {
    asm volatile("{\n"
        " .reg .pred p;\n"
        " .reg .b64 w0, w1, w2, t, u;\n"
        " ld.shared.v2.b64 {w0, w1}, [%0];\n"
        " ld.shared.b64 w2, [%1];\n"
        " setp.ne.u32 p, %4, 0;\n"
        " and.b64 t, %2, 0x000000ffffffffff;\n"
        " add.u64 w0, w0, t;\n"
        " add.u64 w0, w0, 0x0010000000000000;\n"   // (a separate add: ptxas folds it into the IADD3.X of the high word)
        " xor.b64 u, %2, 0x8000000000000000;\n"
        " shr.u64 u, u, 40;\n"
        " and.b64 t, %3, 0xffff;\n"
        " shl.b64 t, t, 36;\n"
        " or.b64 t, t, u;\n"
        " add.u64 w1, w1, t;\n"
        " shr.u64 t, %3, 16;\n"
        " add.u64 w2, w2, t;\n"
        " @p st.shared.v2.b64 [%0], {w0, w1};\n"
        " @p st.shared.b64 [%1], w2;\n"
        "}" ::"r"(pair_addr), "r"(q_addr), "l"(fx), "l"(fq), "r"(pass_bit));
}

template <int MODE, bool MOMENTS, int STAGES, int K>
__global__ void __launch_bounds__(kBulkThreads, sql_ring_min_ctas(MODE, MOMENTS, K)) k_sql_ring(const SqlRingArgs ra) {
    extern __shared__ __align__(128) unsigned char sql_ring_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    const SqlArgs& a = ra.q;
    constexpr int T = kBulkConsumerWarps * 32;
    static_assert(K <= kSqlMaxRowsPerThread, "pass bits live in one register");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* ring = sql_ring_smem;
    unsigned char* bin_mem = sql_ring_smem + ra.ring_bytes;
    SqlBins<MODE, MOMENTS, T> bins;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kBulkConsumerWarps); }
        fence_barrier_init();
    }
    sql_member_load(a, tid, kBulkThreads);
    bins.init(bin_mem, a.n_groups, tid, kBulkThreads);  // ends with __syncthreads()
    bins.with_sums = a.agg_slot >= 0;
    bins.paired = MODE == 1 && a.pair_bins != 0 && a.agg_slot >= 0;
    bins.bias = a.fx_bias; bins.spill_acc = a.global_acc;

    const uint64_t n_main = a.count & ~3ull;  // bulk copies move multiples of 16 bytes: 4 rows of a 4-byte column
    const uint64_t ntiles = (n_main + ra.tile_rows - 1) / ra.tile_rows;
    if (warp == kBulkConsumerWarps) {
        if (lane == 0) {
            // column geometry in registers (one read of the parameter bank, not one per tile)
            const unsigned char* cptr[kSqlMaxCols];
            uint32_t cw[kSqlMaxCols], coff[kSqlMaxCols], row_bytes = 0;
#pragma unroll
            for (int k = 0; k < kSqlMaxCols; ++k) {
                const bool live = k < a.ncols;
                cptr[k] = live ? static_cast<const unsigned char*>(a.cols[k].ptr) : nullptr;
                cw[k] = live ? (a.cols[k].kind == 2 ? 4u : 8u) : 0u;
                coff[k] = live ? ra.col_off[k] : 0u;
                row_bytes += cw[k];
            }
            uint32_t it = 0;
            for (uint64_t c = blockIdx.x; c < ntiles; c += gridDim.x, ++it) {
                const int s = it % STAGES;
                const uint32_t round = it / STAGES;
                if (round > 0) mbar_wait_long(&empty_bar[s], (round - 1) & 1, 20000u);
                const uint64_t row0 = c * (uint64_t)ra.tile_rows;
                const uint32_t rows = (uint32_t)((n_main - row0) < (uint64_t)ra.tile_rows ? (n_main - row0) : (uint64_t)ra.tile_rows);
                unsigned char* stage = ring + (size_t)s * ra.stage_bytes;
                mbar_expect_tx(&full_bar[s], rows * row_bytes);
#pragma unroll
                for (int k = 0; k < kSqlMaxCols; ++k)
                    if (cw[k]) bulk_g2s(stage + coff[k], cptr[k] + row0 * cw[k], rows * cw[k], &full_bar[s]);
            }
        }
    } else {
        // per-query facts, read from the parameter bank once: which columns carry a predicate in which OR branch, where the
        // aggregate / group columns sit in a stage (a chain of indexed constant loads per tile costs 5-15 % with this few warps)
        const int agg_slot = a.agg_slot, group_slot = a.group_slot, n_alt = a.n_alt;
        const int agg_kind = agg_slot >= 0 ? a.cols[agg_slot].kind : -1;
        const int group_kind = group_slot >= 0 ? a.cols[group_slot].kind : -1;
        const uint32_t agg_off = agg_slot >= 0 ? ra.col_off[agg_slot] : 0u, group_off = group_slot >= 0 ? ra.col_off[group_slot] : 0u;
        const bool fuse_agg_pred = n_alt == 1 && agg_kind == 0;  // an f64 aggregate column is tested where it is converted (one LDS per row)
        uint32_t pred_cols[kSqlMaxAlt];
        int mod_slot = -1;
#pragma unroll
        for (int alt = 0; alt < kSqlMaxAlt; ++alt) {
            pred_cols[alt] = 0u;
            for (int k = 0; k < a.ncols; ++k)
                if (alt < n_alt && a.cols[k].pred[alt].has_pred && !(fuse_agg_pred && k == agg_slot)) pred_cols[alt] |= 1u << k;
        }
        for (int k = 0; k < a.ncols; ++k) if (a.cols[k].mod_step > 0) mod_slot = k;
        const SqlPred& ap = a.cols[agg_slot >= 0 ? agg_slot : 0].pred[0];
        const bool agg_has_pred = fuse_agg_pred && ap.has_pred != 0;
        const double agg_lo = __longlong_as_double(ap.lo), agg_hi = __longlong_as_double(ap.hi);
        const double agg_ne = ap.has_ne != 0 ? __longlong_as_double(ap.ne) : __longlong_as_double(0x7ff8000000000000ll);   // NaN: no `!=` in the conjunct
        // MODE 1 with squares, paired bins: this thread's column of bins as shared-memory addresses (sql_private_moments_row)
        const uint32_t my_pair = smem_u32(bins.p_slo) + (uint32_t)tid * 16u, my_q = smem_u32(bins.p_qlo) + (uint32_t)tid * 8u;
        const bool fast_moments = MODE == 1 && MOMENTS && bins.paired && a.pair_bins == 1;   // (AQE_SQL_PAIR_BINS=2: paired bins through SqlBins::add, the A/B)
        const int kmin32 = (int)a.key_min;
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        uint32_t it = 0;
        unsigned int rows_since_drain = 0;
        for (uint64_t c = blockIdx.x; c < ntiles; c += gridDim.x, ++it) {
            if constexpr (MODE == 1) {   // all consumer threads walk the same tiles: the drain's barrier is uniform over them
                if (rows_since_drain + K > a.drain_rows) { bins.drain(a.global_acc, tid); rows_since_drain = 0; }
                rows_since_drain += K;
            }
            const int s = it % STAGES;
            const uint32_t round = it / STAGES;
            mbar_wait_long_at(full0 + 8u * (uint32_t)s, round & 1, 20000u);
            const unsigned char* stage = ring + (size_t)s * ra.stage_bytes;
            // pass bits of rows tid, tid + T, ... of this tile; every tile but the table's last is whole (tile_rows = T * K)
            uint32_t mask = (1u << K) - 1u;
            if (c + 1 == ntiles) {
                const uint32_t rows = (uint32_t)(n_main - c * (uint64_t)ra.tile_rows);
                const uint32_t nk = rows > (uint32_t)tid ? (rows - (uint32_t)tid + T - 1) / T : 0u;
                mask = (1u << (nk < (uint32_t)K ? nk : (uint32_t)K)) - 1u;
            }
            if (n_alt == 1) {
                for (uint32_t pc = pred_cols[0]; pc; pc &= pc - 1) {
                    const int k = __ffs(pc) - 1;
                    mask = sql_pred_pass<T, K>(a.cols[k], a.cols[k].pred[0], stage + ra.col_off[k], tid, mask);
                }
            } else if (n_alt > 1) {  // OR of conjunctions: one mask per branch
                uint32_t any = 0u;
#pragma unroll
                for (int alt = 0; alt < kSqlMaxAlt; ++alt) {
                    if (alt >= n_alt) break;
                    uint32_t m = mask;
                    for (uint32_t pc = pred_cols[alt]; pc; pc &= pc - 1) {
                        const int k = __ffs(pc) - 1;
                        m = sql_pred_pass<T, K>(a.cols[k], a.cols[k].pred[alt], stage + ra.col_off[k], tid, m);
                    }
                    any |= m;
                }
                mask = any;
            }
            if (mod_slot >= 0) mask = sql_mod_pass<T, K>(a.cols[mod_slot], stage + ra.col_off[mod_slot], tid, mask);
            if (ra.samp_step > 1) {
                uint32_t x = (uint32_t)((c * (uint64_t)ra.tile_rows + ra.samp_phase + (uint32_t)tid) % ra.samp_step);  // (row + phase) mod step
                const uint32_t dt = (uint32_t)T % ra.samp_step;                                  // advanced by T mod step per owned row
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (x != 0) mask &= ~(1u << k);
                    x += dt;
                    if (x >= ra.samp_step) x -= ra.samp_step;
                }
            }
            unsigned int g[K];
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = 0;
            if constexpr (MODE != 0) {
                const unsigned char* gb = stage + group_off;
                if (group_kind == 2) {
#pragma unroll
                    for (int k = 0; k < K; ++k) g[k] = (unsigned int)(lds_row<int>(gb, tid + k * T) - kmin32);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) g[k] = (unsigned int)(lds_row<long long>(gb, tid + k * T) - a.key_min);
                }
                if constexpr (MODE == 1) {   // keys outside the layout (none among the live rows of a table the layout was taken from) land in bin G, which no drain reads
#pragma unroll
                    for (int k = 0; k < K; ++k) g[k] = min(g[k], bins.G);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) if (g[k] >= bins.G) mask &= ~(1u << k);  // cannot happen for live rows with this table's own layout
                }
            }
            // shared-atomic bins: rounds of the sparse walk when it pays (sql_sparse_adds), else ~0u
            auto sparse_rounds = [&](uint32_t m) -> unsigned int {
                if constexpr (MODE == 4) {   // (MODE 2 and 3 run under a 56-register cap that this second code path does not fit)
                    const unsigned int r = __reduce_max_sync(0xffffffffu, (unsigned int)__popc(m));
                    return 2u * r <= (unsigned int)K ? r : ~0u;
                } else return ~0u;
            };
            if (agg_kind < 0) {
                if constexpr (MODE == 0) bins.r_cnt += (unsigned long long)__popc(mask);   // COUNT without GROUP BY: the pass bits are the answer
                else {
#pragma unroll
                    for (int k = 0; k < K; ++k) if ((mask >> k) & 1u) bins.add(g[k], tid, false, 0, 0);
                }
            } else if (agg_kind == 0) {
                const unsigned char* ab = stage + agg_off;
                long long fx[K], fq[K];
                if (agg_has_pred) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double d = lds_row<double>(ab, tid + k * T);
                        mask = sql_f64_range_bit(mask, k, d, agg_lo, agg_hi, agg_ne);
                        fx[k] = __double2ll_rn(__dmul_rn(d, a.sum_scale));
                        fq[k] = MOMENTS ? __double2ll_rn(__dmul_rn(__dmul_rn(d, d), a.sq_scale)) : 0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double d = lds_row<double>(ab, tid + k * T);
                        fx[k] = __double2ll_rn(__dmul_rn(d, a.sum_scale));
                        fq[k] = MOMENTS ? __double2ll_rn(__dmul_rn(__dmul_rn(d, d), a.sq_scale)) : 0;
                    }
                }
                const unsigned int sr = sparse_rounds(mask);
                if (sr != ~0u) {
                    if constexpr (MODE == 4)
                        sql_sparse_adds<MOMENTS, T>(SqlSparseConsts{bins.s_sum, bins.s_sq, bins.spill_acc, bins.bias, a.key_min, a.sum_scale, a.sq_scale, bins.G, group_kind, agg_kind},
                                                    stage + group_off, ab, tid, mask, sr);
                } else if (MODE == 3 && mask == (1u << K) - 1u) {   // no WHERE clause: every tile but the table's last is whole -- no test per row
#pragma unroll
                    for (int k = 0; k < K; ++k) bins.add(g[k], tid, true, fx[k], fq[k]);
                } else if (fast_moments) {
#pragma unroll
                    for (int k = 0; k < K; ++k) sql_private_moments_row(my_pair + g[k] * (uint32_t)(T * 16), my_q + g[k] * (uint32_t)(T * 8), fx[k], fq[k], mask & (1u << k));
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) if ((mask >> k) & 1u) bins.add(g[k], tid, true, fx[k], fq[k]);
                }
            } else {
                const unsigned char* ab = stage + agg_off;
                long long fx[K], fq[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    fx[k] = agg_kind == 2 ? (long long)lds_row<int>(ab, tid + k * T) : lds_row<long long>(ab, tid + k * T);
                    const double d = (double)fx[k];
                    fq[k] = MOMENTS ? __double2ll_rn(__dmul_rn(__dmul_rn(d, d), a.sq_scale)) : 0;
                }
                const unsigned int sr = sparse_rounds(mask);
                if (sr != ~0u) {
                    if constexpr (MODE == 4)
                        sql_sparse_adds<MOMENTS, T>(SqlSparseConsts{bins.s_sum, bins.s_sq, bins.spill_acc, bins.bias, a.key_min, a.sum_scale, a.sq_scale, bins.G, group_kind, agg_kind},
                                                    stage + group_off, ab, tid, mask, sr);
                } else if (MODE == 3 && mask == (1u << K) - 1u) {   // no WHERE clause: every tile but the table's last is whole -- no test per row
#pragma unroll
                    for (int k = 0; k < K; ++k) bins.add(g[k], tid, true, fx[k], fq[k]);
                } else if (fast_moments) {
#pragma unroll
                    for (int k = 0; k < K; ++k) sql_private_moments_row(my_pair + g[k] * (uint32_t)(T * 16), my_q + g[k] * (uint32_t)(T * 8), fx[k], fq[k], mask & (1u << k));
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) if ((mask >> k) & 1u) bins.add(g[k], tid, true, fx[k], fq[k]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_at(empty0 + 8u * (uint32_t)s);
        }
        if (blockIdx.x == 0 && tid == 0) {  // the count % 4 tail
            for (uint64_t j = n_main; j < a.count; ++j) {
                long long raw[kSqlMaxCols];
#pragma unroll
                for (int k = 0; k < kSqlMaxCols; ++k) raw[k] = k < a.ncols ? sql_load_raw(a.cols[k], j) : 0;
                if (ra.samp_step > 1 && (j + ra.samp_phase) % ra.samp_step != 0) continue;
                sql_consume(a, bins, tid, raw);
            }
        }
    }
    bins.flush(a.global_acc, tid, kBulkThreads);
    sql_publish(a, tid, kBulkThreads);
}

}  // namespace aqe
