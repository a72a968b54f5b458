// aqe_sql_kernels.cuh -- sm_100a kernels of the SQL-string path.
//
//   k_col_stats   min / max of a column as order-preserving 64-bit keys (+ "ids are first_id + row" check)
//   k_sql_agg     the grouped scan: SELECT agg(col) FROM t [WHERE conjunction] [GROUP BY g] with the
//                 reference's `rowid % step = 0` sampling, one launch
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
//   G <= 4096     CTA-shared bins, 32-bit shared atomics with explicit carry propagation (64-bit shared
//                 atomic adds are CAS loops on sm_100: SASS ATOMS.CAST.SPIN.64)
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

struct SqlCol {
    const void* ptr;
    int kind;          // 0 f64, 1 i64, 2 i32
    int has_pred;      // closed interval [lo, hi] on this column (raw 64-bit: f64 bits or int64)
    int has_ne;
    int mod_step;      // > 0: also requires value % mod_step == 0  (rowid sampling when ids are not dense)
    long long lo, hi, ne;
};

struct SqlArgs {
    SqlCol cols[kSqlMaxCols];
    int ncols;
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
// four consecutive rows starting at a multiple of four (columns are 32-byte aligned on this path)
__device__ __forceinline__ void sql_load_raw4(const SqlCol& c, uint64_t i, long long (&v)[4]) {
    if (c.kind == 2) {
        const Vec<int32_t, 4> r = ldg_stream4(static_cast<const int32_t*>(c.ptr) + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (long long)r.v[e];
    } else {
        const Vec<int64_t, 4> r = ldg_stream4(static_cast<const int64_t*>(c.ptr) + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (long long)r.v[e];
    }
}
__device__ __forceinline__ bool sql_pass(const SqlCol& c, long long raw) {
    bool ok = true;
    if (c.has_pred) {
        if (c.kind == 0) {
            const double d = __longlong_as_double(raw);
            ok = d >= __longlong_as_double(c.lo) && d <= __longlong_as_double(c.hi);
            if (c.has_ne) ok = ok && d != __longlong_as_double(c.ne);
        } else {
            ok = raw >= c.lo && raw <= c.hi;
            if (c.has_ne) ok = ok && raw != c.ne;
        }
    }
    if (c.mod_step > 0) ok = ok && (raw % (long long)c.mod_step == 0);
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
__device__ __forceinline__ void shared_add128(unsigned int* limbs, long long v) {
    const unsigned int sign = v < 0 ? 0xffffffffu : 0u;
    unsigned int c = limb_add(limbs + 0, (unsigned int)v, 0u);
    c = limb_add(limbs + 1, (unsigned int)((unsigned long long)v >> 32), c);
    c = limb_add(limbs + 2, sign, c);
    limb_add(limbs + 3, sign, c);
}

// MODE 0: no GROUP BY (registers) | 1: thread-private shared bins | 2: CTA-shared bins with atomics
// VEC: rows are visited with stride 1 from a multiple of four and every column is 32-byte aligned -> 4 rows per load
template <int MODE, bool MOMENTS, bool VEC>
__global__ void __launch_bounds__(kSqlThreads) k_sql_agg(const SqlArgs a) {
    extern __shared__ __align__(16) unsigned char sql_smem[];
    const unsigned int G = a.n_groups;
    const int tid = threadIdx.x;
    constexpr int T = kSqlThreads;

    // ---- bins ----
    // MODE 1: cnt[G][T] u32 | slo[G][T] u64 | shi[G][T] u64 | (qlo, qhi)
    // MODE 2: cnt[G] u32 | sum limbs [G][4] u32 | (sq limbs [G][4])
    unsigned int* b_cnt = reinterpret_cast<unsigned int*>(sql_smem);
    unsigned long long* p_slo = nullptr; unsigned long long* p_shi = nullptr; unsigned long long* p_qlo = nullptr; unsigned long long* p_qhi = nullptr;
    unsigned int* s_sum = nullptr; unsigned int* s_sq = nullptr;
    if constexpr (MODE == 1) {
        p_slo = reinterpret_cast<unsigned long long*>(sql_smem + (size_t)G * T * 4);
        p_shi = p_slo + (size_t)G * T;
        p_qlo = p_shi + (size_t)G * T;
        p_qhi = p_qlo + (size_t)G * T;
        for (unsigned int g = 0; g < G; ++g) {
            b_cnt[g * T + tid] = 0; p_slo[g * T + tid] = 0; p_shi[g * T + tid] = 0;
            if constexpr (MOMENTS) { p_qlo[g * T + tid] = 0; p_qhi[g * T + tid] = 0; }
        }
    } else if constexpr (MODE == 2) {
        s_sum = b_cnt + G;
        s_sq = s_sum + (size_t)G * 4;
        for (unsigned int i = tid; i < G; i += T) b_cnt[i] = 0;
        for (unsigned int i = tid; i < G * 4; i += T) { s_sum[i] = 0; if constexpr (MOMENTS) s_sq[i] = 0; }
        __syncthreads();
    }
    // MODE 0 registers
    unsigned long long r_cnt = 0, r_slo = 0, r_qlo = 0;
    long long r_shi = 0, r_qhi = 0;

    auto consume = [&](const long long (&raw)[kSqlMaxCols]) {
        bool pass = true;
#pragma unroll
        for (int c = 0; c < kSqlMaxCols; ++c)
            if (c < a.ncols) pass = pass && sql_pass(a.cols[c], raw[c]);
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
            if (g >= G) return;  // cannot happen when the layout came from this table's statistics
        }
        if constexpr (MODE == 0) {
            r_cnt += 1;
            r_slo += (unsigned long long)fx & 0xffffffffull; r_shi += fx >> 32;
            if constexpr (MOMENTS) { r_qlo += (unsigned long long)fq & 0xffffffffull; r_qhi += fq >> 32; }
        } else if constexpr (MODE == 1) {
            const unsigned int s = g * T + tid;
            b_cnt[s] += 1;
            p_slo[s] += (unsigned long long)fx & 0xffffffffull; p_shi[s] += (unsigned long long)(fx >> 32);
            if constexpr (MOMENTS) { p_qlo[s] += (unsigned long long)fq & 0xffffffffull; p_qhi[s] += (unsigned long long)(fq >> 32); }
        } else {
            atomicAdd(b_cnt + g, 1u);
            if (a.agg_slot >= 0) {
                shared_add128(s_sum + g * 4, fx);
                if constexpr (MOMENTS) shared_add128(s_sq + g * 4, fq);
            }
        }
    };

    const uint64_t gsz = (uint64_t)gridDim.x * T;
    const uint64_t gtid = (uint64_t)blockIdx.x * T + tid;
    if constexpr (VEC) {
        const uint64_t units = a.count / 4;
        for (uint64_t u = gtid; u < units; u += gsz) {
            long long raw4[kSqlMaxCols][4];
#pragma unroll
            for (int c = 0; c < kSqlMaxCols; ++c)
                if (c < a.ncols) sql_load_raw4(a.cols[c], a.first + u * 4, raw4[c]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                long long raw[kSqlMaxCols];
#pragma unroll
                for (int c = 0; c < kSqlMaxCols; ++c) raw[c] = c < a.ncols ? raw4[c][e] : 0;
                consume(raw);
            }
        }
        if (gtid == 0) {
            for (uint64_t j = units * 4; j < a.count; ++j) {
                long long raw[kSqlMaxCols];
#pragma unroll
                for (int c = 0; c < kSqlMaxCols; ++c) raw[c] = c < a.ncols ? sql_load_raw(a.cols[c], a.first + j) : 0;
                consume(raw);
            }
        }
    } else {
        // strided / unaligned visit: four independent rows in flight per thread
        uint64_t j = gtid;
        for (; j + 3 * gsz < a.count; j += 4 * gsz) {
            long long raw4[4][kSqlMaxCols];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int c = 0; c < kSqlMaxCols; ++c)
                    raw4[e][c] = c < a.ncols ? sql_load_raw(a.cols[c], a.first + (j + (uint64_t)e * gsz) * a.stride) : 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) consume(raw4[e]);
        }
        for (; j < a.count; j += gsz) {
            long long raw[kSqlMaxCols];
#pragma unroll
            for (int c = 0; c < kSqlMaxCols; ++c) raw[c] = c < a.ncols ? sql_load_raw(a.cols[c], a.first + j * a.stride) : 0;
            consume(raw);
        }
    }

    // ---- CTA totals -> global accumulators (integer atomics: order does not matter) ----
    if constexpr (MODE == 0) {
        r_cnt = warp_reduce_u64(r_cnt);
        r_slo = warp_reduce_u64(r_slo); r_shi = (long long)warp_reduce_u64((unsigned long long)r_shi);
        if constexpr (MOMENTS) { r_qlo = warp_reduce_u64(r_qlo); r_qhi = (long long)warp_reduce_u64((unsigned long long)r_qhi); }
        if ((tid & 31) == 0 && r_cnt) {
            unsigned long long lo, hi;
            atomicAdd(a.global_acc + 0, r_cnt);
            split_to_128(r_slo, r_shi, lo, hi); global_add128(a.global_acc + 1, lo, hi);
            if constexpr (MOMENTS) { split_to_128(r_qlo, r_qhi, lo, hi); global_add128(a.global_acc + 3, lo, hi); }
        }
    } else if constexpr (MODE == 1) {
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (unsigned int g = warp; g < G; g += T / 32) {
            unsigned long long c = 0, sl = 0, sh = 0, ql = 0, qh = 0;
#pragma unroll
            for (int k = 0; k < T / 32; ++k) {
                const unsigned int s = g * T + lane + 32 * k;
                c += b_cnt[s]; sl += p_slo[s]; sh += p_shi[s];
                if constexpr (MOMENTS) { ql += p_qlo[s]; qh += p_qhi[s]; }
            }
            c = warp_reduce_u64(c); sl = warp_reduce_u64(sl); sh = warp_reduce_u64(sh);
            if constexpr (MOMENTS) { ql = warp_reduce_u64(ql); qh = warp_reduce_u64(qh); }
            if (lane == 0 && c) {
                unsigned long long lo, hi;
                unsigned long long* ga = a.global_acc + (size_t)g * 5;
                atomicAdd(ga + 0, c);
                split_to_128(sl, (long long)sh, lo, hi); global_add128(ga + 1, lo, hi);
                if constexpr (MOMENTS) { split_to_128(ql, (long long)qh, lo, hi); global_add128(ga + 3, lo, hi); }
            }
        }
    } else {
        __syncthreads();
        for (unsigned int g = tid; g < G; g += T) {
            const unsigned int c = b_cnt[g];
            if (!c) continue;
            unsigned long long* ga = a.global_acc + (size_t)g * 5;
            atomicAdd(ga + 0, (unsigned long long)c);
            const unsigned int* l = s_sum + g * 4;
            global_add128(ga + 1, ((unsigned long long)l[1] << 32) | l[0], ((unsigned long long)l[3] << 32) | l[2]);
            if constexpr (MOMENTS) {
                const unsigned int* m = s_sq + g * 4;
                global_add128(ga + 3, ((unsigned long long)m[1] << 32) | m[0], ((unsigned long long)m[3] << 32) | m[2]);
            }
        }
    }

    // ---- last CTA: publish and re-arm ----
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (unsigned int i = tid; i < G * 5; i += T) {
        a.out[i] = __ldcg(a.global_acc + i);
        a.global_acc[i] = 0ull;
    }
    if (tid == 0) *a.ticket = 0u;
}

}  // namespace aqe
