// aqe_device.cuh -- device-side building blocks shared by the sm_100a kernels of libaqe_b200.
//
// Nothing here is generic CUDA plumbing for its own sake: every helper exists because one of the
// kernels K1..K7 needs it.  Compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "aqe_b200.h"

namespace aqe {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG).  Same constants / round function as oracle/aqe_oracle.c section 1.
// ------------------------------------------------------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

constexpr uint32_t kSynthStream = 0x41514544u;  // "AQED": synthetic table stream (counter word 2)
constexpr uint32_t kSeedStream = 0x53454544u;   // "SEED": stand-in for std::random_device draws
constexpr uint32_t kDrawStream = 0x53525330u;   // "SRS0": sample positions of the persistent CLT kernel

__host__ __device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// j-th sample position of the draw stream: uniform in [0, units) by multiply-shift.
__host__ __device__ __forceinline__ uint64_t draw_position(uint64_t seed, uint32_t design, uint64_t j, uint64_t units) {
    const uint64_t c = j >> 1;
    const u32x4 r = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), kDrawStream | design, 0u, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    const uint64_t u = (j & 1) ? (((uint64_t)r.w << 32) | r.z) : (((uint64_t)r.y << 32) | r.x);
    return mulhi64(u, units);
}

// Student-t quantile from the normal quantile z at the same probability (Cornish-Fisher / Fisher's expansion in 1/df):
// t = z + (z^3 + z)/(4 df) + (5 z^5 + 16 z^3 + 3 z)/(96 df^2).  df >= 15 here (the smallest look is 16 units): error < 1e-4.
__host__ __device__ __forceinline__ double t_from_z(double z, double df) {
    const double z2 = z * z;
    return z + z * (z2 + 1.0) / (4.0 * df) + z * ((5.0 * z2 + 16.0) * z2 + 3.0) / (96.0 * df * df);
}

// Pseudo-random permutation of [0, n): a 4-round balanced Feistel network over the smallest even number of bits
// covering n, splitmix64 as the round function, cycle-walking back into [0, n).  perm(0), perm(1), ... perm(k-1) are
// k DISTINCT, uniformly scattered positions -- simple random sampling WITHOUT replacement in O(1) memory, computed
// where it is needed (replaces copy-all + std::shuffle of sample_records, custom_bplus_db.cpp:345-363).
__host__ __device__ __forceinline__ uint64_t feistel_round(uint64_t r, uint32_t round, uint64_t seed) {
    uint64_t z = (r + 0x9E3779B97F4A7C15ull * (uint64_t)(round + 1)) ^ seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t feistel_half_bits(uint64_t n) {  // half width h: 2^(2h) >= n, h >= 1
    uint32_t h = 1;
    while (h < 32 && (1ull << (2 * h)) < n) ++h;
    return h;
}
__host__ __device__ __forceinline__ uint64_t feistel_perm(uint64_t k, uint64_t n, uint64_t seed, uint32_t half) {
    const uint64_t mask = (1ull << half) - 1;
    uint64_t x = k;
    do {
        uint64_t l = x >> half, r = x & mask;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            const uint64_t f = feistel_round(r, i, seed) & mask;
            const uint64_t t = l ^ f;
            l = r; r = t;
        }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}

// One row of the synthetic sales table.  UNIFORM is bit-identical on host and device.
__host__ __device__ __forceinline__ void synth_row(uint64_t seed, uint64_t row, int dist, aqe_record& r) {
    const u32x4 o = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), kSynthStream, 0u, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    const uint64_t u = (((uint64_t)o.x << 32) | o.y) >> 11;
    const double x = (double)u * (1.0 / 9007199254740992.0);
    r.id = (int64_t)row + 1;
    if (dist == AQE_SYNTH_LOGNORMAL) {
        const u32x4 q = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), kSynthStream, 1u, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
        const double u1 = (double)((((uint64_t)q.x << 32) | q.y) >> 11) * (1.0 / 9007199254740992.0);
        const double u2 = (double)((((uint64_t)q.z << 32) | q.w) >> 11) * (1.0 / 9007199254740992.0);
        const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
        r.amount = exp(4.0 + 1.5 * z);
    } else {
#ifdef __CUDA_ARCH__
        r.amount = __dadd_rn(1.0, __dmul_rn(999.0, x));  // two roundings, never an FMA
#else
        volatile double t = 999.0 * x;
        r.amount = 1.0 + t;
#endif
    }
    r.region = (int32_t)(o.z & 7u);
    r.product_id = (int32_t)(o.w % 1000u);
    r.timestamp = 1700000000LL + (int64_t)row;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Compensated fp64 accumulation.  (s, c): s is the running rounded sum, c collects the rounding errors
// (Knuth TwoSum, branch free).  s + c is the sum to roughly twice the working precision, so the result is
// independent of grid geometry / shard count except in razor-edge roundings.
// ------------------------------------------------------------------------------------------------
struct DD { double s, c; };

__device__ __forceinline__ void dd_add(DD& a, double x) {
    const double t = __dadd_rn(a.s, x);
    const double z = __dadd_rn(t, -a.s);
    a.c = __dadd_rn(a.c, __dadd_rn(__dadd_rn(a.s, -__dadd_rn(t, -z)), __dadd_rn(x, -z)));
    a.s = t;
}
__device__ __forceinline__ void dd_merge(DD& a, const DD& b) {
    const double t = __dadd_rn(a.s, b.s);
    const double z = __dadd_rn(t, -a.s);
    const double e = __dadd_rn(__dadd_rn(a.s, -__dadd_rn(t, -z)), __dadd_rn(b.s, -z));
    a.c = __dadd_rn(__dadd_rn(a.c, b.c), e);
    a.s = t;
}
// renormalise so that s is the correctly rounded value of s + c
__device__ __forceinline__ void dd_norm(DD& a) {
    const double t = __dadd_rn(a.s, a.c);
    a.c = __dadd_rn(a.c, -__dadd_rn(t, -a.s));
    a.s = t;
}

// ------------------------------------------------------------------------------------------------
// Streaming loads: read-only path, no L1 allocation (each byte is used once per scan).  256-bit forms
// are sm_100+ (SASS LDG.E.256); they need 32-byte alignment.
// ------------------------------------------------------------------------------------------------
template <typename T, int N> struct Vec { T v[N]; };

__device__ __forceinline__ Vec<double, 4> ldg_stream4(const double* p) {
    Vec<double, 4> r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<int64_t, 4> ldg_stream4(const int64_t* p) {
    Vec<int64_t, 4> r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<int32_t, 4> ldg_stream4(const int32_t* p) {
    Vec<int32_t, 4> r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<double, 2> ldg_stream2(const double* p) {
    Vec<double, 2> r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<int64_t, 2> ldg_stream2(const int64_t* p) {
    Vec<int64_t, 2> r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.v[0]), "=l"(r.v[1]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<int32_t, 2> ldg_stream2(const int32_t* p) {
    Vec<int32_t, 2> r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "l"(p));
    return r;
}
__device__ __forceinline__ Vec<int32_t, 8> ldg_stream8(const int32_t* p) {
    Vec<int32_t, 8> r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
template <int W, typename T> __device__ __forceinline__ Vec<T, W> ldg_stream(const T* p) {
    if constexpr (W == 8) return ldg_stream8(p);
    else if constexpr (W == 4) return ldg_stream4(p);
    else return ldg_stream2(p);
}

// Scheduling fence: an empty volatile asm that "modifies" every register of a loaded vector.  Volatile asms keep
// their program order, so placing these after a batch of ldg_stream() calls forces ALL loads of the batch to be
// issued before the first value is consumed (ptxas otherwise sinks later loads below earlier uses to save
// registers, halving the bytes in flight -- seen on the integer-predicate scan, 3.9 instead of 7.3 TB/s).
__device__ __forceinline__ void pin(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void pin(int64_t& x) { asm volatile("" : "+l"(x)); }
__device__ __forceinline__ void pin(int32_t& x) { asm volatile("" : "+r"(x)); }
template <typename T, int N> __device__ __forceinline__ void pin(Vec<T, N>& v) {
#pragma unroll
    for (int i = 0; i < N; ++i) pin(v.v[i]);
}

// ------------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) for the smem-staged scan variant.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same with a suspend-time hint: the thread may stay suspended in the barrier unit for up to `ns` before try_wait returns false, so a
// warp that waits long polls rarely (k_sql_ring: the consumers of the shared-bin kernels spent 15 % of their instructions on these polls)
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITL_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEL_%=;\n"
        "bra WAITL_%=;\n"
        "DONEL_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
// the same two on a barrier's shared-memory address (taken once outside a tile loop: the generic-to-shared conversion of a
// __shared__ array element costs an S2UR / ULEA sequence per call)
__device__ __forceinline__ void mbar_arrive_at(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_long_at(uint32_t bar_addr, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEA_%=;\n"
        "bra WAITA_%=;\n"
        "DONEA_%=:\n"
        "}\n" ::"r"(bar_addr), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy, completion signalled on `bar` as transaction bytes; L2 evict-first hint
// is NOT set: a 10 M-row column (80 MB) fits the 126 MB L2 and repeated queries should hit it.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// Fixed-order reductions: shuffle-down tree inside a warp, then warp partials through shared memory in
// warp order.  No floating-point atomics anywhere, so results are run-to-run bit stable.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_down_f64(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t shfl_down_u64(uint64_t v, int d) {
    return (uint64_t)__shfl_down_sync(0xffffffffu, (unsigned long long)v, d);
}
__device__ __forceinline__ DD warp_reduce_dd(DD a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        DD b{shfl_down_f64(a.s, d), shfl_down_f64(a.c, d)};
        dd_merge(a, b);
    }
    return a;
}
__device__ __forceinline__ uint64_t warp_reduce_u64(uint64_t a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += shfl_down_u64(a, d);
    return a;
}
__device__ __forceinline__ double warp_reduce_min(double a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = fmin(a, shfl_down_f64(a, d));
    return a;
}
__device__ __forceinline__ double warp_reduce_max(double a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = fmax(a, shfl_down_f64(a, d));
    return a;
}
#endif  // __CUDACC__

}  // namespace aqe
