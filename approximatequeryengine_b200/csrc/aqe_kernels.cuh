// aqe_kernels.cuh -- the sm_100a kernels of libaqe_b200.
//
//   K1/K2  k_scan_ring / k_scan   exact SUM/COUNT[/WHERE] full scan, f64 (compensated) and int128
//   K3/K5  k_plan_stats           moments of a column over a sample plan (affine segments / index list)
//   K4     k_approx               persistent Philox + Welford-by-shifted-sums + in-kernel CLT stop rule
//   K6     k_plan_gather          rows at plan positions -> 32-byte AoS (legacy list[Record] path)
//   K7     k_aos_to_soa           ingest: 32-byte file rows -> 5 column arrays
//          k_synth                device-side generator of the synthetic sales table
//
// Reference loops replaced (src/aqe_backend/core/custom_bplus_db.cpp): sum_amount :242-251,
// sum_amount_where :263-274, every `samples.push_back(all_records[i])` loop (:737-1960), the thread pool
// of clt_validated_dual_pointer_sample :885-1043, load_from_file :685-711.
#pragma once

#include <cooperative_groups.h>
#include <type_traits>

#include "aqe_device.cuh"

namespace aqe {
namespace cg = cooperative_groups;

// Read a POD written by another SM: L2 (ld.global.cg), never a stale L1 line.
template <typename T> __device__ __forceinline__ T load_cg(const T* p) {
    static_assert(sizeof(T) % 8 == 0, "8-byte multiple");
    T r;
    const unsigned long long* s = reinterpret_cast<const unsigned long long*>(p);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 8); ++i) d[i] = __ldcg(s + i);
    return r;
}

// ================================================================================================
// Shared accumulator of the scan kernels and its fixed-order block reduction
// ================================================================================================
struct ScanAcc {
    uint64_t count;
    DD sum;
    uint64_t ilo;  // sum of the low 32 bits of each integer value (as u64)
    int64_t ihi;   // sum of the (arithmetic) high 32 bits
    DD sq;
    double mn, mx;
};

__device__ __forceinline__ ScanAcc scan_identity() {
    ScanAcc a;
    a.count = 0; a.sum = DD{0.0, 0.0}; a.ilo = 0; a.ihi = 0; a.sq = DD{0.0, 0.0};
    a.mn = __longlong_as_double(0x7ff0000000000000LL); a.mx = __longlong_as_double(0xfff0000000000000LL);
    return a;
}
__device__ __forceinline__ void scan_merge(ScanAcc& a, const ScanAcc& b) {
    a.count += b.count; dd_merge(a.sum, b.sum); a.ilo += b.ilo; a.ihi += b.ihi; dd_merge(a.sq, b.sq);
    a.mn = fmin(a.mn, b.mn); a.mx = fmax(a.mx, b.mx);
}
template <bool IS_INT, bool MOMENTS> __device__ __forceinline__ ScanAcc scan_warp_reduce(ScanAcc a) {
    a.count = warp_reduce_u64(a.count);
    if constexpr (IS_INT) {
        a.ilo = warp_reduce_u64(a.ilo);
        a.ihi = (int64_t)warp_reduce_u64((uint64_t)a.ihi);
    } else {
        a.sum = warp_reduce_dd(a.sum);
        if constexpr (MOMENTS) {
            a.sq = warp_reduce_dd(a.sq);
            a.mn = warp_reduce_min(a.mn);
            a.mx = warp_reduce_max(a.mx);
        }
    }
    return a;
}
// Result valid in thread 0.  Order: lanes by shuffle-down tree, then warps 0..nw-1 by the same tree.
template <bool IS_INT, bool MOMENTS> __device__ __forceinline__ ScanAcc scan_block_reduce(ScanAcc a, ScanAcc* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = scan_warp_reduce<IS_INT, MOMENTS>(a);
    __syncthreads();  // sm may still be read by a previous call
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (warp == 0) {
        a = lane < nw ? sm[lane] : scan_identity();
        a = scan_warp_reduce<IS_INT, MOMENTS>(a);
    }
    return a;
}

// Cross-GPU exchange fused into the scan kernel (one process per GPU, peers mapped with CUDA IPC over
// NVLink).  Every rank owns a mailbox of kMaxRanks x 2 slots; the LAST block of rank g stores its 64-byte
// shard partial + the query's sequence number into slot [g][seq & 1] of EVERY rank's mailbox (peer stores),
// then waits until its own mailbox holds sequence `seq` from all ranks and folds them in rank order -- the
// all-gather + merge that followed the scan (NCCL launch + kernel) becomes a few NVLink stores inside it.
// Two slots per sender suffice: a rank can write query k+2 only after it merged query k+1, which needs every
// peer's k+1 partial, which a peer publishes only after it consumed query k.
constexpr int kMaxRanks = 16;
struct ExSlot { aqe_partial p; unsigned long long seq; unsigned long long pad[7]; };  // 128 bytes
struct Exchange {
    int world;                 // 0/1 = disabled
    int rank;
    int is_integer;
    int split;                 // 1: the scan kernel leaves its partial in `local`; k_scan_merge (next in the stream) publishes it, waits for the peers and folds
    aqe_partial* local;        // this GPU's memory
    unsigned long long seq;
    unsigned long long timeout_cycles;
    ExSlot* peers[kMaxRanks];  // peers[r] = rank r's mailbox as mapped in this process (peers[rank] = own)
    unsigned int* status;      // set to 1 if a peer did not show up in time (a word of the handle's mapped pinned slot: the host reads it without a copy)
};

struct ScanArgs {
    const void* agg;
    const void* pred;
    uint64_t n;
    double lo, hi;
    long long ilo, ihi;     // the same closed interval on an INTEGER predicate column: {v : lo <= (double)v <= hi} = [ilo, ihi]
    ScanAcc* partials;      // [gridDim.x]
    unsigned int* ticket;   // zero before launch; reset by the last block
    aqe_partial* out;       // device-visible (device memory or mapped pinned host memory)
    Exchange ex;
    // Programmatic dependent launch (back-to-back scans of one stream): != 0 = every CTA says at its start that the NEXT scan of the
    // stream may be launched.  The ring kernel is launched with exactly as many CTAs as the GPU holds (2 per SM, shared memory padded
    // so that a third cannot fit), so the next scan's CTAs wait for slots and take them one by one as this scan's CTAs exit: no idle
    // tail, no launch gap, still two CTAs on every SM (a first version that let a third CTA squeeze in next to a draining scan left
    // the next scan's CTAs unevenly spread over the SMs and ran 10-25 % SLOWER, profiles/r2_pdl_ab.jsonl).  Everything a scan writes
    // (block partials, ticket, mailboxes, result) happens behind griddep_wait(), i.e. after the previous scan has completed:
    // stream order is kept for every side effect, only the read-only streaming overlaps the previous scan's fold and exchange.
    unsigned int pdl_tail;
    // Static tile schedule of the ring kernel.  Rounds [0, even_rounds) deal tiles round-robin to all CTAs (tile = round * grid +
    // block); the remaining tiles go round-robin to the first `long_ctas` CTAs only (long_ctas = grid: the plain schedule).  With two
    // CTAs per SM and long_ctas = half the grid, the two CTAs of an SM finish a few tiles apart, so that in a stream of back-to-back
    // scans (programmatic dependent launch) one of them is always streaming while the other folds its partial and the next scan's CTA
    // fills its pipeline -- otherwise both slots of every SM go through that ~4 us hand-over at the same time.
    unsigned long long even_rounds;
    unsigned int long_ctas;
};

// it-th tile of CTA `block` under the schedule above (>= ntiles: the CTA is done)
__device__ __forceinline__ uint64_t scan_tile_of(const ScanArgs& a, uint64_t it, unsigned int block, unsigned int grid, uint64_t ntiles) {
    if (it < a.even_rounds) return it * grid + block;
    if (block >= a.long_ctas) return ntiles;
    return a.even_rounds * grid + (it - a.even_rounds) * a.long_ctas + block;
}

__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// (hi:lo) two's-complement 128-bit integer -> double, correctly rounded like the host's (double)(__int128): through the
// magnitude, so a small negative total is not lost in the rounding of its 2^64-complement low word.
__device__ __forceinline__ double i128_to_double(unsigned long long lo, long long hi) {
    const bool neg = hi < 0;
    unsigned long long mlo = lo, mhi = (unsigned long long)hi;
    if (neg) { mlo = ~lo + 1ull; mhi = ~(unsigned long long)hi + (mlo == 0ull ? 1ull : 0ull); }
    // mhi * 2^64 + mlo with ONE rounding: split mlo so that the partial sums below are exact, then add once
    double r;
    if (mhi == 0ull) r = __ull2double_rn(mlo);
    else {
        // 128-bit magnitude: normalise to 64 significant bits + sticky, convert, scale
        const int lz = __clzll((long long)mhi);
        unsigned long long top = lz ? ((mhi << lz) | (mlo >> (64 - lz))) : mhi;
        const unsigned long long rest = lz ? (mlo << lz) : mlo;
        if (rest) top |= 1ull;   // sticky bit: top has 64 bits, a double keeps 53, so bit 0 never reaches the mantissa
        r = ldexp(__ull2double_rn(top), 64 - lz);
    }
    return neg ? -r : r;
}

// Fixed-rank-order fold of shard partials; the same IEEE operations as the host's aqe_merge_partials
// (aqe_engine.cu), so device-merged and host-merged results are bit-identical.
__device__ __forceinline__ void merge_partials_dev(const aqe_partial* parts, int n, bool is_integer, aqe_partial* out) {
    aqe_partial r;
    r.count = 0; r.isum_lo = 0; r.isum_hi = 0;
    r.minv = __longlong_as_double(0x7ff0000000000000LL); r.maxv = __longlong_as_double(0xfff0000000000000LL);
    double s = 0.0, c = 0.0, q = 0.0;
    for (int i = 0; i < n; ++i) {
        const aqe_partial& p = parts[i];
        r.count += p.count;
        const unsigned long long lo = r.isum_lo + p.isum_lo;
        r.isum_hi += p.isum_hi + (lo < r.isum_lo ? 1 : 0);
        r.isum_lo = lo;
        const double t = __dadd_rn(s, p.sum);
        const double z = __dadd_rn(t, -s);
        const double e = __dadd_rn(__dadd_rn(s, -__dadd_rn(t, -z)), __dadd_rn(p.sum, -z));
        c = __dadd_rn(c, __dadd_rn(p.comp, e));
        s = t;
        q = __dadd_rn(q, p.sumsq);
        if (p.count) { r.minv = fmin(r.minv, p.minv); r.maxv = fmax(r.maxv, p.maxv); }
    }
    const double t = __dadd_rn(s, c);
    r.comp = __dadd_rn(c, -__dadd_rn(t, -s));
    r.sum = t;
    r.sumsq = q;
    if (is_integer) {
        r.sum = i128_to_double(r.isum_lo, r.isum_hi);
        r.comp = 0.0;
    }
    *out = r;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool IS_INT> __device__ __forceinline__ void scan_write_out(const ScanAcc& t, aqe_partial* out) {
    aqe_partial r;
    r.count = t.count;
    if constexpr (IS_INT) {
        // value = ihi * 2^32 + ilo  in 128-bit two's complement
        const uint64_t lo = t.ilo + ((uint64_t)t.ihi << 32);
        const int64_t hi = (t.ihi >> 32) + (lo < t.ilo ? 1 : 0);
        r.isum_lo = lo; r.isum_hi = hi;
        r.sum = i128_to_double(lo, hi);
        r.comp = 0.0; r.sumsq = 0.0; r.minv = 0.0; r.maxv = 0.0;
    } else {
        DD s = t.sum; dd_norm(s);
        DD q = t.sq; dd_norm(q);
        r.sum = s.s; r.comp = s.c; r.sumsq = q.s; r.minv = t.mn; r.maxv = t.mx;
        r.isum_lo = 0; r.isum_hi = 0;
    }
    *out = r;
}

// Second stage: the last block to finish folds the per-block partials in block order.
template <bool IS_INT, bool MOMENTS> __device__ __forceinline__ void scan_finish(ScanAcc block_total, const ScanArgs& a, ScanAcc* sm) {
    __shared__ bool is_last;
    griddep_wait();   // the previous scan of this stream (launched with programmatic serialization) has completed and is visible
    if (threadIdx.x == 0) {
        a.partials[blockIdx.x] = block_total;
        __threadfence();
        const unsigned int t = atomicAdd(a.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    ScanAcc acc = scan_identity();
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        scan_merge(acc, load_cg(a.partials + b));
    }
    acc = scan_block_reduce<IS_INT, MOMENTS>(acc, sm);
    if (a.ex.world <= 1) {
        if (threadIdx.x == 0) {
            scan_write_out<IS_INT>(acc, a.out);
            *a.ticket = 0u;
        }
        return;
    }
    if (a.ex.split) {   // publish / wait / fold happen in k_scan_merge: this CTA's SM slot is wanted by the next scan of the stream
        if (threadIdx.x == 0) {
            scan_write_out<IS_INT>(acc, a.ex.local);
            *a.ticket = 0u;
        }
        return;
    }
    // ---- fused all-gather + merge over peer memory ----
    __shared__ aqe_partial sh_parts[kMaxRanks];
    __shared__ aqe_partial sh_local;
    const int world = a.ex.world, par = (int)(a.ex.seq & 1ull);
    if (threadIdx.x == 0) {
        scan_write_out<IS_INT>(acc, &sh_local);
        *a.ticket = 0u;
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        ExSlot* dst = a.ex.peers[threadIdx.x] + (a.ex.rank * 2 + par);
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&sh_local);
        volatile unsigned long long* d = reinterpret_cast<volatile unsigned long long*>(&dst->p);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(aqe_partial) / 8); ++i) d[i] = src[i];
        __threadfence_system();
        st_release_sys(&dst->seq, a.ex.seq);
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const ExSlot* src = a.ex.peers[a.ex.rank] + (threadIdx.x * 2 + par);
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(&src->seq) != a.ex.seq) {
            if ((unsigned long long)(clock64() - t0) > a.ex.timeout_cycles) { ok = false; break; }
            __nanosleep(64);
        }
        if (!ok) *(volatile unsigned int*)a.ex.status = 1u;
        const volatile unsigned long long* sp = reinterpret_cast<const volatile unsigned long long*>(&src->p);
        unsigned long long* dp = reinterpret_cast<unsigned long long*>(&sh_parts[threadIdx.x]);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(aqe_partial) / 8); ++i) dp[i] = sp[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) merge_partials_dev(sh_parts, world, a.ex.is_integer != 0, a.out);
}

// The exchange proper as its own one-warp kernel, launched right behind the scan on the same stream: stores this shard's partial
// + the query's sequence number into every rank's mailbox (peer stores over NVLink), waits until this rank's mailbox holds `seq`
// from every rank, folds the partials in rank order and writes the table-level result.  Keeping all of that out of the scan grid
// matters for back-to-back queries: the scan CTA that would publish (system-scope fence + NVLink stores) and wait for a slower GPU
// holds an SM slot the next scan's CTA is queued for (static tile assignment), and the whole next query ends that much later
// (measured at 8 GPUs: 148.0 us per query with publish + wait inside the scan, 146.8 with the wait moved out, 141.7 for the scan
// alone).  One warp, 1 KiB of shared memory: it fits next to the two scan CTAs of any SM.
__global__ void __launch_bounds__(32) k_scan_merge(const Exchange ex, aqe_partial* out) {
    __shared__ aqe_partial sh_parts[kMaxRanks];
    griddep_launch_dependents();   // the next scan of the stream may be launched (its CTAs queue for SM slots)
    griddep_wait();                // this rank's scan has completed: its partial is in ex.local
    const int world = ex.world, par = (int)(ex.seq & 1ull);
    if ((int)threadIdx.x < world) {
        ExSlot* dst = ex.peers[threadIdx.x] + (ex.rank * 2 + par);
        const unsigned long long* lp = reinterpret_cast<const unsigned long long*>(ex.local);
        volatile unsigned long long* d = reinterpret_cast<volatile unsigned long long*>(&dst->p);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(aqe_partial) / 8); ++i) d[i] = __ldcg(lp + i);
        __threadfence_system();
        st_release_sys(&dst->seq, ex.seq);
    }
    __syncwarp();
    if ((int)threadIdx.x < world) {
        const ExSlot* src = ex.peers[ex.rank] + (threadIdx.x * 2 + par);
        const long long t0 = clock64();
        while (ld_acquire_sys(&src->seq) != ex.seq) {
            if ((unsigned long long)(clock64() - t0) > ex.timeout_cycles) { *(volatile unsigned int*)ex.status = 1u; break; }
            __nanosleep(64);
        }
        const volatile unsigned long long* sp = reinterpret_cast<const volatile unsigned long long*>(&src->p);
        unsigned long long* dp = reinterpret_cast<unsigned long long*>(&sh_parts[threadIdx.x]);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(aqe_partial) / 8); ++i) dp[i] = sp[i];
    }
    __syncwarp();
    if (threadIdx.x == 0) merge_partials_dev(sh_parts, world, ex.is_integer != 0, out);
}

template <typename T> __device__ __forceinline__ double as_f64(T v) { return (double)v; }

// PRED: 0 none, 1 predicate on the aggregate column itself, 2 predicate on another column (PredT)
// Integer predicate columns are compared as integers against [ilo, ihi] (computed once on the host so that it
// selects exactly the rows `lo <= (double)v <= hi` would): no per-row I2F.F64 conversions on the FP64 pipe.
template <typename AggT, int PRED, typename PredT, bool MOMENTS>
__device__ __forceinline__ void scan_consume(ScanAcc& acc, DD& alt, AggT v, PredT pv, const ScanArgs& a, int parity) {
    const double lo = a.lo, hi = a.hi;
    bool pass = true;
    if constexpr (PRED == 1) {
        if constexpr (std::is_integral_v<AggT>) pass = ((long long)v >= a.ilo) && ((long long)v <= a.ihi);
        else { const double d = as_f64(v); pass = (d >= lo) && (d <= hi); }
    }
    if constexpr (PRED == 2) {
        if constexpr (std::is_integral_v<PredT>) pass = ((long long)pv >= a.ilo) && ((long long)pv <= a.ihi);
        else { const double d = as_f64(pv); pass = (d >= lo) && (d <= hi); }
    }
    acc.count += pass ? 1u : 0u;
    if constexpr (std::is_integral_v<AggT>) {  // int64 / int32: split so a thread never overflows
        const int64_t x = pass ? (int64_t)v : 0;
        acc.ilo += (uint64_t)x & 0xffffffffull; acc.ihi += x >> 32;
    } else {  // f64
        const double x = pass ? (double)v : 0.0;
        if (parity) dd_add(alt, x); else dd_add(acc.sum, x);
        if constexpr (MOMENTS) {
            acc.sq.s = fma(x, x, acc.sq.s);
            acc.mn = fmin(acc.mn, pass ? (double)v : acc.mn);
            acc.mx = fmax(acc.mx, pass ? (double)v : acc.mx);
        }
    }
}

constexpr int kScanThreads = 256;

// K1/K2, register-staged variant: W-element vector loads (W=4: LDG.256 for 8-byte columns), U
// independent loads in flight per thread, grid-stride so every warp instruction reads one contiguous
// 32*W*sizeof(T)-byte run.
template <typename AggT, int PRED, typename PredT, int W, int U, bool MOMENTS, int MINB>
__global__ void __launch_bounds__(kScanThreads, MINB) k_scan(const ScanArgs a) {
    constexpr bool IS_INT = std::is_integral_v<AggT>;
    __shared__ ScanAcc sm[32];
    const AggT* __restrict__ agg = static_cast<const AggT*>(a.agg);
    const PredT* __restrict__ pred = static_cast<const PredT*>(a.pred);
    const uint64_t units = a.n / W;
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ScanAcc acc = scan_identity();
    DD alt{0.0, 0.0};

    for (; u + (uint64_t)(U - 1) * G < units; u += (uint64_t)U * G) {
        Vec<AggT, W> av[U];
        Vec<PredT, W> pv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            av[j] = ldg_stream<W>(agg + (u + (uint64_t)j * G) * W);
            if constexpr (PRED == 2) pv[j] = ldg_stream<W>(pred + (u + (uint64_t)j * G) * W);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {  // all 2U loads are in flight before the first use
            pin(av[j]);
            if constexpr (PRED == 2) pin(pv[j]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
            for (int e = 0; e < W; ++e)
                scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, av[j].v[e], PRED == 2 ? pv[j].v[e] : PredT(0), a, e & 1);
    }
    for (; u < units; u += G) {
        const Vec<AggT, W> av = ldg_stream<W>(agg + u * W);
        Vec<PredT, W> pv;
        if constexpr (PRED == 2) pv = ldg_stream<W>(pred + u * W);
#pragma unroll
        for (int e = 0; e < W; ++e)
            scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, av.v[e], PRED == 2 ? pv.v[e] : PredT(0), a, e & 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the n % W tail
        for (uint64_t i = units * W; i < a.n; ++i)
            scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, agg[i], PRED == 2 ? pred[i] : PredT(0), a, 0);
    }
    dd_merge(acc.sum, alt);
    acc = scan_block_reduce<IS_INT, MOMENTS>(acc, sm);
    scan_finish<IS_INT, MOMENTS>(acc, a, sm);
}

// Scalar-load fallback for columns whose base pointer is not vector aligned (attached torch slices).
template <typename AggT, int PRED, typename PredT, bool MOMENTS>
__global__ void __launch_bounds__(kScanThreads) k_scan_unaligned(const ScanArgs a) {
    constexpr bool IS_INT = std::is_integral_v<AggT>;
    __shared__ ScanAcc sm[32];
    const AggT* __restrict__ agg = static_cast<const AggT*>(a.agg);
    const PredT* __restrict__ pred = static_cast<const PredT*>(a.pred);
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    ScanAcc acc = scan_identity();
    DD alt{0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += G)
        scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, __ldg(agg + i), PRED == 2 ? __ldg(pred + i) : PredT(0), a, 0);
    acc = scan_block_reduce<IS_INT, MOMENTS>(acc, sm);
    scan_finish<IS_INT, MOMENTS>(acc, a, sm);
}

// K1/K2, TMA-staged ring (the default scan kernel for every column-type combination): one producer warp streams
// tiles of the aggregate column (and, for a predicate on another column, the matching tile of that column) into a
// STAGES-deep shared-memory ring with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers;
// 8 consumer warps reduce tiles out of shared memory with conflict-free 64/128-bit LDS.  Bytes in flight are set
// by the ring (STAGES x 16 KiB x CTAs/SM), not by how ptxas schedules loads.  Tile c goes to CTA (c mod gridDim.x):
// static, deterministic.  Tiles hold a power-of-two number of rows so that <= 16 KiB per stage is used.
constexpr int kBulkConsumerWarps = 8;
constexpr int kBulkThreads = (kBulkConsumerWarps + 1) * 32;
constexpr int kStageBytes = 16384;

template <typename AggT, int PRED, typename PredT> struct RingGeom {
    static constexpr int kRowBytes = (int)sizeof(AggT) + (PRED == 2 ? (int)sizeof(PredT) : 0);
    static constexpr int kRows = kRowBytes <= 4 ? 4096 : (kRowBytes <= 8 ? 2048 : 1024);   // rows per tile
    static constexpr int kAggBytes = kRows * (int)sizeof(AggT);
    static constexpr int kPredBytes = PRED == 2 ? kRows * (int)sizeof(PredT) : 0;
    static_assert(kAggBytes + kPredBytes <= kStageBytes, "tile exceeds the stage");
    // rows handled per consumer-thread step: 2 if any column is 8 bytes wide (LDS.128 / LDS.64), else 4 (LDS.128)
    static constexpr int kUnit = (sizeof(AggT) == 8 || (PRED == 2 && sizeof(PredT) == 8)) ? 2 : 4;
};

template <typename T, int N> __device__ __forceinline__ Vec<T, N> lds_vec(const unsigned char* base, uint32_t unit_index) {
    Vec<T, N> r;
    if constexpr (sizeof(T) * N == 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(base + (size_t)unit_index * 16);
        *reinterpret_cast<uint4*>(&r) = q;
    } else {
        static_assert(sizeof(T) * N == 8, "8- or 16-byte units");
        const uint2 q = *reinterpret_cast<const uint2*>(base + (size_t)unit_index * 8);
        *reinterpret_cast<uint2*>(&r) = q;
    }
    return r;
}

template <typename AggT, int PRED, typename PredT, int STAGES, bool MOMENTS>
__global__ void __launch_bounds__(kBulkThreads) k_scan_ring(const ScanArgs a) {
    using G = RingGeom<AggT, PRED, PredT>;
    constexpr bool IS_INT = std::is_integral_v<AggT>;
    constexpr int U = G::kUnit;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ ScanAcc sm[32];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    const unsigned char* __restrict__ agg = static_cast<const unsigned char*>(a.agg);
    const unsigned char* __restrict__ pred = static_cast<const unsigned char*>(a.pred);
    const uint64_t n_main = a.n & ~3ull;  // bulk copies move multiples of 16 bytes: 4 rows of the narrowest type
    const uint64_t ntiles = (n_main + G::kRows - 1) / G::kRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kBulkConsumerWarps); }
        fence_barrier_init();
    }
    __syncthreads();
    if (a.pdl_tail) griddep_launch_dependents();

    ScanAcc acc = scan_identity();
    DD alt{0.0, 0.0};
    if (warp == kBulkConsumerWarps) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t c; (c = scan_tile_of(a, it, blockIdx.x, gridDim.x, ntiles)) < ntiles; ++it) {
                const int s = it % STAGES;
                const uint32_t round = it / STAGES;
                if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                const uint64_t row0 = c * (uint64_t)G::kRows;
                const uint32_t rows = (uint32_t)((n_main - row0) < (uint64_t)G::kRows ? (n_main - row0) : (uint64_t)G::kRows);
                unsigned char* stage = smem_raw + (size_t)s * kStageBytes;
                mbar_expect_tx(&full_bar[s], rows * (uint32_t)G::kRowBytes);
                bulk_g2s(stage, agg + row0 * sizeof(AggT), rows * (uint32_t)sizeof(AggT), &full_bar[s]);
                if constexpr (PRED == 2) bulk_g2s(stage + G::kAggBytes, pred + row0 * sizeof(PredT), rows * (uint32_t)sizeof(PredT), &full_bar[s]);
            }
        }
    } else {
        const uint32_t ct = threadIdx.x;  // 0 .. 255
        uint32_t it = 0;
        for (uint64_t c; (c = scan_tile_of(a, it, blockIdx.x, gridDim.x, ntiles)) < ntiles; ++it) {
            const int s = it % STAGES;
            const uint32_t round = it / STAGES;
            mbar_wait(&full_bar[s], round & 1);
            const uint64_t row0 = c * (uint64_t)G::kRows;
            const uint32_t rows = (uint32_t)((n_main - row0) < (uint64_t)G::kRows ? (n_main - row0) : (uint64_t)G::kRows);
            const unsigned char* stage = smem_raw + (size_t)s * kStageBytes;
            const uint32_t nunits = rows / U;
            if constexpr (!IS_INT) {
                // f64 sums: the (at most 8) values a thread takes from one tile are added in plain double -- two independent chains --
                // and only the tile's subtotal enters the compensated (TwoSum) accumulator: 1 + 7/8 FP64 adds per value instead of 7.
                // The FP64 pipe was 48 % busy with the per-value TwoSum and held the board at its 1 kW power cap in sustained loops
                // (SM clock 1.71 GHz, 5.6 TB/s; profiles/r2_bench_n1_a.json).  Error: a subtotal of <= 8 same-sign values is off by
                // <= 3 ulp of ITSELF, the compensated sum of the subtotals adds nothing to that: <= 2 ulp of the total in the worst case,
                // ~1e-20 relative in practice (the subtotals' errors are independent); the reference's serial sum is off by ~5e-13 at 1 B rows.
                double b0 = 0.0, b1 = 0.0;
                uint32_t passed = 0;
#pragma unroll 4
                for (uint32_t i = ct; i < nunits; i += kBulkConsumerWarps * 32) {
                    const Vec<AggT, U> av = lds_vec<AggT, U>(stage, i);
                    Vec<PredT, U> pv;
                    if constexpr (PRED == 2) pv = lds_vec<PredT, U>(stage + G::kAggBytes, i);
#pragma unroll
                    for (int e = 0; e < U; ++e) {
                        bool pass = true;
                        if constexpr (PRED == 1) pass = (av.v[e] >= a.lo) && (av.v[e] <= a.hi);
                        if constexpr (PRED == 2) {
                            if constexpr (std::is_integral_v<PredT>) pass = ((long long)pv.v[e] >= a.ilo) && ((long long)pv.v[e] <= a.ihi);
                            else { const double d = as_f64(pv.v[e]); pass = (d >= a.lo) && (d <= a.hi); }
                        }
                        passed += pass ? 1u : 0u;
                        const double x = pass ? (double)av.v[e] : 0.0;
                        if (e & 1) b1 = __dadd_rn(b1, x); else b0 = __dadd_rn(b0, x);
                        if constexpr (MOMENTS) {   // the synchronous API also reports sum of squares / min / max; count, sum, comp are the same bits
                            acc.sq.s = fma(x, x, acc.sq.s);
                            acc.mn = fmin(acc.mn, pass ? (double)av.v[e] : acc.mn);
                            acc.mx = fmax(acc.mx, pass ? (double)av.v[e] : acc.mx);
                        }
                    }
                }
                acc.count += passed;
                dd_add(acc.sum, __dadd_rn(b0, b1));
            } else {
#pragma unroll 4
                for (uint32_t i = ct; i < nunits; i += kBulkConsumerWarps * 32) {
                    const Vec<AggT, U> av = lds_vec<AggT, U>(stage, i);
                    Vec<PredT, U> pv;
                    if constexpr (PRED == 2) pv = lds_vec<PredT, U>(stage + G::kAggBytes, i);
#pragma unroll
                    for (int e = 0; e < U; ++e)
                        scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, av.v[e], PRED == 2 ? pv.v[e] : PredT(0), a, e & 1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {  // the n % 4 tail
            const AggT* ag = static_cast<const AggT*>(a.agg);
            const PredT* pr = static_cast<const PredT*>(a.pred);
            for (uint64_t i = n_main; i < a.n; ++i)
                scan_consume<AggT, PRED, PredT, MOMENTS>(acc, alt, ag[i], PRED == 2 ? pr[i] : PredT(0), a, 0);
        }
    }
    dd_merge(acc.sum, alt);
    acc = scan_block_reduce<IS_INT, MOMENTS>(acc, sm);
    scan_finish<IS_INT, MOMENTS>(acc, a, sm);
}

// ================================================================================================
// Sample plans on the device
// ================================================================================================
struct PlanDev {
    const aqe_segment* segs;    // nseg affine runs, or
    const uint64_t* seg_start;  // exclusive prefix of counts, nseg + 1 entries
    const int64_t* idx;         // explicit positions (nseg == 0)
    const int64_t* perm;        // optional: positions index this permutation (amount-sorted order)
    uint32_t nseg;
    uint64_t count;
};

__device__ __forceinline__ int64_t plan_position(const PlanDev& P, uint64_t k) {
    int64_t pos;
    if (P.nseg == 0) {
        pos = __ldg(P.idx + k);
    } else {
        uint32_t lo = 0, hi = P.nseg;  // find seg with seg_start[seg] <= k < seg_start[seg+1]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(P.seg_start + mid) <= k) lo = mid; else hi = mid;
        }
        const aqe_segment s = P.segs[lo];
        const uint64_t r = k - __ldg(P.seg_start + lo);
        if (s.kind == 1) pos = (int64_t)(uint64_t)__dmul_rn((double)r, s.scale);
        else if (s.kind == 2) pos = (int64_t)feistel_perm(r, (uint64_t)s.base, (uint64_t)s.outer_step, (uint32_t)s.inner_len);
        else if (s.kind == 3) {  // jittered stride: (r*stride + U{0..stride/2}) mod M   (address_arithmetic_sample)
            const uint64_t stride = (uint64_t)s.outer_step, M = (uint64_t)s.base;
            const u32x4 q = philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), kSeedStream, (uint32_t)AQE_M_ADDRESS_ARITHMETIC, (uint32_t)s.inner_len,
                                          (uint32_t)((uint64_t)s.inner_len >> 32));
            pos = (int64_t)((r * stride + mulhi64(((uint64_t)q.y << 32) | q.x, stride / 2 + 1)) % M);
        }
        else if (s.inner_len == 1) pos = s.base + (int64_t)r * s.outer_step;
        else pos = s.base + (int64_t)(r / (uint64_t)s.inner_len) * s.outer_step + (int64_t)(r % (uint64_t)s.inner_len);
    }
    if (P.perm) pos = __ldg(P.perm + pos);
    return pos;
}

// An f64 column of a table whose row ranges live in several allocations (the shards of a range-sharded table, mapped into
// this GPU's address space by peer access): row i lives in the part p with first[p] <= i < first[p + 1].
struct GlobalF64 {
    const double* base[kMaxRanks];
    uint64_t first[kMaxRanks + 1];
    int parts;
    __device__ __forceinline__ double at(uint64_t i) const {
        int p = 0;
        while (p + 1 < parts && i >= first[p + 1]) ++p;
        return __ldg(base[p] + (i - first[p]));
    }
};

struct Columns {
    const int64_t* id;
    const double* amount;
    const int32_t* region;
    const int32_t* product_id;
    const int64_t* timestamp;
};
__device__ __forceinline__ double column_value(const Columns& C, int col, int64_t i) {
    switch (col) {
        case AQE_COL_ID: return (double)__ldg(C.id + i);
        case AQE_COL_AMOUNT: return __ldg(C.amount + i);
        case AQE_COL_REGION: return (double)__ldg(C.region + i);
        case AQE_COL_PRODUCT_ID: return (double)__ldg(C.product_id + i);
        default: return (double)__ldg(C.timestamp + i);
    }
}

// ================================================================================================
// K3/K5: moments of a column over a plan.  Sums are taken of d = x - K with K = the first sampled value,
// so M2 = sum d^2 - (sum d)^2 / n has no catastrophic cancellation; sum x is carried separately
// (compensated) because the CLI estimator is sum * N/n (enhanced_aqe_cli.py:190-193).
// ================================================================================================
struct StatAcc { uint64_t n; DD sx, sd, sdd; };
__device__ __forceinline__ StatAcc stat_identity() { return StatAcc{0, DD{0, 0}, DD{0, 0}, DD{0, 0}}; }
__device__ __forceinline__ void stat_merge(StatAcc& a, const StatAcc& b) {
    a.n += b.n; dd_merge(a.sx, b.sx); dd_merge(a.sd, b.sd); dd_merge(a.sdd, b.sdd);
}
__device__ __forceinline__ StatAcc stat_warp_reduce(StatAcc a) {
    a.n = warp_reduce_u64(a.n); a.sx = warp_reduce_dd(a.sx); a.sd = warp_reduce_dd(a.sd); a.sdd = warp_reduce_dd(a.sdd);
    return a;
}
__device__ __forceinline__ StatAcc stat_block_reduce(StatAcc a, StatAcc* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = stat_warp_reduce(a);
    __syncthreads();
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (warp == 0) { a = lane < nw ? sm[lane] : stat_identity(); a = stat_warp_reduce(a); }
    return a;
}

struct StatArgs {
    PlanDev plan;
    Columns cols;
    int col;
    int pred_col;   // AQE_COL_NONE, or rows failing lo <= pred_col <= hi contribute 0 (still counted in n)
    double lo, hi;
    StatAcc* partials;
    unsigned int* ticket;
    aqe_stats* out;
    // Range-sharded tables: the plan's positions are rows of the WHOLE table, `cols` hold rows [win_first, win_first + win_n)
    // of it; positions outside the window belong to another shard and are skipped here (every shard walks the same plan, the
    // gathers split).  The whole table on one GPU: win_first = 0, win_n = 2^64 - 1.
    uint64_t win_first, win_n;
    aqe_stats_partial* raw_out;  // windowed form: the shard's mergeable sums (aqe_stats_merge) instead of finished moments
};

// GU = independent gathers in flight per thread: 8 for scattered positions (stride / permutation / index-list plans), 4 for plans of
// contiguous tiles (coalesced already; the 16 extra registers of GU = 8 cost them occupancy: 44.9 -> 63.8 us under ncu, 400 M rows).
template <int GU> __global__ void __launch_bounds__(256) k_plan_stats(const StatArgs a) {
    __shared__ StatAcc sm[32];
    __shared__ bool is_last;
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    StatAcc acc = stat_identity();
    const bool windowed = a.raw_out != nullptr;
    auto local_value = [&](int64_t row) {
        double x = column_value(a.cols, a.col, row);
        if (a.pred_col != AQE_COL_NONE) {
            const double pv = column_value(a.cols, a.pred_col, row);
            x = (pv >= a.lo && pv <= a.hi) ? x : 0.0;
        }
        return x;
    };
    // in[j]: position k is a row of this shard; x = its value (0 with in = false)
    auto value_at = [&](uint64_t k, bool& in) {
        const uint64_t u = (uint64_t)plan_position(a.plan, k) - a.win_first;
        in = u < a.win_n;
        return in ? local_value((int64_t)u) : 0.0;
    };
    // shift K: the plan's first sampled value; a shard of a larger table takes its own first row (any value of the data
    // does: K only keeps sum d^2 - (sum d)^2 / n free of cancellation, the merge is exact about differing shifts)
    double K = 0.0;
    if (windowed) { if (a.win_n) K = column_value(a.cols, a.col, 0); }
    else if (a.plan.count) { bool in; K = value_at(0, in); }
    // 8 independent gathers in flight per thread.  A random 8-byte read costs a whole 128-byte line of L2 / DRAM traffic on B200
    // (tools/microbench.cu under ncu: 4 sectors per load whatever the load form or the L2 fetch-granularity limit), so the ceiling
    // is HBM bandwidth / 128 B, 48-56 G samples/s, reached only with many loads in flight (profiles/r2_mb_gather.jsonl)
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; k + (uint64_t)(GU - 1) * G < a.plan.count; k += (uint64_t)GU * G) {
        double x[GU];
        bool in[GU];
#pragma unroll
        for (int j = 0; j < GU; ++j) x[j] = value_at(k + (uint64_t)j * G, in[j]);
#pragma unroll
        for (int j = 0; j < GU; ++j) {
            if (!in[j]) continue;
            const double d = __dadd_rn(x[j], -K);
            acc.n += 1; dd_add(acc.sx, x[j]); dd_add(acc.sd, d); dd_add(acc.sdd, __dmul_rn(d, d));
        }
    }
    for (; k < a.plan.count; k += G) {
        bool in;
        const double x = value_at(k, in);
        if (!in) continue;
        const double d = __dadd_rn(x, -K);
        acc.n += 1; dd_add(acc.sx, x); dd_add(acc.sd, d); dd_add(acc.sdd, __dmul_rn(d, d));
    }
    acc = stat_block_reduce(acc, sm);
    if (threadIdx.x == 0) {
        a.partials[blockIdx.x] = acc;
        __threadfence();
        is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    acc = stat_identity();
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) stat_merge(acc, load_cg(a.partials + b));
    acc = stat_block_reduce(acc, sm);
    if (threadIdx.x == 0 && windowed) {
        dd_norm(acc.sx); dd_norm(acc.sd); dd_norm(acc.sdd);
        aqe_stats_partial r;
        r.n = acc.n; r.sum = acc.sx.s; r.sum_c = acc.sx.c; r.shift = K;
        r.sd = acc.sd.s; r.sd_c = acc.sd.c; r.sdd = acc.sdd.s; r.sdd_c = acc.sdd.c;
        *a.raw_out = r;
        *a.ticket = 0u;
    } else if (threadIdx.x == 0) {
        dd_norm(acc.sx); dd_norm(acc.sd); dd_norm(acc.sdd);
        aqe_stats r;
        r.n = acc.n; r.sum = acc.sx.s;
        const double n = (double)acc.n;
        r.mean = acc.n ? acc.sx.s / n : 0.0;
        double m2 = acc.n ? acc.sdd.s - (acc.sd.s * acc.sd.s) / n : 0.0;
        r.m2 = m2 > 0.0 ? m2 : 0.0;
        *a.out = r;
        *a.ticket = 0u;
    }
}

// K6: rows at plan positions, AoS, in plan order.
// Sharded tables (win_*: see StatArgs): only positions inside this shard's window are written; every shard of the table writes its
// own rows into the SAME output buffer (plan order), so the shards together fill it.
__global__ void __launch_bounds__(256) k_plan_gather(const PlanDev plan, const Columns C, aqe_record* __restrict__ out, uint64_t first, uint64_t n,
                                                     uint64_t win_first, uint64_t win_n, unsigned long long* counter) {
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long wrote = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += G) {
        const uint64_t u = (uint64_t)plan_position(plan, first + k) - win_first;
        if (u >= win_n) continue;
        ++wrote;
        const int64_t i = (int64_t)u;
        aqe_record r;
        r.id = C.id ? __ldg(C.id + i) : 0;
        r.amount = C.amount ? __ldg(C.amount + i) : 0.0;
        r.region = C.region ? __ldg(C.region + i) : 0;
        r.product_id = C.product_id ? __ldg(C.product_id + i) : 0;
        r.timestamp = C.timestamp ? __ldg(C.timestamp + i) : 0;
        uint4* o = reinterpret_cast<uint4*>(out + k);
        const uint4* s = reinterpret_cast<const uint4*>(&r);
        o[0] = s[0]; o[1] = s[1];
    }
    if (counter) {
        wrote = warp_reduce_u64(wrote);
        if ((threadIdx.x & 31) == 0 && wrote) atomicAdd(counter, wrote);
    }
}

// contiguous rows [first, first+n) -> AoS (collect_all_records :660, save_to_file :665)
__global__ void __launch_bounds__(256) k_soa_to_aos(const Columns C, aqe_record* __restrict__ out, uint64_t first, uint64_t n) {
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += G) {
        const uint64_t i = first + k;
        aqe_record r;
        r.id = C.id ? C.id[i] : 0;
        r.amount = C.amount ? C.amount[i] : 0.0;
        r.region = C.region ? C.region[i] : 0;
        r.product_id = C.product_id ? C.product_id[i] : 0;
        r.timestamp = C.timestamp ? C.timestamp[i] : 0;
        uint4* o = reinterpret_cast<uint4*>(out + k);
        const uint4* s = reinterpret_cast<const uint4*>(&r);
        o[0] = s[0]; o[1] = s[1];
    }
}

// K7: ingest.  A warp reads 32 rows = 1 KiB contiguous; column stores are 256/128-byte coalesced.
// `unsorted` is raised if ids are not strictly ascending: bit 0 for an id below its predecessor (the host re-orders,
// custom_bplus_db.cpp:198-200), bit 1 for an id equal to it (duplicate ids: the host works out where the reference's tree
// would have put them, aqe_order.cpp).
struct MutColumns { int64_t* id; double* amount; int32_t* region; int32_t* product_id; int64_t* timestamp; };
__global__ void __launch_bounds__(256) k_aos_to_soa(const aqe_record* __restrict__ rows, uint64_t n, MutColumns C, uint64_t dst,
                                                    int64_t prev_last_id, int has_prev, unsigned int* unsorted) {
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += G) {
        const uint4* s = reinterpret_cast<const uint4*>(rows + k);
        const uint4 a = s[0], b = s[1];
        aqe_record r;
        reinterpret_cast<uint4*>(&r)[0] = a; reinterpret_cast<uint4*>(&r)[1] = b;
        C.id[dst + k] = r.id; C.amount[dst + k] = r.amount; C.region[dst + k] = r.region;
        C.product_id[dst + k] = r.product_id; C.timestamp[dst + k] = r.timestamp;
        if (k || has_prev) {
            const int64_t prev = k ? rows[k - 1].id : prev_last_id;
            if (r.id <= prev) atomicOr(unsorted, r.id < prev ? 1u : 2u);
        }
    }
}

__global__ void __launch_bounds__(256) k_synth(uint64_t seed, uint64_t first_row, uint64_t n, int dist, MutColumns C) {
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += G) {
        aqe_record r;
        synth_row(seed, first_row + k, dist, r);
        if (C.id) C.id[k] = r.id;
        if (C.amount) C.amount[k] = r.amount;
        if (C.region) C.region[k] = r.region;
        if (C.product_id) C.product_id[k] = r.product_id;
        if (C.timestamp) C.timestamp[k] = r.timestamp;
    }
}

// ================================================================================================
// K4: persistent CLT estimator.  One cooperative launch; rounds ("looks") at cumulative sample sizes
// n_0 < n_1 < ...; sample j is the j-th element of a Philox stream, so the sample set after a look is a
// prefix of one deterministic sequence (seed, design).  Per look: every block reduces its share to
// (units, rows, pass count, sum d, sum d^2) with d = y - K, writes it to a double-buffered slot, one grid
// barrier, then EVERY block folds all slots in block order (identical arithmetic -> identical decision in
// every block, no broadcast needed) and evaluates the stop rule  z * sqrt(var / n) * scale / |estimate| <= eps.
// Replaces the std::async fast/slow thread pool + should_stop atomics of custom_bplus_db.cpp:885-1043.
// ================================================================================================
struct ApproxAcc { uint64_t units, rows; DD sc, sd, sdd; };
__device__ __forceinline__ ApproxAcc approx_identity() { return ApproxAcc{0, 0, DD{0, 0}, DD{0, 0}, DD{0, 0}}; }
__device__ __forceinline__ void approx_merge(ApproxAcc& a, const ApproxAcc& b) {
    a.units += b.units; a.rows += b.rows; dd_merge(a.sc, b.sc); dd_merge(a.sd, b.sd); dd_merge(a.sdd, b.sdd);
}
__device__ __forceinline__ ApproxAcc approx_warp_reduce(ApproxAcc a) {
    a.units = warp_reduce_u64(a.units); a.rows = warp_reduce_u64(a.rows);
    a.sc = warp_reduce_dd(a.sc); a.sd = warp_reduce_dd(a.sd); a.sdd = warp_reduce_dd(a.sdd);
    return a;
}
__device__ __forceinline__ ApproxAcc approx_block_reduce(ApproxAcc a, ApproxAcc* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = approx_warp_reduce(a);
    __syncthreads();
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (warp == 0) { a = lane < nw ? sm[lane] : approx_identity(); a = approx_warp_reduce(a); }
    return a;
}

struct ApproxArgs {
    Columns cols;
    uint64_t n_rows, units;
    uint32_t block_rows;  // 1 for SRS
    int design, agg, agg_col, pred_col;
    double lo, hi, eps, z;   // z: normal quantile at the (guarded) confidence level; the kernel turns it into t(df)
    int stein;               // 1: the half width at look r uses the variance of look r-1 (include/aqe_b200.h, aqe_ci_mode)
    uint64_t seed, n0, nmax;
    ApproxAcc* slots;  // [2][gridDim.x]
    aqe_approx_result* out;
    // multi-GPU (shards are strata, one GLOBAL stop rule): ex.world > 1.  ex.seq = index of this query's first message.
    Exchange ex;
    uint64_t units_total, rows_total, n0_total, nmax_total;
};

// What a rank publishes after every look: its cumulative sample moments (about its own shift K) and its stratum size.
struct ApproxMsg { unsigned long long n_units, n_rows; double sc, sd, sdd, K; unsigned long long pop_units, pop_rows; };
static_assert(sizeof(ApproxMsg) == sizeof(aqe_partial), "messages travel in 64-byte mailbox payloads");

// rank's share of a global cumulative sample size T (proportional allocation)
__host__ __device__ __forceinline__ uint64_t approx_share(uint64_t T, uint64_t units_g, uint64_t units_total) {
    if (units_g == 0) return 0;
    const double x = ceil((double)T * ((double)units_g / (double)units_total));
    const uint64_t v = x >= (double)units_g ? units_g : (uint64_t)x;
    return v < 1 ? 1 : v;
}

// Stratified estimate over all ranks' messages (rank order).  Returns the relative half width in percent.
// SUM/COUNT: T = sum_g U_g mean_g, Var = sum_g U_g^2 s_g^2 / n_g.  AVG: the same over the total row count.  AVG with a
// predicate: ratio estimator R = Num/Den with the linearised residual e = y - R c in every stratum.
// var_io / nvar_io (per stratum; *have_prev says whether they hold the previous look): with `stein` the half width uses the
// PREVIOUS look's stratum variances (and the t quantile at their degrees of freedom) while means and n are the current ones;
// on return they hold the current look's.  *rel_next = the relative half width with the current variances: sizes the next look.
__host__ __device__ inline double approx_global(const ApproxMsg* msgs, int world, int agg, bool ratio, double z, int stein, double* var_io, uint64_t* nvar_io,
                                                bool* have_prev, double* est_out, double* half_out, uint64_t* n_tot_out, uint64_t* rows_tot_out, double* rel_next,
                                                double* pass_fraction) {
    const double inf = HUGE_VAL;
    double T = 0.0, V = 0.0, Vcur = 0.0, num = 0.0, den = 0.0, pop_rows = 0.0, df_use = 0.0, df_cur = 0.0;
    uint64_t n_tot = 0, rows_tot = 0;
    for (int g = 0; g < world; ++g) {
        const ApproxMsg& m = msgs[g];
        n_tot += m.n_units; rows_tot += m.n_rows; pop_rows += (double)m.pop_rows;
        if (m.pop_units == 0 || m.n_units == 0) continue;
        const double n = (double)m.n_units, U = (double)m.pop_units;
        const double sy = m.sd + n * m.K;
        num += U * (sy / n); den += U * (m.sc / n);
    }
    const double R = den > 0.0 ? num / den : 0.0;
    const bool use_prev = stein && *have_prev;
    for (int g = 0; g < world; ++g) {
        const ApproxMsg& m = msgs[g];
        if (m.pop_units == 0 || m.n_units == 0) continue;
        const double n = (double)m.n_units, U = (double)m.pop_units;
        const double sy = m.sd + n * m.K;
        const double mu = sy / n;
        double ss = m.sdd - (m.sd * m.sd) / n;
        if (ss < 0.0) ss = 0.0;
        double var;
        if (ratio) {
            const double syy = ss + sy * mu;
            double se2 = syy - 2.0 * R * sy + R * R * m.sc;
            const double ebar = (sy - R * m.sc) / n;
            se2 -= n * ebar * ebar;
            if (se2 < 0.0) se2 = 0.0;
            var = m.n_units > 1 ? se2 / (n - 1.0) : inf;
        } else {
            var = m.n_units > 1 ? ss / (n - 1.0) : inf;
        }
        const bool take_prev = use_prev && nvar_io[g] > 1 && var_io[g] > var;   // the larger of the two variances (see k_approx)
        const double v_use = take_prev ? var_io[g] : var;
        df_use += take_prev ? (double)(nvar_io[g] - 1) : n - 1.0;
        df_cur += n - 1.0;
        T += U * mu;
        V += U * U * v_use / n;
        Vcur += U * U * var / n;
        var_io[g] = var; nvar_io[g] = m.n_units;
    }
    *have_prev = true;
    const double tq = t_from_z(z, df_use > 4.0 ? df_use : 4.0), tq_next = t_from_z(z, df_cur > 4.0 ? df_cur : 4.0);
    double est, half, half_next;
    if (ratio) { est = R; half = den > 0.0 ? tq * sqrt(V) / den : inf; half_next = den > 0.0 ? tq_next * sqrt(Vcur) / den : inf; }
    else if (agg == AQE_AGG_AVG) { est = pop_rows > 0.0 ? T / pop_rows : 0.0; half = pop_rows > 0.0 ? tq * sqrt(V) / pop_rows : inf; half_next = pop_rows > 0.0 ? tq_next * sqrt(Vcur) / pop_rows : inf; }
    else { est = T; half = tq * sqrt(V); half_next = tq_next * sqrt(Vcur); }
    *est_out = est; *half_out = half; *n_tot_out = n_tot; *rows_tot_out = rows_tot;
    *pass_fraction = pop_rows > 0.0 ? den / pop_rows : 0.0;
    if (ratio && !(den > 0.0)) { *rel_next = inf; return inf; }
    *rel_next = est != 0.0 ? half_next / fabs(est) * 100.0 : inf;
    return est != 0.0 ? half / fabs(est) * 100.0 : inf;
}

__device__ __forceinline__ void approx_row(const ApproxArgs& a, uint64_t row, double& y, double& c) {
    bool pass = true;
    if (a.pred_col != AQE_COL_NONE) { const double pv = column_value(a.cols, a.pred_col, (int64_t)row); pass = (pv >= a.lo) && (pv <= a.hi); }
    c = pass ? 1.0 : 0.0;
    y = pass ? (a.agg == AQE_AGG_COUNT ? 1.0 : column_value(a.cols, a.agg_col, (int64_t)row)) : 0.0;
}

__global__ void __launch_bounds__(256) k_approx(const ApproxArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ ApproxAcc sm[32];
    __shared__ ApproxAcc sh_total;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool ratio = (a.agg == AQE_AGG_AVG && a.pred_col != AQE_COL_NONE);

    // shift K: value of the first row of the first drawn unit, times rows per unit
    double K = 0.0;
    if (a.units) {
        const uint64_t u0 = draw_position(a.seed, (uint32_t)a.design, 0, a.units);
        double y0, c0; approx_row(a, u0 * a.block_rows, y0, c0);
        K = y0 * (double)a.block_rows;
    }

    const bool multi = a.ex.world > 1;
    __shared__ ApproxMsg sh_msgs[kMaxRanks];
    uint64_t Tg = a.n0_total, n_tot = 0, rows_tot = 0;  // global cumulative look size (multi)
    ApproxAcc cum = approx_identity();
    uint64_t n_prev = 0, target = multi ? approx_share(Tg, a.units, a.units_total) : a.n0;
    uint32_t rounds = 0;
    int status = AQE_DRIFTING;
    double est = 0.0, half = 0.0, rel = 0.0, rel_next = 0.0, mean = 0.0, m2 = 0.0, pass_fraction = 1.0;
    // the previous look's variance(s): what the Stein-type interval is built from (include/aqe_b200.h, aqe_ci_mode)
    double var_prev[kMaxRanks];
    uint64_t nvar_prev[kMaxRanks];
    bool have_prev = false;
#pragma unroll
    for (int g = 0; g < kMaxRanks; ++g) { var_prev[g] = 0.0; nvar_prev[g] = 0; }

    for (;;) {
        ApproxAcc acc = approx_identity();
        if (a.design == AQE_DESIGN_SRS) {
            // a thread owns Philox counter q: samples 2q and 2q+1
            for (uint64_t q = (n_prev >> 1) + gtid; q < ((target + 1) >> 1); q += gthreads) {
                const u32x4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), kDrawStream | (uint32_t)a.design, 0u,
                                              (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                const uint64_t j0 = 2 * q, j1 = 2 * q + 1;
                const uint64_t p0 = mulhi64(((uint64_t)r.y << 32) | r.x, a.units);
                const uint64_t p1 = mulhi64(((uint64_t)r.w << 32) | r.z, a.units);
                double y0 = 0, c0 = 0, y1 = 0, c1 = 0;
                const bool v0 = j0 >= n_prev && j0 < target, v1 = j1 >= n_prev && j1 < target;
                if (v0) approx_row(a, p0, y0, c0);
                if (v1) approx_row(a, p1, y1, c1);
                if (v0) { const double d = __dadd_rn(y0, -K); acc.units++; acc.rows++; dd_add(acc.sc, c0); dd_add(acc.sd, d); dd_add(acc.sdd, __dmul_rn(d, d)); }
                if (v1) { const double d = __dadd_rn(y1, -K); acc.units++; acc.rows++; dd_add(acc.sc, c1); dd_add(acc.sd, d); dd_add(acc.sdd, __dmul_rn(d, d)); }
            }
        } else {
            // a warp owns a tile: lanes stride the tile's rows (coalesced), fixed-order warp fold
            const uint64_t gwarps = gthreads >> 5, gwarp = gtid >> 5;
            for (uint64_t j = n_prev + gwarp; j < target; j += gwarps) {
                const uint64_t u = draw_position(a.seed, (uint32_t)a.design, j, a.units);
                const uint64_t r0 = u * a.block_rows;
                const uint64_t r1 = (r0 + a.block_rows < a.n_rows) ? r0 + a.block_rows : a.n_rows;
                DD ty{0, 0}, tc{0, 0};
                for (uint64_t r = r0 + lane; r < r1; r += 32) { double y, c; approx_row(a, r, y, c); dd_add(ty, y); dd_add(tc, c); }
                ty = warp_reduce_dd(ty); tc = warp_reduce_dd(tc);
                if (lane == 0) {
                    dd_norm(ty); dd_norm(tc);
                    const double d = __dadd_rn(ty.s, -K);
                    acc.units++; acc.rows += (r1 - r0); dd_add(acc.sc, tc.s); dd_add(acc.sd, d); dd_add(acc.sdd, __dmul_rn(d, d));
                }
            }
        }
        acc = approx_block_reduce(acc, sm);
        ApproxAcc* slot = a.slots + (size_t)(rounds & 1) * gridDim.x;
        if (threadIdx.x == 0) slot[blockIdx.x] = acc;
        grid.sync();
        // every block folds all slots in block order
        ApproxAcc t = approx_identity();
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) approx_merge(t, load_cg(slot + b));
        t = approx_block_reduce(t, sm);
        if (threadIdx.x == 0) sh_total = t;
        __syncthreads();
        approx_merge(cum, sh_total);
        ++rounds;

        if (multi) {
            // ---- exchange the per-rank cumulative moments through the mailboxes, then ONE global stop rule ----
            const int world = a.ex.world;
            const unsigned long long m = a.ex.seq + (rounds - 1), tag = m + 1;
            const int par = (int)(m & 1ull);
            if (blockIdx.x == 0 && (int)threadIdx.x < world) {
                ApproxAcc s = cum; dd_norm(s.sd); dd_norm(s.sdd); dd_norm(s.sc);
                ApproxMsg msg{s.units, s.rows, s.sc.s, s.sd.s, s.sdd.s, K, a.units, a.n_rows};
                ExSlot* dst = a.ex.peers[threadIdx.x] + (2 * kMaxRanks + a.ex.rank * 2 + par);  // second half of the mailbox
                const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&msg);
                volatile unsigned long long* d = reinterpret_cast<volatile unsigned long long*>(&dst->p);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = src[i];
                __threadfence_system();
                st_release_sys(&dst->seq, tag);
            }
            if ((int)threadIdx.x < world) {
                const ExSlot* src = a.ex.peers[a.ex.rank] + (2 * kMaxRanks + threadIdx.x * 2 + par);
                const long long t0 = clock64();
                while (ld_acquire_sys(&src->seq) != tag) {
                    if ((unsigned long long)(clock64() - t0) > a.ex.timeout_cycles) { *(volatile unsigned int*)a.ex.status = 1u; break; }
                    __nanosleep(32);
                }
                const volatile unsigned long long* sp = reinterpret_cast<const volatile unsigned long long*>(&src->p);
                unsigned long long* dp = reinterpret_cast<unsigned long long*>(&sh_msgs[threadIdx.x]);
#pragma unroll
                for (int i = 0; i < 8; ++i) dp[i] = sp[i];
            }
            __syncthreads();
            rel = approx_global(sh_msgs, world, a.agg, ratio, a.z, a.stein, var_prev, nvar_prev, &have_prev, &est, &half, &n_tot, &rows_tot, &rel_next, &pass_fraction);
            mean = 0.0; m2 = 0.0;
            n_prev = target;
            if (rel <= a.eps) { status = AQE_STABLE; break; }
            if (Tg >= a.nmax_total) { status = AQE_DRIFTING; break; }
            const double rn = rel_next / a.eps;
            const double want = ceil(1.1 * ((double)n_tot * rn * rn));
            const uint64_t lo_n = Tg + Tg / 4 + 1, hi_n = Tg * 8;
            uint64_t t2 = want >= (double)hi_n ? hi_n : (want <= (double)lo_n ? lo_n : (uint64_t)want);
            if (t2 > a.nmax_total) t2 = a.nmax_total;
            Tg = t2;
            const uint64_t nt = approx_share(Tg, a.units, a.units_total);
            target = nt > n_prev ? nt : n_prev;
            continue;
        }

        // ---- stop rule (same formulas as oracle/aqe_oracle.c orc_approx) ----
        ApproxAcc s = cum; dd_norm(s.sd); dd_norm(s.sdd); dd_norm(s.sc);
        const double n = (double)s.units;
        const double sy = s.sd.s + n * K;            // sum y
        const double mu = sy / n;
        double ss = s.sdd.s - (s.sd.s * s.sd.s) / n;  // sum (y - mean)^2
        if (ss < 0.0) ss = 0.0;
        mean = mu; m2 = ss;
        double var, scale;
        if (ratio) {
            if (s.sc.s <= 0.0) { est = 0.0; var = __longlong_as_double(0x7ff0000000000000LL); scale = 1.0; }
            else {
                const double R = sy / s.sc.s;
                // residual e = y - R c ; sum e^2 = sum y^2 - R sum y   (c^2 = c, y c = y)
                const double syy = ss + sy * mu;
                double rss = syy - R * sy; if (rss < 0.0) rss = 0.0;
                const double cbar = s.sc.s / n;
                var = rss / (n - 1.0) / (cbar * cbar);
                est = R; scale = 1.0;
            }
        } else {
            var = s.units > 1 ? ss / (n - 1.0) : __longlong_as_double(0x7ff0000000000000LL);
            if (a.agg == AQE_AGG_AVG) { scale = (double)a.units / (double)a.n_rows; est = mu * scale; }
            else { scale = (double)a.units; est = mu * scale; }
        }
        // the interval: the LARGER of the previous look's variance (the one that chose this look's size -- Stein) and the current one,
        // at the t quantile of its degrees of freedom; the first look has only its own.  (The previous variance alone under-covers
        // on heavy tails, where a 16 k-sample variance is still noisy: 0.937 on the lognormal column, profiles/r2_coverage_config4_4000seeds_stein_prev_only.json.)
        const bool use_prev = a.stein && have_prev && nvar_prev[0] > 1 && var_prev[0] > var;
        const double v_use = use_prev ? var_prev[0] : var;
        const double df_use = use_prev ? (double)(nvar_prev[0] - 1) : n - 1.0;
        half = t_from_z(a.z, df_use > 4.0 ? df_use : 4.0) * sqrt(v_use / n) * scale;
        const double half_next = t_from_z(a.z, n - 1.0 > 4.0 ? n - 1.0 : 4.0) * sqrt(var / n) * scale;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        rel = est != 0.0 ? half / fabs(est) * 100.0 : inf;
        rel_next = est != 0.0 ? half_next / fabs(est) * 100.0 : inf;
        var_prev[0] = var; nvar_prev[0] = s.units; have_prev = true;
        pass_fraction = s.rows ? s.sc.s / (double)s.rows : 0.0;
        n_prev = target;
        if (rel <= a.eps) { status = AQE_STABLE; break; }
        if (n_prev >= a.nmax) { status = AQE_DRIFTING; break; }
        const double rn = rel_next / a.eps;   // the next look is sized for the variance seen now: that is the one its interval will use
        const double want = ceil(1.1 * (n * rn * rn));
        const uint64_t lo_n = n_prev + n_prev / 4 + 1, hi_n = n_prev * 8;
        uint64_t t2 = want >= (double)hi_n ? hi_n : (want <= (double)lo_n ? lo_n : (uint64_t)want);
        if (t2 > a.nmax) t2 = a.nmax;
        target = t2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        aqe_approx_result r;
        r.estimate = est; r.ci_lower = est - half; r.ci_upper = est + half;
        r.error_margin = rel / 100.0; r.confidence_level = 0.0;
        r.n_samples = multi ? rows_tot : cum.rows; r.n_units = multi ? n_tot : cum.units; r.population = multi ? a.rows_total : a.n_rows;
        r.mean = mean; r.m2 = m2; r.rounds = rounds; r.status = status; r.elapsed_us = 0.0; r.pass_fraction = pass_fraction;
        *a.out = r;
    }
}

}  // namespace aqe
