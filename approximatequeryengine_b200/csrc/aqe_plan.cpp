// aqe_plan.cpp -- closed forms of the sampler position lists.
//
// The reference materialises every sample with a loop of `samples.push_back(all_records[i])` over a
// fresh O(N) copy of the table (custom_bplus_db.cpp:737-1960).  Here each sampler becomes a handful of
// affine segments {base, outer_step, inner_len, count} whose lengths are computed arithmetically, and the
// gather kernels expand them on the fly.  Samplers that have no affine form (mt19937 + std::set, the
// tree-shape walks, per-sample random offsets) produce an explicit list.  "cbd" below is
// src/aqe_backend/core/custom_bplus_db.cpp of the reference.
#include "aqe_plan.hpp"

#include <algorithm>
#include <cmath>
#include <random>

#include "aqe_device.cuh"

namespace aqe {

namespace {

inline int64_t target_of(uint64_t n, double p) { return (int64_t)((double)n * p / 100.0); }  // e.g. cbd:745
inline uint64_t cdiv(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline uint64_t umin(uint64_t a, uint64_t b) { return a < b ? a : b; }
inline uint64_t umax(uint64_t a, uint64_t b) { return a > b ? a : b; }

void add_seg(aqe_plan& pl, int64_t base, int64_t outer_step, int64_t inner_len, uint64_t count) {
    if (count == 0) return;
    aqe_segment s{};
    s.base = base; s.outer_step = outer_step; s.inner_len = inner_len; s.count = (int64_t)count; s.kind = 0; s.scale = 0.0;
    pl.segs.push_back(s);
}
void add_run(aqe_plan& pl, uint64_t first, uint64_t step, uint64_t count) { add_seg(pl, (int64_t)first, (int64_t)step, 1, count); }
void add_all(aqe_plan& pl, uint64_t n) { add_seg(pl, 0, 0, (int64_t)n, n); }
// number of terms of first, first+step, ... that are < limit
uint64_t run_len(uint64_t first, uint64_t step, uint64_t limit) { return first < limit ? cdiv(limit - first, step) : 0; }

// `budget` rows from the selected blocks j = j0 .. j1-1 whose row ranges are [j*stride_rows, +B) clipped
// to n; only the table's final block can be short and it can only be the last selected one.
uint64_t add_blocks(aqe_plan& pl, uint64_t n, uint64_t B, uint64_t j0, uint64_t j1, uint64_t stride_rows, uint64_t budget) {
    if (j1 <= j0 || budget == 0) return 0;
    const uint64_t nfull = j1 - j0 - 1;
    const uint64_t a = umin(budget, nfull * B);
    add_seg(pl, (int64_t)(j0 * stride_rows), (int64_t)stride_rows, (int64_t)B, a);
    const uint64_t last_base = (j1 - 1) * stride_rows;
    const uint64_t last_len = last_base < n ? umin(B, n - last_base) : 0;
    const uint64_t b = umin(budget - a, last_len);
    add_seg(pl, (int64_t)last_base, 0, (int64_t)umax(last_len, 1), b);
    return a + b;
}

}  // namespace

uint64_t cache_rows(uint64_t n) { return n >= 1000 ? n - n % 1000 : n; }
uint64_t leaf_count(uint64_t n) { return n < 255 ? 1 : (n - 255) / 127 + 2; }  // split 255 -> 127 | 128 (cbd:45-56, 215)
static uint64_t parent_count(uint64_t children) { return children <= 255 ? 1 : (children - 256) / 128 + 2; }  // cbd:63-70, 235
uint64_t tree_height(uint64_t n) {
    uint64_t h = 1, c = leaf_count(n);
    while (c > 1) { c = parent_count(c); ++h; }
    return h;
}

uint64_t seeded_u64(uint64_t seed, uint32_t method, uint64_t draw) {
    const u32x4 r = philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), kSeedStream, method, (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((uint64_t)r.y << 32) | r.x;
}
uint64_t seeded_below(uint64_t seed, uint32_t method, uint64_t draw, uint64_t bound) { return mulhi64(seeded_u64(seed, method, draw), bound); }

void plan_finalize(aqe_plan& pl) {
    pl.seg_start.clear();
    if (pl.segs.empty()) { pl.count = pl.idx.size(); return; }
    uint64_t acc = 0;
    for (const auto& s : pl.segs) { pl.seg_start.push_back(acc); acc += (uint64_t)s.count; }
    pl.seg_start.push_back(acc);
    pl.count = acc;
}

int64_t plan_position_host(const aqe_plan& pl, uint64_t k) {
    if (pl.segs.empty()) return pl.idx[k];
    size_t lo = 0, hi = pl.segs.size();
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (pl.seg_start[mid] <= k) lo = mid; else hi = mid; }
    const aqe_segment& s = pl.segs[lo];
    const uint64_t r = k - pl.seg_start[lo];
    if (s.kind == 1) { volatile double t = (double)r * s.scale; return (int64_t)(uint64_t)t; }
    if (s.kind == 2) return (int64_t)feistel_perm(r, (uint64_t)s.base, (uint64_t)s.outer_step, (uint32_t)s.inner_len);
    if (s.kind == 3) {
        const uint64_t stride = (uint64_t)s.outer_step, M = (uint64_t)s.base;
        return (int64_t)((r * stride + seeded_below((uint64_t)s.inner_len, (uint32_t)AQE_M_ADDRESS_ARITHMETIC, r, stride / 2 + 1)) % M);
    }
    if (s.inner_len == 1) return s.base + (int64_t)r * s.outer_step;
    return s.base + (int64_t)(r / (uint64_t)s.inner_len) * s.outer_step + (int64_t)(r % (uint64_t)s.inner_len);
}

int clt_threads(uint64_t N, const aqe_sample_params& p, std::vector<CltThread>& out, int64_t& T, std::string& err) {
    out.clear();
    T = target_of(N, p.sample_percent);
    if (N == 0 || T <= 0) return AQE_OK;
    const int64_t Th = p.num_threads, F = Th / 2, S = Th - F;
    if (F <= 0 || S <= 0 || p.check_interval < 2 || T / F == 0 || T / S == 0) {
        err = "clt_validated_dual_pointer_sample: needs num_threads >= 2, check_interval >= 2 and a target of at least "
              "one row per thread (the reference divides by zero, custom_bplus_db.cpp:927/936/981)";
        return AQE_ERR_INVALID;
    }
    for (int64_t q = 0; q < Th; ++q) {
        const bool fast = q < F;
        const uint64_t t = (uint64_t)(fast ? q : q - F), G = (uint64_t)(fast ? F : S);
        const uint64_t a = (N * t) / G, b = (N * (t + 1)) / G;                         // cbd:925-926 / 979-980
        const int64_t s = (int64_t)(int)((b - a) / (uint64_t)(T / (int64_t)G));
        CltThread c;
        c.fast = fast;
        c.step = (uint64_t)(fast ? std::max<int64_t>(3, s) : std::max<int64_t>(1, s));  // cbd:927 / 981
        c.first = fast ? a : a + c.step / 2;                                            // cbd:984
        c.len = run_len(c.first, c.step, b);
        out.push_back(c);
    }
    return AQE_OK;
}

int plan_build(uint64_t N, int method, const aqe_sample_params& P, const PlanData& data, aqe_plan& pl, std::string& err) {
    pl.segs.clear(); pl.idx.clear(); pl.by_amount_order = false;
    const double p = P.sample_percent;
    if (!(p == p)) { err = "sample_percent is NaN"; return AQE_ERR_INVALID; }
    const int64_t T = target_of(N, p);

    switch (method) {
        case AQE_M_SLOW_POINTER:   // cbd:759-778
        case AQE_M_FAST_POINTER: { // cbd:737-757
            if (N == 0 || T <= 0) break;
            const int64_t mult = method == AQE_M_FAST_POINTER ? P.step_size : 1;
            if (mult <= 0) { err = "fast_pointer_sample: step_size must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t step = umax(1, N / (uint64_t)T) * (uint64_t)mult;
            add_run(pl, 0, step, umin((uint64_t)T, run_len(0, step, N)));
            break;
        }
        case AQE_M_DUAL_POINTER: { // cbd:780-812
            if (N == 0 || T <= 0) break;
            const int64_t Tf = T / 3, Ts = T - Tf;
            if (Tf == 0) { err = "dual_pointer_sample: target < 3 rows (the reference divides by zero, custom_bplus_db.cpp:796)"; return AQE_ERR_INVALID; }
            const uint64_t fs = umax(1, N / (uint64_t)Tf) * 3;
            const uint64_t cf = umin((uint64_t)Tf, run_len(0, fs, N));
            add_run(pl, 0, fs, cf);
            const uint64_t ss = umax(1, N / (uint64_t)Ts);
            add_run(pl, fs / 2, ss, umin((uint64_t)T - cf, run_len(fs / 2, ss, N)));
            break;
        }
        case AQE_M_PARALLEL_POINTER: { // cbd:814-854
            if (N == 0 || T <= 0) break;
            const int64_t Th = P.num_threads;
            if (Th <= 0) { err = "parallel_pointer_sample: num_threads must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t spt = (uint64_t)(T / Th), step = umax(1, N / (uint64_t)T);
            for (int64_t t = 0; t < Th; ++t) {
                const uint64_t start = (N / (uint64_t)Th) * (uint64_t)t;
                add_run(pl, start, step, umin(spt, run_len(start, step, N)));
            }
            break;
        }
        case AQE_M_RANDOM_POINTER: { // cbd:856-882: mt19937(seed), uniform_int_distribution<size_t>, std::set
            if (N == 0 || T <= 0) break;
            const uint64_t want = umin((uint64_t)T, N);
            std::mt19937 rng((unsigned int)P.seed);
            std::uniform_int_distribution<size_t> dist(0, (size_t)N - 1);
            std::vector<bool> seen(N, false);
            pl.idx.reserve(want);
            while (pl.idx.size() < want) {
                const size_t x = dist(rng);
                if (!seen[x]) { seen[x] = true; pl.idx.push_back((int64_t)x); }
            }
            std::sort(pl.idx.begin(), pl.idx.end());  // std::set iterates ascending
            break;
        }
        case AQE_M_MEMORY_STRIDE:                 // cbd:1526-1566
        case AQE_M_RANDOM_START_MEMORY_STRIDE: {  // cbd:1838-1878
            const uint64_t M = cache_rows(N);
            const int64_t Tm = target_of(M, p);
            if (M == 0 || Tm <= 0) break;
            if (P.block_size < 0) { err = "stride_bytes must be >= 0"; return AQE_ERR_INVALID; }
            const uint64_t stride = P.block_size == 0 ? umax(1, M / (uint64_t)Tm) : umax(1, (uint64_t)P.block_size / 32);
            const uint64_t start = method == AQE_M_MEMORY_STRIDE ? 0 : seeded_below(P.seed, (uint32_t)method, 0, stride);
            add_run(pl, start, stride, umin((uint64_t)Tm, run_len(start, stride, M)));
            break;
        }
        case AQE_M_OPT_ADDRESS_ARITHMETIC: { // cbd:1667-1703
            const uint64_t M = cache_rows(N);
            const int64_t Tm = target_of(M, p);
            if (M == 0 || Tm <= 0) break;
            const uint64_t stride = umax(1, M / (uint64_t)Tm);
            add_run(pl, 0, stride, umin((uint64_t)Tm, run_len(0, stride, M)));
            break;
        }
        case AQE_M_BYTE_OFFSET:   // cbd:1461-1481 -> index_based_sample
            if (N == 0 || T <= 0) break;
            [[fallthrough]];
        case AQE_M_INDEX_BASED: { // cbd:444-487: i_k = floor(k * N/T)
            if (N == 0 || p <= 0.0) break;
            if (p >= 100.0) { add_all(pl, N); break; }
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0);
            if (Tu == 0) break;
            aqe_segment s{};
            s.kind = 1; s.scale = (double)N / (double)Tu; s.count = (int64_t)Tu; s.inner_len = 1;
            pl.segs.push_back(s);
            break;
        }
        case AQE_M_OPTIMIZED_CLT: { // cbd:1046-1147 (the CLT check has no effect on the returned rows)
            if (N == 0 || p <= 0.0) break;
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0);
            if (Tu == 0) break;
            const int64_t opt = std::min<int64_t>(P.num_threads, std::max<int64_t>(1, (int64_t)(int)(Tu / 100)));
            if (opt <= 0) { err = "optimized_clt_sample: num_threads must be >= 1"; return AQE_ERR_INVALID; }
            if (N < 5000 || Tu < 200 || opt == 1) {
                const uint64_t step = umax(1, N / Tu);
                add_run(pl, 0, step, umin(Tu, run_len(0, step, N)));
                break;
            }
            const uint64_t spt = Tu / (uint64_t)opt, rpt = N / (uint64_t)opt;
            for (int64_t t = 0; t < opt; ++t) {
                const uint64_t a = (uint64_t)t * rpt, b = t == opt - 1 ? N : (uint64_t)(t + 1) * rpt;
                const uint64_t lt = t == opt - 1 ? Tu - (uint64_t)(opt - 1) * spt : spt;
                if (lt == 0) continue;
                const uint64_t stride = umax(1, (b - a) / lt);
                add_run(pl, a, stride, umin(lt, run_len(a, stride, b)));
            }
            break;
        }
        case AQE_M_BLOCK:  // cbd:1151-1181
        case AQE_M_PAGE: { // cbd:1183-1216: rows per page = page_size / sizeof(Record)
            if (N == 0 || T <= 0) break;
            if (P.block_size < 0) { err = "block_size must be >= 0"; return AQE_ERR_INVALID; }
            uint64_t B = (uint64_t)P.block_size;
            if (method == AQE_M_PAGE) B = umax(1, B / 32);
            if (B == 0) { err = "block_sample: block_size must be >= 1 (the reference divides by zero)"; return AQE_ERR_INVALID; }
            const uint64_t nb = cdiv(N, B);
            const uint64_t k = umax(1, (uint64_t)((double)nb * p / 100.0));
            const uint64_t iv = umax(1, nb / k);
            add_blocks(pl, N, B, 0, cdiv(nb, iv), iv * B, (uint64_t)T);
            break;
        }
        case AQE_M_PARALLEL_BLOCK: { // cbd:1218-1271
            if (N == 0 || T <= 0) break;
            const int64_t Th = P.num_threads;
            if (Th <= 0 || P.block_size <= 0) { err = "parallel_block_sample: num_threads and block_size must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t B = (uint64_t)P.block_size, nb = cdiv(N, B);
            const uint64_t k = umax(1, (uint64_t)((double)nb * p / 100.0));
            const uint64_t bpt = umax(1, k / (uint64_t)Th), iv = umax(1, nb / k), tt = (uint64_t)(T / Th);
            for (int64_t t = 0; t < Th; ++t) {
                const uint64_t sb = (uint64_t)t * bpt, eb = umin(sb + bpt, k);
                add_blocks(pl, N, B, sb, eb, iv * B, tt);
            }
            break;
        }
        case AQE_M_NODE_SKIP: { // cbd:489-532: every skip-th leaf (1-based counter), first rows of each
            if (N == 0 || p <= 0.0) break;
            if (p >= 100.0) { add_all(pl, N); break; }
            if (P.step_size <= 0) { err = "node_skip_sample: skip_factor must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0), skip = (uint64_t)P.step_size, L = leaf_count(N);
            const uint64_t inner = (L - 1) / skip;  // selected leaves that are not the last one: 127 rows each
            const uint64_t a = umin(Tu, inner * 127);
            add_seg(pl, (int64_t)(127 * (skip - 1)), (int64_t)(127 * skip), 127, a);
            if (L % skip == 0) add_seg(pl, (int64_t)(127 * (L - 1)), 0, (int64_t)(N - 127 * (L - 1)), umin(Tu - a, N - 127 * (L - 1)));
            break;
        }
        case AQE_M_DIRECT_ACCESS: { // cbd:584-644
            if (N == 0 || p <= 0.0) break;
            if (p >= 100.0) { add_all(pl, N); break; }
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0), L = leaf_count(N);
            const uint64_t nodes = umax(1, Tu / 10);
            const double node_step = (double)L / (double)nodes;
            pl.idx.reserve(Tu);
            for (uint64_t i = 0; i < nodes && pl.idx.size() < Tu; ++i) {
                const uint64_t leaf = (uint64_t)((double)i * node_step);
                if (leaf >= L) continue;
                const int kc = (int)(leaf + 1 < L ? 127 : N - 127 * (L - 1));
                const int per = std::min(std::max(1, (int)(Tu / nodes)), kc);
                const double rs = (double)kc / per;
                for (int j = 0; j < per && pl.idx.size() < Tu; ++j) {
                    const int r = (int)(j * rs);
                    if (r < kc) pl.idx.push_back((int64_t)(127 * leaf + (uint64_t)r));
                }
            }
            break;
        }
        case AQE_M_BALANCED_TREE: { // cbd:534-582: proportional allocation down the bulk-load shape
            if (N == 0 || p <= 0.0) break;
            if (p >= 100.0) { add_all(pl, N); break; }
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0);
            std::vector<uint64_t> width{leaf_count(N)};  // nodes per level, level 0 = leaves
            while (width.back() > 1) width.push_back(parent_count(width.back()));
            const uint64_t L = width[0];
            // [first leaf, last leaf) spanned by node g of level l
            auto span = [&](int l, uint64_t g, uint64_t& lo, uint64_t& hi) {
                lo = g; hi = g + 1;
                for (int q = l; q > 0; --q) { const bool last = hi == width[q]; lo *= 128; hi = last ? width[q - 1] : hi * 128; }
            };
            auto rows_of = [&](uint64_t lo, uint64_t hi) { return (hi == L ? N : 127 * hi) - 127 * lo; };
            struct Item { int level; uint64_t node, want; };
            std::vector<Item> stack{{(int)width.size() - 1, 0, Tu}};
            pl.idx.reserve(Tu);
            while (!stack.empty() && pl.idx.size() < Tu) {
                const Item it = stack.back(); stack.pop_back();
                if (it.want == 0) continue;
                if (it.level == 0) {
                    const int kc = (int)(it.node + 1 < L ? 127 : N - 127 * (L - 1));
                    const int take = std::min((int)it.want, kc);
                    const double step = (double)kc / take;
                    for (int i = 0; i < take && pl.idx.size() < Tu; ++i) {
                        const int r = (int)(i * step);
                        if (r < kc) pl.idx.push_back((int64_t)(127 * it.node + (uint64_t)r));
                    }
                    continue;
                }
                uint64_t lo, hi; span(it.level, it.node, lo, hi);
                const uint64_t node_rows = rows_of(lo, hi);
                const uint64_t c0 = 128 * it.node, c1 = it.node + 1 == width[it.level] ? width[it.level - 1] : 128 * (it.node + 1);
                for (uint64_t ch = c1; ch-- > c0;) {  // push in reverse so children pop in order
                    uint64_t clo, chi; span(it.level - 1, ch, clo, chi);
                    const uint64_t crow = rows_of(clo, chi);
                    if (crow > 0) stack.push_back({it.level - 1, ch, (it.want * crow) / node_rows});
                }
            }
            break;
        }
        case AQE_M_ADAPTIVE_BLOCK: { // cbd:1273-1329; zone variances come from the device (exact scans)
            if (N == 0 || T <= 0) break;
            if (!data.zone_var) { err = "adaptive_block_sample needs the table on a device"; return AQE_ERR_STATE; }
            const uint64_t mn = (uint64_t)P.block_size, mx = (uint64_t)P.block_size_max, zs = N / 10;
            if (zs == 0) { err = "adaptive_block_sample: needs at least 10 rows (the reference divides 0/0)"; return AQE_ERR_INVALID; }
            double maxvar = data.zone_var[0];
            for (int z = 1; z < 10; ++z) maxvar = std::max(maxvar, data.zone_var[z]);
            uint64_t left = (uint64_t)T;
            for (uint64_t z = 0; z < 10 && left > 0; ++z) {
                const uint64_t s = z * zs, e = umin(s + zs, N);
                const double ratio = data.zone_var[z] / maxvar;
                const uint64_t bs = mn + (uint64_t)((double)(mx - mn) * (1.0 - ratio));
                if (bs == 0) { err = "adaptive_block_sample: block size 0 (the reference would not terminate)"; return AQE_ERR_INVALID; }
                const uint64_t nfull = (e - s) / bs, rem = (e - s) % bs;
                const uint64_t cfull = umin(umax(1, (uint64_t)((double)bs * p / 100.0)), bs);
                const uint64_t a = umin(left, nfull * cfull);
                add_seg(pl, (int64_t)s, (int64_t)bs, (int64_t)cfull, a);
                left -= a;
                if (rem && left) {
                    const uint64_t c = umin(umin(umax(1, (uint64_t)((double)rem * p / 100.0)), rem), left);
                    add_seg(pl, (int64_t)(s + nfull * bs), 0, (int64_t)c, c);
                    left -= c;
                }
            }
            break;
        }
        case AQE_M_STRATIFIED_BLOCK: { // cbd:1331-1379; positions are into the amount-sorted order
            pl.by_amount_order = true;
            if (N == 0 || T <= 0) break;
            const int64_t K = P.block_size_max;
            if (K <= 0 || P.block_size <= 0) { err = "stratified_block_sample: block_size and strata_count must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t B = (uint64_t)P.block_size, ssz = N / (uint64_t)K, sps = (uint64_t)(T / K);
            uint64_t got = 0;
            for (int64_t s = 0; s < K && got < (uint64_t)T; ++s) {
                const uint64_t a = (uint64_t)s * ssz, b = s == K - 1 ? N : a + ssz;
                const uint64_t nb = cdiv(b - a, B);
                const uint64_t k = umax(1, (uint64_t)((double)nb * p / 100.0)), iv = umax(1, nb / k);
                for (uint64_t bi = 0; bi < nb && got < (uint64_t)T; bi += iv) {
                    const uint64_t bs = a + bi * B, be = umin(bs + B, b);
                    const uint64_t take = umin(umin(sps, (uint64_t)T - got), be - bs);
                    add_seg(pl, (int64_t)bs, 0, (int64_t)umax(take, 1), take);
                    got += take;
                }
            }
            break;
        }
        case AQE_M_SAMPLE_RECORDS: { // cbd:345-363: SRSWOR of floor(N p/100) rows = prefix of a seeded permutation of [0,N)
            if (N == 0) break;
            if (p >= 100.0) { add_all(pl, N); break; }
            if (p <= 0.0) break;
            const uint64_t k = umin((uint64_t)((double)N * p / 100.0), N);
            if (k == 0) break;
            aqe_segment s{};
            s.kind = 2; s.base = (int64_t)N; s.outer_step = (int64_t)P.seed; s.inner_len = (int64_t)feistel_half_bits(N); s.count = (int64_t)k;
            pl.segs.push_back(s);
            break;
        }
        case AQE_M_OPTIMIZED_SEQUENTIAL: { // cbd:366-428: 1-based count c is taken when c >= next; next += step
            if (p >= 100.0) { add_all(pl, N); break; }
            if (p <= 0.0 || N == 0) break;
            const uint64_t Tu = (uint64_t)((double)N * p / 100.0);
            if (Tu == 0) break;
            const double step = 100.0 / p;
            volatile double next = step * ((double)(seeded_u64(P.seed, (uint32_t)method, 0) >> 11) * (1.0 / 9007199254740992.0));
            pl.idx.reserve(Tu);
            uint64_t c = 0;  // last count taken
            while (pl.idx.size() < Tu) {
                const double nx = next;
                uint64_t want = nx <= 1.0 ? 1 : (uint64_t)std::ceil(nx);
                if (want <= c) want = c + 1;
                if (want > N) break;
                pl.idx.push_back((int64_t)(want - 1));
                c = want;
                next = nx + step;
            }
            break;
        }
        case AQE_M_RANDOM_START_NTH: { // cbd:1483-1524
            if (N == 0 || T <= 0) break;
            if (P.step_size <= 0) { err = "random_start_nth_sample: nth must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t nth = (uint64_t)P.step_size, start = seeded_below(P.seed, (uint32_t)method, 0, N);
            const uint64_t a = umin((uint64_t)T, run_len(start, nth, N));
            add_run(pl, start, nth, a);
            add_run(pl, 0, nth, umin((uint64_t)T - a, run_len(0, nth, start)));
            break;
        }
        case AQE_M_ADDRESS_ARITHMETIC: { // cbd:1605-1665: i*stride + U{0..stride/2}, wrapped
            if (N == 0 || T <= 0) break;
            const uint64_t M = cache_rows(N), stride = umax(1, M / (uint64_t)T);
            aqe_segment s{};
            s.kind = 3; s.base = (int64_t)M; s.outer_step = (int64_t)stride; s.inner_len = (int64_t)P.seed; s.count = T;
            pl.segs.push_back(s);
            break;
        }
        case AQE_M_MULTITHREADED_MEMORY_STRIDE: { // cbd:1880-1960 (and the index sets of cbd:1962-2048)
            const uint64_t M = cache_rows(N);
            if (M == 0) break;
            const int64_t Th = P.num_threads;
            if (Th <= 0) { err = "multithreaded_memory_stride_sample: num_threads must be >= 1"; return AQE_ERR_INVALID; }
            const double pp = p / (double)Th;
            const uint64_t rs = M / (uint64_t)Th, rem = M % (uint64_t)Th;
            for (int64_t t = 0; t < Th; ++t) {
                const uint64_t a = (uint64_t)t * rs;
                if (a >= M) continue;
                const uint64_t b = umin(a + rs + ((uint64_t)t < rem ? 1 : 0), M), rt = b - a;
                const uint64_t tt = (uint64_t)((double)rt * pp / 100.0);
                if (tt == 0) continue;
                const uint64_t start = a + seeded_below(P.seed, (uint32_t)method, (uint64_t)t, umin(rt / 10, 100) + 1);
                const uint64_t stride = umax(1, rt / tt);
                add_run(pl, start, stride, umin(tt, run_len(start, stride, b)));
            }
            break;
        }
        case AQE_M_SIGNAL_BASED_CLT: { // cbd:1705-1818 under the lock-step schedule
            const uint64_t M = cache_rows(N);
            const int64_t Tm = target_of(M, p);
            if (M == 0 || Tm <= 0) break;
            if (P.check_interval <= 0) { err = "signal_based_clt_sample: check_interval must be >= 1"; return AQE_ERR_INVALID; }
            const uint64_t ci = (uint64_t)P.check_interval, Tu = (uint64_t)Tm;
            const uint64_t fs = umax(2, M / (Tu * 2));
            const uint64_t first_hit = umax(ci, ci * cdiv(Tu / 2, ci));  // first multiple of ci that is >= T/2
            const uint64_t nf = umin(umin(run_len(0, fs, M), Tu), first_hit);
            add_run(pl, 0, fs, nf);
            add_run(pl, 0, 1, umin(umin(nf, Tu / 4), umin(M, Tu - nf)));
            break;
        }
        case AQE_M_CLT_VALIDATED_DUAL_POINTER: { // cbd:885-1043 under the lock-step schedule
            std::vector<CltThread> th;
            int64_t Tc = 0;
            const int rc = clt_threads(N, P, th, Tc, err);
            if (rc != AQE_OK) return rc;
            if (th.empty()) break;
            if (data.clt_kstop < 0) { err = "clt_validated_dual_pointer_sample needs the table on a device"; return AQE_ERR_STATE; }
            uint64_t got = 0;
            for (size_t q = 0; q < th.size(); ++q) {
                uint64_t take = th[q].len;
                if (data.clt_kstop > 0) take = umin(take, (int64_t)q <= data.clt_stopper ? (uint64_t)data.clt_kstop : (uint64_t)data.clt_kstop - 1);
                add_run(pl, th[q].first, th[q].step, take);
                got += take;
            }
            if ((int64_t)got < Tc / 4) {  // cbd:1032-1040 top-up
                const uint64_t step = (uint64_t)std::max<int64_t>(1, (int64_t)(int)(N / (uint64_t)(Tc / 4)));
                add_run(pl, 0, step, umin((uint64_t)Tc - got, run_len(0, step, N)));
            }
            break;
        }
        default:
            err = "unknown sampler id";
            return AQE_ERR_UNSUPPORTED;
    }
    plan_finalize(pl);
    return AQE_OK;
}

}  // namespace aqe
