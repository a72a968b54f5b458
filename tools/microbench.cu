// microbench.cu -- measured ceilings that the sampled-path rooflines are quoted against (tools/, not product code):
//
//   gather   random / strided 8-byte reads out of a large column: samples/s and (under ncu) DRAM bytes per sample, for the load
//            forms a gather kernel can use (ld.global.nc, ld.global.ca, ld.global.cg, L1::no_allocate) x loads in flight per
//            thread x the device's L2 fetch-granularity limit (32 / 64 / 128 bytes).
//   h2d      pinned host -> device cudaMemcpyAsync bandwidth, 1 GiB, per GPU and all GPUs at once (one thread per GPU).
//   atoms    issue rate of 32-bit shared-memory atomics per SM (clocks per warp instruction) by address pattern, use of the return
//            value, dependence between atomics and active lanes: the ceiling of the shared-bin GROUP BY kernels (k_sql_ring, G > 16).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/microbench.cu -o tools/_bin/microbench
//   tools/_bin/microbench gather [elems] | h2d [n_gpus] | atoms [groups]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <int MODE> __device__ __forceinline__ double load8(const double* p) {
    double v;
    if constexpr (MODE == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if constexpr (MODE == 1) asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if constexpr (MODE == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// pattern 0: uniformly random element; 1: element k * stride (memory_stride_sample's progression)
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ col, uint64_t n, uint64_t count, int pattern, uint64_t stride, double* out) {
    const uint64_t G = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; k + (uint64_t)(U - 1) * G < count; k += (uint64_t)U * G) {
        double x[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t kk = k + (uint64_t)j * G;
            const uint64_t pos = pattern == 0 ? __umul64hi(mix(kk), n) : (kk * stride) % n;
            x[j] = load8<MODE>(col + pos);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) acc += x[j];
    }
    if (acc == 123.456) *out = acc;
}

template <int MODE, int U> static float run_gather(const double* col, uint64_t n, uint64_t count, int pattern, uint64_t stride, double* out, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 8;
    for (int i = 0; i < 2; ++i) k_gather<MODE, U><<<grid, 256>>>(col, n, count, pattern, stride, out);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_gather<MODE, U><<<grid, 256>>>(col, n, count, pattern, stride, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

static int gather_main(uint64_t n) {
    double *col = nullptr, *out = nullptr;
    CK(cudaMalloc(&col, n * 8));
    CK(cudaMalloc(&out, 8));
    CK(cudaMemset(col, 0, n * 8));
    const uint64_t count = 16u << 20;
    const bool quick = std::getenv("MB_QUICK") != nullptr;   // the short list (for a run under ncu)
    for (size_t gran : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
        if (gran) {
            if (quick && gran == 64) continue;
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            if (e != cudaSuccess) { std::printf("{\"l2_fetch_granularity\": %zu, \"error\": \"%s\"}\n", gran, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        }
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        for (int pattern = 0; pattern < 2; ++pattern) {
            auto line = [&](const char* mode, int u, float ms) {
                std::printf("{\"bench\": \"gather\", \"l2_fetch_granularity\": %zu, \"pattern\": \"%s\", \"load\": \"%s\", \"in_flight\": %d, \"ms\": %.4f, "
                            "\"Gsamples_per_s\": %.2f, \"sector_GBps\": %.1f}\n", got, pattern ? "stride100" : "random", mode, u, ms, count / ms / 1e6, count * 32.0 / ms / 1e6);
                std::fflush(stdout);
            };
            const int reps = 5;
            line("nc", 4, run_gather<0, 4>(col, n, count, pattern, 100, out, reps));
            line("cg", 4, run_gather<2, 4>(col, n, count, pattern, 100, out, reps));
            if (quick) continue;
            line("nc", 1, run_gather<0, 1>(col, n, count, pattern, 100, out, reps));
            line("nc", 2, run_gather<0, 2>(col, n, count, pattern, 100, out, reps));
            line("nc", 8, run_gather<0, 8>(col, n, count, pattern, 100, out, reps));
            line("nc", 16, run_gather<0, 16>(col, n, count, pattern, 100, out, reps));
            line("ca", 4, run_gather<1, 4>(col, n, count, pattern, 100, out, reps));
            line("cg", 8, run_gather<2, 8>(col, n, count, pattern, 100, out, reps));
            line("cg", 16, run_gather<2, 16>(col, n, count, pattern, 100, out, reps));
            line("nc.no_allocate", 4, run_gather<3, 4>(col, n, count, pattern, 100, out, reps));
            line("nc.no_allocate", 8, run_gather<3, 8>(col, n, count, pattern, 100, out, reps));
        }
    }
    return 0;
}

static int h2d_main(int ngpu) {
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (ngpu <= 0 || ngpu > have) ngpu = have;
    const size_t bytes = 1ull << 30;
    std::vector<void*> host(ngpu), dev(ngpu);
    std::vector<cudaStream_t> st(ngpu);
    for (int g = 0; g < ngpu; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaHostAlloc(&host[g], bytes, cudaHostAllocPortable));
        std::memset(host[g], 1, bytes);
        CK(cudaMalloc(&dev[g], bytes));
        CK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
    }
    auto one = [&](int g, int reps) -> double {   // GB/s of `reps` back-to-back 1 GiB copies on GPU g
        CK(cudaSetDevice(g));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemcpyAsync(dev[g], host[g], bytes, cudaMemcpyHostToDevice, st[g]));
        CK(cudaEventRecord(e0, st[g]));
        for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(dev[g], host[g], bytes, cudaMemcpyHostToDevice, st[g]));
        CK(cudaEventRecord(e1, st[g]));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        return reps * (double)bytes / ms / 1e6;
    };
    for (int g = 0; g < ngpu; ++g) std::printf("{\"bench\": \"h2d\", \"gpu\": %d, \"alone_GBps\": %.2f}\n", g, one(g, 4));
    std::vector<double> r(ngpu);
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; ++g) th.emplace_back([&, g] { r[g] = one(g, 6); });
    for (auto& t : th) t.join();
    double sum = 0;
    for (int g = 0; g < ngpu; ++g) sum += r[g];
    std::printf("{\"bench\": \"h2d\", \"gpus\": %d, \"concurrent_total_GBps\": %.2f, \"per_gpu_min_GBps\": %.2f}\n", ngpu, sum, *std::min_element(r.begin(), r.end()));
    return 0;
}

// h2dnuma: does the placement of the pinned host buffers limit the all-GPUs-at-once rate?  Buffers are mmap'ed, given a NUMA
// policy with mbind(2) (default / interleaved over all nodes / bound to node k), touched, registered with CUDA and copied from on
// all GPUs at once.  Prints what the kernel allows (nodes online, this process' allowed memory nodes).
static int h2dnuma_main(int ngpu) {
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (ngpu <= 0 || ngpu > have) ngpu = have;
    int nodes = 0;
    for (int k = 0; k < 64; ++k) { char path[64]; std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d", k); if (access(path, F_OK) == 0) nodes = k + 1; }
    if (nodes < 1) nodes = 1;
    {
        FILE* f = std::fopen("/proc/self/status", "r");
        char line[512];
        while (f && std::fgets(line, sizeof(line), f))
            if (!std::strncmp(line, "Mems_allowed_list", 17) || !std::strncmp(line, "Cpus_allowed_list", 17)) { line[std::strcspn(line, "\n")] = 0; std::printf("{\"bench\": \"h2dnuma\", \"status\": \"%s\"}\n", line); }
        if (f) std::fclose(f);
    }
    const size_t bytes = 1ull << 30;
    std::vector<void*> dev(ngpu);
    std::vector<cudaStream_t> st(ngpu);
    for (int g = 0; g < ngpu; ++g) { CK(cudaSetDevice(g)); CK(cudaMalloc(&dev[g], bytes)); CK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking)); }
    for (int policy = -2; policy < nodes; ++policy) {   // -2 default, -1 interleave over all nodes, k >= 0 bind to node k
        std::vector<void*> host(ngpu, nullptr);
        bool ok = true;
        for (int g = 0; g < ngpu && ok; ++g) {
            void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) { ok = false; break; }
            host[g] = p;
            if (policy != -2) {
                unsigned long mask = policy == -1 ? ((nodes >= 64 ? ~0ul : ((1ul << nodes) - 1))) : (1ul << policy);
                const long rc = syscall(SYS_mbind, p, bytes, policy == -1 ? 3 /* MPOL_INTERLEAVE */ : 2 /* MPOL_BIND */, &mask, (unsigned long)(nodes + 1), 0u);
                if (rc != 0) { std::printf("{\"bench\": \"h2dnuma\", \"policy\": %d, \"error\": \"mbind failed (errno %d)\"}\n", policy, errno); ok = false; }
            }
            if (ok) { std::memset(p, 1, bytes); CK(cudaSetDevice(g)); if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); ok = false; } }
        }
        if (ok) {
            std::vector<double> r(ngpu);
            std::vector<std::thread> th;
            for (int g = 0; g < ngpu; ++g)
                th.emplace_back([&, g] {
                    CK(cudaSetDevice(g));
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    CK(cudaMemcpyAsync(dev[g], host[g], bytes, cudaMemcpyHostToDevice, st[g]));
                    CK(cudaEventRecord(e0, st[g]));
                    for (int i = 0; i < 6; ++i) CK(cudaMemcpyAsync(dev[g], host[g], bytes, cudaMemcpyHostToDevice, st[g]));
                    CK(cudaEventRecord(e1, st[g]));
                    CK(cudaEventSynchronize(e1));
                    float ms = 0;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    r[g] = 6.0 * (double)bytes / ms / 1e6;
                });
            for (auto& t : th) t.join();
            double sum = 0;
            for (double v : r) sum += v;
            std::printf("{\"bench\": \"h2dnuma\", \"gpus\": %d, \"numa_nodes\": %d, \"policy\": \"%s%d\", \"concurrent_total_GBps\": %.2f, \"per_gpu_min_GBps\": %.2f, \"per_gpu_max_GBps\": %.2f}\n",
                        ngpu, nodes, policy == -2 ? "default" : (policy == -1 ? "interleave" : "bind node "), policy < 0 ? 0 : policy, sum,
                        *std::min_element(r.begin(), r.end()), *std::max_element(r.begin(), r.end()));
            std::fflush(stdout);
        }
        for (int g = 0; g < ngpu; ++g) if (host[g]) { cudaHostUnregister(host[g]); cudaGetLastError(); munmap(host[g], bytes); }
    }
    return 0;
}


// ---- atoms: shared-memory atomic issue rate ------------------------------------------------------------------------------
// Every thread makes `iters` x 8 "rows"; a row is one of the shapes below on pseudo-random bins of a G-word table (limb-major
// words w0 = s[g], w1 = s[G + g], ...).  Reported: SM clocks per warp-level ATOMS instruction and per row.
//  0 one atomic, result unused            1 one atomic, result summed into a register     2 one atomic whose addend depends on the previous result
//  3 like 0, bank = lane (no conflicts)   4 like 1, bank = lane                            5 atomicAdd(p, 1) (ATOMS.POPC.INC)
//  6 like 0, 4 of 32 lanes active         7 like 0, 8 replicas of the table, replica = lane % 8
//  8 two atomics, the second takes the first's carry (k_sql_ring packed bins, w0 -> w1), rows independent, no branch
//  9 the packed row: w2 (result tested against a limit), w0 -> w1, branch on w1's wrap    10 three independent atomics, results unused
// 11 like 9 with a fourth independent atomic (general form: count word + limbs)          12 like 10 plus a count word (four, results unused)
template <int V> __global__ void __launch_bounds__(256) k_atoms(unsigned int G, int iters, int pattern, unsigned int* out, long long* clk) {
    extern __shared__ unsigned int s[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (unsigned int i = tid; i < 8 * G; i += blockDim.x) s[i] = 0;
    __syncthreads();
    unsigned int x = (blockIdx.x * 256u + tid) * 2654435761u + 12345u, acc = 0, carry = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x = x * 1664525u + 1013904223u;
            unsigned int h = x ^ (x >> 16);            // the LCG alone keeps the lanes of a warp in an arithmetic progression (a low-discrepancy,
            h *= 0x7feb352du; h ^= h >> 15;            // nearly conflict-free set of bins): hash it so that the 32 bins of a warp row are independent
            h *= 0x846ca68bu; h ^= h >> 16;
            unsigned int g = pattern == 0 ? __umulhi(h, G) : __umulhi(x, G);
            const unsigned int v = h | 0x80000000u;
            if (V == 0) atomicAdd(s + g, v);
            if (V == 1) acc += atomicAdd(s + g, v);
            if (V == 2) { const unsigned int a = v + carry; const unsigned int o = atomicAdd(s + g, a); carry = (o + a < o) ? 1u : 0u; }
            if (V == 3 || V == 4) { g = (g & ~31u) | lane; if (g >= G) g -= 32; if (V == 3) atomicAdd(s + g, v); else acc += atomicAdd(s + g, v); }
            if (V == 5) atomicAdd(s + g, 1u);
            if (V == 6) { if ((lane & 7) == 0) atomicAdd(s + g, v); }
            if (V == 7) atomicAdd(s + g * 8 + (lane & 7), v);
            if (V == 8) { const unsigned int o = atomicAdd(s + g, v); atomicAdd(s + G + g, (v >> 10) + ((o + v < o) ? 1u : 0u)); }
            if (V == 9 || V == 11) {
                const unsigned int o2 = atomicAdd(s + 2 * G + g, (v >> 24) + (1u << 20));
                if (V == 11) atomicAdd(s + 4 * G + g, 1u);
                const unsigned int o = atomicAdd(s + g, v);
                const unsigned int a1 = (v >> 10) + ((o + v < o) ? 1u : 0u);
                const unsigned int o1 = atomicAdd(s + G + g, a1);
                if (o1 + a1 < o1) atomicAdd(s + 3 * G + g, 1u);
                if ((o2 >> 20) >= 4000u) { atomicExch(s + 2 * G + g, 0u); acc += 1; }
            }
            if (V == 10 || V == 12) { atomicAdd(s + g, v >> 11); atomicAdd(s + G + g, (v >> 5) & 0x1fffffu); atomicAdd(s + 2 * G + g, v & 0xfffffu); if (V == 12) atomicAdd(s + 3 * G + g, 1u); }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u || carry == 77u) out[0] = s[tid % G];   // keep the results alive
}

template <int V> static void run_atoms(unsigned int G, int ctas_per_sm, int pattern, int atoms_per_row, const char* what, unsigned int* out, long long* clk, int sms) {
    const int iters = 2000;
    const size_t smem = (size_t)8 * G * 4;
    cudaFuncSetAttribute(k_atoms<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sms * ctas_per_sm;
    k_atoms<V><<<grid, 256, smem>>>(G, 10, pattern, out, clk);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_atoms<V><<<grid, 256, smem>>>(G, iters, pattern, out, clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (long long c : h) mean += (double)c;
    mean /= grid;
    const double rows_per_sm_warp = (double)iters * 8 * ctas_per_sm * 8;   // warp-rows per SM
    printf("{\"variant\": %d, \"what\": \"%s\", \"bins\": \"%s\", \"groups\": %u, \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"clk_per_warp_row_per_sm\": %.2f, "
           "\"clk_per_warp_atomic_per_sm\": %.2f, \"ms_per_1e9_rows_148_sms\": %.3f, \"err\": \"%s\"}\n",
           V, what, pattern == 0 ? "independent per lane" : "arithmetic progression over the lanes (few bank conflicts)", G, ctas_per_sm, ctas_per_sm * 8, ms, mean / rows_per_sm_warp, mean / rows_per_sm_warp / atoms_per_row,
           ms * 1e9 / ((double)iters * 8 * 256 * grid), cudaGetErrorString(cudaGetLastError()));
}

static int atoms_main(unsigned int G) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    unsigned int* out; long long* clk;
    cudaMalloc(&out, 4096); cudaMalloc(&clk, sizeof(long long) * 4096);
    for (int pattern : {0, 1})
        for (int c : {3, 4}) {
            const int n = prop.multiProcessorCount;
            run_atoms<0>(G, c, pattern, 1, "one atomic, result unused", out, clk, n);
            run_atoms<1>(G, c, pattern, 1, "one atomic, result summed", out, clk, n);
            run_atoms<2>(G, c, pattern, 1, "one atomic, addend depends on the previous result", out, clk, n);
            run_atoms<3>(G, c, pattern, 1, "result unused, bank = lane", out, clk, n);
            run_atoms<5>(G, c, pattern, 1, "atomicAdd(p, 1): POPC.INC", out, clk, n);
            run_atoms<6>(G, c, pattern, 1, "result unused, 4 of 32 lanes", out, clk, n);
            run_atoms<7>(G, c, pattern, 1, "result unused, 8 replicas by lane", out, clk, n);
            run_atoms<8>(G, c, pattern, 2, "two atomics, carry from the first into the second", out, clk, n);
            run_atoms<9>(G, c, pattern, 3, "packed row: w2 / w0 -> w1, branch on the wrap", out, clk, n);
            run_atoms<10>(G, c, pattern, 3, "three independent atomics, results unused", out, clk, n);
            run_atoms<11>(G, c, pattern, 4, "packed row + a count word", out, clk, n);
            run_atoms<12>(G, c, pattern, 4, "four independent atomics, results unused", out, clk, n);
        }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "atoms")) return atoms_main(argc >= 3 ? (unsigned int)std::atoi(argv[2]) : 1000u);
    if (argc >= 2 && !std::strcmp(argv[1], "h2dnuma")) return h2dnuma_main(argc >= 3 ? std::atoi(argv[2]) : 0);
    if (argc >= 2 && !std::strcmp(argv[1], "gather")) return gather_main(argc >= 3 ? std::strtoull(argv[2], nullptr, 10) : (1ull << 30));
    if (argc >= 2 && !std::strcmp(argv[1], "h2d")) return h2d_main(argc >= 3 ? std::atoi(argv[2]) : 0);
    std::fprintf(stderr, "usage: microbench gather [elems] | h2d [n_gpus] | h2dnuma [n_gpus]\n");
    return 2;
}
