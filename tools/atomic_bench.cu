// atomic_bench.cu -- throughput of returning 64-bit global atomics at random addresses, the
// access pattern of the frontier schedule (csrc/push_frontier.cu): region resident in L2 or in
// DRAM, with or without a plain load of the sector just before the atomic.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_bench atomic_bench.cu && ./atomic_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t rng(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

// mode 0: atomicAdd(s) + atomicAdd(r) with return; 1: __ldcg of the 16-byte pair first, then the atomics;
// 2: plain load + store of the pair (no atomics)
template <int MODE>
__global__ void k(unsigned long long *buf, uint64_t entries, int iters, unsigned long long *sink)
{
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; ++i) {
        const uint64_t e = rng(s) % entries;
        unsigned long long *p = buf + 2 * e;
        if (MODE == 1) { const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p)); acc += v.x; }
        if (MODE == 2) {
            ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
            v.x += 3; v.y += 3;
            __stcg(reinterpret_cast<ulonglong2 *>(p), v);
            acc += v.x;
        } else {
            acc += atomicAdd(p, 3ull);
            acc += atomicAdd(p + 1, 3ull);
        }
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

template <int MODE> float run(unsigned long long *buf, uint64_t entries, int blocks, int threads, int iters, unsigned long long *sink)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(buf, entries, 4, sink);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(buf, entries, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
    unsigned long long *buf, *sink;
    const uint64_t big = (uint64_t)96 << 30;
    cudaMalloc(&buf, big); cudaMemset(buf, 0, big); cudaMalloc(&sink, 8);
    const int iters = 200;
    // region sweep: L2-resident, then DRAM-resident with a growing number of 2 MB pages (TLB reach)
    for (uint64_t mb : {48ull, 256ull, 1024ull, 4096ull, 16384ull, 65536ull, 98304ull}) {
        const uint64_t bytes = mb << 20;
        const uint64_t entries = bytes / 16;
        for (int blocks : {148 * 8}) {
            const int threads = 256;
            const double n = (double)blocks * threads * iters;
            float m0 = run<0>(buf, entries, blocks, threads, iters, sink);
            float m1 = run<1>(buf, entries, blocks, threads, iters, sink);
            float m2 = run<2>(buf, entries, blocks, threads, iters, sink);
            printf("region %6.0f MB  threads %7d : atomics %7.2f G pairs/s | load+atomics %7.2f | plain ld/st %7.2f\n",
                   bytes / 1048576.0, blocks * threads, n / m0 / 1e6, n / m1 / 1e6, n / m2 / 1e6);
        }
    }
    return 0;
}
