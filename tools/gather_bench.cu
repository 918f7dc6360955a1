// Micro-benchmark: random 16-byte read-modify-write over a large array, the access pattern
// of the push kernel's {s, r} state.  Measures touches/s for several load/store flavours
// and working-set sizes.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

template <int MODE>
__device__ __forceinline__ double2 ld16(const double2 *p) {
    double2 v;
    if (MODE == 0) v = *p;
    else if (MODE == 1) asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.cv.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else asm volatile("ld.global.L1::no_allocate.L2::64B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <int MODE>
__device__ __forceinline__ void st16(double2 *p, double2 v) {
    if (MODE == 0) *p = v;
    else if (MODE == 1 || MODE == 3) asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y));
    else asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y));
}

// each warp owns a private region (like a slot); lanes touch random elements of it
template <int MODE, int STRIDE16>  // STRIDE16: element stride in units of 16 bytes (1 = packed, 2 = 32-byte entries)
__global__ void k_rmw(double2 *base, uint64_t region_elems, int iters, uint64_t seed) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    double2 *reg = base + warp * region_elems * STRIDE16;
    uint64_t h = mix(seed + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
        h = mix(h + i);
        const uint64_t idx = (h % region_elems) * STRIDE16;
        double2 v = ld16<MODE>(reg + idx);
        v.x += 1.0; v.y += 0.5;
        st16<MODE>(reg + idx, v);
    }
}

template <int MODE, int STRIDE16>
static void run(const char *name, double2 *buf, uint64_t region_elems, int warps, int iters) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int threads = 256, blocks = warps * 32 / threads;
    k_rmw<MODE, STRIDE16><<<blocks, threads>>>(buf, region_elems, 8, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_rmw<MODE, STRIDE16><<<blocks, threads>>>(buf, region_elems, iters, 2);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double touches = (double)warps * 32 * iters;
    printf("%-28s region=%8.2f MB x %5d warps (%7.1f MB)  %.2f ms  %.2f Gtouch/s\n", name,
           region_elems * 16.0 * STRIDE16 / 1e6, warps, warps * region_elems * 16.0 * STRIDE16 / 1e6, ms, touches / ms / 1e6);
}

int main(int argc, char **argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)); }
    size_t g; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity));
    printf("MaxL2FetchGranularity=%zu\n", g);
    const size_t bytes = size_t(16) << 30;
    double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    const int iters = 2000;
    // (region elements, warps): flickr-like 80K-node slots; youtube-like 1.1M-node slots; L2-resident
    struct { uint64_t elems; int warps; } cfgs[] = {{1024, 4736}, {4096, 1184}, {8192, 592}, {16384, 296}, {32768, 148}, {32768, 296}, {32768, 592}, {80513, 148}, {80513, 296}};
    for (auto c : cfgs) {
        if ((double)c.elems * 32 * c.warps > (double)bytes) continue;
        run<0, 1>("default ld/st 16B", buf, c.elems, c.warps, iters);
        run<1, 1>("ld.cg/st.cg 16B", buf, c.elems, c.warps, iters);
        run<2, 1>("L1::no_allocate 16B", buf, c.elems, c.warps, iters);
        run<3, 1>("ld.cv/st.cg 16B", buf, c.elems, c.warps, iters);
        run<4, 1>("no_alloc L2::64B 16B", buf, c.elems, c.warps, iters);
        run<0, 2>("default, 32B entries", buf, c.elems, c.warps, iters);
        run<2, 2>("no_allocate, 32B entries", buf, c.elems, c.warps, iters);
    }
    return 0;
}
