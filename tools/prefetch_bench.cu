// Does prefetch.global.L2 lift the random-RMW ceiling?  Each thread knows its address D
// iterations ahead and prefetches it into L2 before the read-modify-write reaches it.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
template <int D>
__global__ void k_rmw(double2 *base, uint64_t region_elems, int iters, uint64_t seed) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    double2 *reg = base + warp * region_elems;
    const uint64_t h0 = mix(seed + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    for (int i = 0; i < iters; ++i) {
        if (D > 0) {
            const uint64_t ip = mix(h0 + i + D) % region_elems;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(reg + ip));
        }
        const uint64_t idx = mix(h0 + i) % region_elems;
        double2 v = __ldcg(reg + idx);
        v.x += 1.0; v.y += 0.5;
        __stcg(reg + idx, v);
        if (D == 0 && (i & 1023) == 1023) __syncwarp();
    }
}
template <int D> void run(double2 *buf, uint64_t elems, int warps, int iters) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int threads = 256, blocks = warps * 32 / threads;
    k_rmw<D><<<blocks, threads>>>(buf, elems, 8, 1); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); k_rmw<D><<<blocks, threads>>>(buf, elems, iters, 2); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("prefetch distance %2d  region %.2f MB x %d warps: %.2f ms  %.2f Gtouch/s\n", D, elems * 16.0 / 1e6, warps, ms,
           (double)warps * 32 * iters / ms / 1e6);
}
int main() {
    const size_t bytes = size_t(90) << 30;
    double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    for (uint64_t elems : {80513ull, 1138499ull}) {
        const int warps = 4736;
        run<0>(buf, elems, warps, 1000); run<1>(buf, elems, warps, 1000); run<2>(buf, elems, warps, 1000);
        run<4>(buf, elems, warps, 1000); run<8>(buf, elems, warps, 1000); run<16>(buf, elems, warps, 1000);
    }
    return 0;
}
