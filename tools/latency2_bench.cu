// latency2_bench.cu -- dependent-chain latency of random 16-byte accesses (plain load, returning
// atomic pair, load-then-atomics) at a given number of resident warps: what one step of a walk
// waits for when few requests are in flight.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned long long *buf, uint64_t entries, int iters, unsigned long long *sink, long long *cycles)
{
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 777;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long *p = buf + 2 * (x % entries);
        unsigned long long got;
        if (MODE == 0) got = __ldcg(p);
        else if (MODE == 1) { got = atomicAdd(p, 0ull); got += atomicAdd(p + 1, 0ull); }
        else { got = __ldcg(p); got = atomicAdd(p, got >> 63) + atomicAdd(p + 1, got >> 63); }
        x += got;   // next address depends on the result (buffer is all zero, so the sequence is unchanged)
    }
    if (threadIdx.x == 0) atomicAdd((unsigned long long *)cycles, (unsigned long long)(clock64() - t0));
    if (x == 42) *sink = x;
}
int main()
{
    unsigned long long *buf, *sink; long long *cyc;
    const uint64_t big = (uint64_t)16 << 30;
    cudaMalloc(&buf, big); cudaMemset(buf, 0, big); cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    for (uint64_t mb : {32ull, 16384ull})
        for (int blocks : {148, 148 * 4, 148 * 16})
            for (int threads : {32, 256}) {
                const int iters = 2000;
                long long h[3];
                for (int mode = 0; mode < 3; ++mode) {
                    cudaMemset(cyc, 0, 8);
                    if (mode == 0) k<0><<<blocks, threads>>>(buf, (mb << 20) / 16, iters, sink, cyc);
                    if (mode == 1) k<1><<<blocks, threads>>>(buf, (mb << 20) / 16, iters, sink, cyc);
                    if (mode == 2) k<2><<<blocks, threads>>>(buf, (mb << 20) / 16, iters, sink, cyc);
                    cudaMemcpy(&h[mode], cyc, 8, cudaMemcpyDeviceToHost);
                }
                printf("region %6llu MB  %5d CTAs x %3d thr: cycles per dependent step: load %7.0f | atomic pair %7.0f | load+atomic pair %7.0f\n",
                       (unsigned long long)mb, blocks, threads, (double)h[0] / blocks / iters, (double)h[1] / blocks / iters, (double)h[2] / blocks / iters);
            }
    return 0;
}
