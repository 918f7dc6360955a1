// Dependent-load latency of one thread per CTA under the block kernel's conditions.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

// mode 0: lone thread chases; others exit.  mode 1: others wait at a barrier each step.
// mode 2: others do one random RMW gather per step (like a push batch), then barrier.
__global__ void k_chase(double2 *base, uint64_t region_elems, int steps, int mode, unsigned long long *out_ns) {
    double2 *reg = base + (uint64_t)blockIdx.x * region_elems;
    uint64_t h = mix(blockIdx.x * 7919ull + threadIdx.x);
    unsigned long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    double acc = 0;
    uint64_t idx = h % region_elems;
    for (int s = 0; s < steps; ++s) {
        if (threadIdx.x == 0) {
            double2 v = __ldcg(reg + idx);          // dependent chain: next index depends on the value
            acc += v.x;
            idx = mix(idx + (uint64_t)(v.y) + s) % region_elems;
        } else if (mode == 2) {
            h = mix(h + s);
            double2 *p = reg + (h % region_elems);
            double2 v = __ldcg(p); v.x += 1.0; __stcg(p, v);
        }
        if (mode >= 1) __syncthreads();
        else if (threadIdx.x != 0) return;
    }
    if (threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out_ns[blockIdx.x] = t1 - t0;
        if (acc == 12345.678) printf("x");
    }
}

int main() {
    const size_t bytes = size_t(24) << 30;
    double2 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    unsigned long long *d_ns; CK(cudaMalloc(&d_ns, 8 * 4096));
    unsigned long long h_ns[4096];
    struct { uint64_t elems; int ctas; int threads; int mode; const char *name; } cfgs[] = {
        {80513, 296, 512, 0, "lone thread, 1.29MB regions x296 (382MB)"},
        {80513, 296, 512, 1, "lone thread + 511 at barrier"},
        {80513, 296, 512, 2, "lone thread + 511 random RMW + barrier"},
        {80513, 148, 512, 2, "same, 148 CTAs"},
        {4096, 296, 512, 2, "same, 64KB regions (L2 resident)"},
        {4096, 296, 512, 0, "lone thread, 64KB regions (L2 resident)"},
        {1138499, 296, 512, 0, "lone thread, 18MB regions x296 (5.4GB)"},
        {4000000, 296, 512, 0, "lone thread, 64MB regions x296 (19GB)"},
    };
    const int steps = 2000;
    for (auto c : cfgs) {
        k_chase<<<c.ctas, c.threads>>>(buf, c.elems, 50, c.mode, d_ns);
        CK(cudaDeviceSynchronize());
        k_chase<<<c.ctas, c.threads>>>(buf, c.elems, steps, c.mode, d_ns);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h_ns, d_ns, 8 * c.ctas, cudaMemcpyDeviceToHost));
        double sum = 0; for (int i = 0; i < c.ctas; ++i) sum += h_ns[i];
        printf("%-48s  %.3f us per step\n", c.name, sum / c.ctas / steps / 1e3);
    }
    return 0;
}
