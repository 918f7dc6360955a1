// primitives.cu -- exclusive scan + stable LSD radix sort (see primitives.cuh).
#include "primitives.cuh"

namespace arcte {

// ------------------------------------------------------------------------------------
// Exclusive scan: three-phase (tile sums -> recursive scan of tile sums -> tile rescan).
// ------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total)
{
    // warp inclusive scan, then scan of the 8 warp totals
    __shared__ int64_t warp_tot[kScanThreads / 32];
    __shared__ int64_t block_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) {
            int64_t t = warp_tot[w];
            warp_tot[w] = run;
            run += t;
        }
        block_tot = run;
    }
    __syncthreads();
    *total = block_tot;
    const int64_t res = warp_tot[warp] + inc - v;
    __syncthreads();
    return res;
}

template <typename Tin>
__global__ void __launch_bounds__(kScanThreads)
k_scan_tile_sums(const Tin *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) local += (int64_t)in[idx];
    }
    int64_t total;
    block_exclusive_scan(local, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename Tin>
__global__ void __launch_bounds__(kScanThreads)
k_scan_tiles(const Tin *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_prefix,
             int64_t *__restrict__ out)
{
    // each thread owns kScanItems CONSECUTIVE elements so the scan order is the index order
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t v[kScanItems];
    int64_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        v[i] = idx < n ? (int64_t)in[idx] : 0;
        local += v[i];
    }
    int64_t total;
    int64_t run = block_exclusive_scan(local, &total) + tile_prefix[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
    // the very last tile also writes the grand total
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = run;
}

__global__ void k_scan_single_zero(int64_t *out) { out[0] = 0; }

template <typename Tin>
static int scan_impl(const Tin *in, int64_t *out, int64_t n, DevBuf &scratch, size_t scratch_off,
                     cudaStream_t stream, int64_t *launches)
{
    if (n == 0) {
        k_scan_single_zero<<<1, 1, 0, stream>>>(out);
        ++*launches;
        return ARCTE_OK;
    }
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    // scratch layout at this recursion level: tile_sums[tiles], tile_prefix[tiles+1]
    int64_t *tile_sums = scratch.as<int64_t>() + scratch_off;
    int64_t *tile_prefix = tile_sums + tiles;
    if (tiles == 1) {
        k_scan_single_zero<<<1, 1, 0, stream>>>(tile_prefix);
        ++*launches;
    } else {
        k_scan_tile_sums<Tin><<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, n, tile_sums);
        ++*launches;
        ARCTE_TRY(scan_impl<int64_t>(tile_sums, tile_prefix, tiles, scratch,
                                     scratch_off + (size_t)(2 * tiles + 2), stream, launches));
    }
    k_scan_tiles<Tin><<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, n, tile_prefix, out);
    ++*launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

static size_t scan_scratch_elems(int64_t n)
{
    size_t total = 8;
    while (n > 0) {
        const int64_t tiles = (n + kScanTile - 1) / kScanTile;
        total += (size_t)(2 * tiles + 2);
        if (tiles == 1) break;
        n = tiles;
    }
    return total;
}

int exclusive_scan_i32(const int32_t *in, int64_t *out, int64_t n, DevBuf &scratch,
                       cudaStream_t stream, int64_t *launches)
{
    ARCTE_TRY(dev_reserve(scratch, scan_scratch_elems(n) * sizeof(int64_t)));
    return scan_impl<int32_t>(in, out, n, scratch, 0, stream, launches);
}

int exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, DevBuf &scratch,
                       cudaStream_t stream, int64_t *launches)
{
    ARCTE_TRY(dev_reserve(scratch, scan_scratch_elems(n) * sizeof(int64_t)));
    return scan_impl<int64_t>(in, out, n, scratch, 0, stream, launches);
}

// ------------------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits.
//   pass = histogram kernel -> exclusive scan of hist[digit][tile] -> scatter kernel.
// Stability inside a tile: each warp owns a contiguous sub-tile and walks it in index
// order 32 elements at a time; __match_any_sync ranks equal digits by lane.
// The scatter kernel first REORDERS the tile in shared memory (digit-major, tile order kept
// inside a digit) and then writes it out front to back: consecutive threads write consecutive
// addresses of one digit's global run, whole sectors at a time.  (Round 1 let every lane store
// its pair straight to its own global position: 510 GB/s on 517 M pairs)
// ------------------------------------------------------------------------------------
constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 16;                          // 32-element rounds per warp
constexpr int kSortTile = kSortWarps * kSortRounds * 32; // 8192
constexpr int kDigits = 256;

__global__ void __launch_bounds__(kSortThreads)
k_radix_hist(const uint32_t *__restrict__ keys, int64_t n, int shift, int64_t n_tiles,
             int32_t *__restrict__ hist)
{
    __shared__ int32_t sh[kDigits];
    if (threadIdx.x < kDigits) sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int i = 0; i < kSortTile / kSortThreads; ++i) {
        const int64_t idx = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(keys[idx] >> shift) & 0xff], 1);
    }
    __syncthreads();
    if (threadIdx.x < kDigits) hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

template <typename V> struct SortShared {
    uint32_t key[kSortTile];
    V val[kSortTile];
    int32_t woff[kSortWarps][kDigits + 1];   // per warp and digit: count, then local start
    int32_t dstart[kDigits + 1];             // local start of every digit in the reordered tile
    int64_t gbase[kDigits];                  // global start of (digit, tile) minus dstart[digit]
};

template <typename V>
__global__ void __launch_bounds__(kSortThreads)
k_radix_scatter(const uint32_t *__restrict__ keys, const V *__restrict__ vals, int64_t n, int shift,
                int64_t n_tiles, const int64_t *__restrict__ hist_prefix,
                uint32_t *__restrict__ keys_out, V *__restrict__ vals_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortShared<V> &S = *reinterpret_cast<SortShared<V> *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    for (int i = threadIdx.x; i < kSortWarps * (kDigits + 1); i += kSortThreads) (&S.woff[0][0])[i] = 0;
    __syncthreads();

    const int64_t tbase = (int64_t)blockIdx.x * kSortTile;
    const int64_t wbase = tbase + (int64_t)warp * (kSortRounds * 32);
    const int tile_n = (int)((n - tbase) < (int64_t)kSortTile ? (n - tbase) : (int64_t)kSortTile);
    uint32_t key[kSortRounds];
    // sweep 1: per-warp digit counts
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys[idx] : 0u;
        const int d = valid ? (int)((key[r] >> shift) & 0xff) : kDigits;
        const unsigned peers = __match_any_sync(kFull, d);
        if ((peers & lt) == 0) S.woff[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: totals of the tile -> exclusive scan over the digits -> per-warp local starts
    int total = 0;
    if (threadIdx.x < kDigits) {
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) total += S.woff[w][threadIdx.x];
    }
    {   // exclusive scan of `total` over threads 0..255 (8 warps)
        int incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (threadIdx.x < kDigits && lane == 31) S.dstart[warp] = incl;   // warp sums, parked in dstart[0..7]
        __syncthreads();
        int before = 0;
        if (threadIdx.x < kDigits)
            for (int w = 0; w < warp; ++w) before += S.dstart[w];
        __syncthreads();
        if (threadIdx.x < kDigits) {
            const int start = before + incl - total;
            S.dstart[threadIdx.x] = start;
            S.gbase[threadIdx.x] = hist_prefix[(int64_t)threadIdx.x * n_tiles + blockIdx.x] - start;
            int run = start;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const int c = S.woff[w][threadIdx.x];
                S.woff[w][threadIdx.x] = run;
                run += c;
            }
        }
    }
    __syncthreads();
    // sweep 2: stable reorder of the tile in shared memory
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        const int d = valid ? (int)((key[r] >> shift) & 0xff) : kDigits;
        const unsigned peers = __match_any_sync(kFull, d);
        const int start = S.woff[warp][d];
        __syncwarp();
        if ((peers & lt) == 0) S.woff[warp][d] = start + __popc(peers);
        __syncwarp();
        if (valid) {
            const int pos = start + __popc(peers & lt);
            S.key[pos] = key[r];
            S.val[pos] = vals[idx];
        }
    }
    __syncthreads();
    // write the reordered tile front to back: runs of one digit go to consecutive global addresses
    for (int i = threadIdx.x; i < tile_n; i += kSortThreads) {
        const uint32_t k = S.key[i];
        const int64_t dst = S.gbase[(k >> shift) & 0xff] + i;
        keys_out[dst] = k;
        vals_out[dst] = S.val[i];
    }
}

int radix_sort_pairs(uint32_t *k0, void *v0, uint32_t *k1, void *v1, int64_t n, int bits,
                     int value_bytes, DevBuf &scratch_hist, DevBuf &scratch_scan,
                     DevBuf &scratch_scan2, cudaStream_t stream, bool *result_in_second,
                     int64_t *launches)
{
    *result_in_second = false;
    if (n <= 1 || bits <= 0) return ARCTE_OK;
    const int64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    const int64_t hist_n = n_tiles * kDigits;
    ARCTE_TRY(dev_reserve(scratch_hist, (size_t)hist_n * sizeof(int32_t)));
    ARCTE_TRY(dev_reserve(scratch_scan, (size_t)(hist_n + 1) * sizeof(int64_t)));
    // more than 48 KB of dynamic shared memory needs the opt-in (per device: set on every call, it is cheap)
    if (value_bytes == 4)
        ARCTE_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(SortShared<uint32_t>)));
    else
        ARCTE_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(SortShared<uint64_t>)));
    uint32_t *ki = k0, *ko = k1;
    void *vi = v0, *vo = v1;
    bool second = false;
    for (int shift = 0; shift < bits; shift += 8) {
        k_radix_hist<<<(unsigned)n_tiles, kSortThreads, 0, stream>>>(ki, n, shift, n_tiles,
                                                                      scratch_hist.as<int32_t>());
        ++*launches;
        ARCTE_TRY(exclusive_scan_i32(scratch_hist.as<int32_t>(), scratch_scan.as<int64_t>(),
                                     hist_n, scratch_scan2, stream, launches));
        if (value_bytes == 4)
            k_radix_scatter<uint32_t><<<(unsigned)n_tiles, kSortThreads, sizeof(SortShared<uint32_t>), stream>>>(
                ki, (const uint32_t *)vi, n, shift, n_tiles, scratch_scan.as<int64_t>(), ko,
                (uint32_t *)vo);
        else
            k_radix_scatter<uint64_t><<<(unsigned)n_tiles, kSortThreads, sizeof(SortShared<uint64_t>), stream>>>(
                ki, (const uint64_t *)vi, n, shift, n_tiles, scratch_scan.as<int64_t>(), ko,
                (uint64_t *)vo);
        ++*launches;
        ARCTE_CUDA_TRY(cudaGetLastError());
        uint32_t *tk = ki; ki = ko; ko = tk;
        void *tv = vi; vi = vo; vo = tv;
        second = !second;
    }
    *result_in_second = second;
    return ARCTE_OK;
}

}  // namespace arcte
