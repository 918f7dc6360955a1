// assemble.cu -- K5: pack the per-seed member segments column-major, transpose them to
// row-major with a stable radix sort, and splice [I + pattern(A) | local] into one
// canonical n x 2n CSR.
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   local block   COO(ones,(member, seed)) -> CSR, summed over chunks/workers
//                                                     embedding/arcte/arcte.py:379-386, 670-673
//   base block    identity + binarised adjacency      embedding/arcte/arcte.py:676-679
//   hstack        sparse.hstack([base, local]).tocsr() embedding/arcte/arcte.py:683
//
// The output is what scipy returns: sorted column indices per row, no duplicates,
// float64 data (1.0; 2.0 on the diagonal of a node with a self loop), column id of a
// community = n + seed node id (arcte.py:376).
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "common.cuh"
#include "primitives.cuh"

namespace arcte {

struct SegPart {
    int64_t n_segments;
    const int32_t *seg_seed;
    const int32_t *seg_count;
    const int64_t *seg_offset;
    const int32_t *members;
};

__global__ void __launch_bounds__(256)
k_scatter_counts(SegPart part, int32_t *__restrict__ cnt_by_node)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= part.n_segments) return;
    const int32_t m = part.seg_count[k];
    if (m > 0) cnt_by_node[part.seg_seed[k]] = m;
}

// One warp per segment: members -> (row = member, col = seed) pairs laid out by
// ascending seed id, plus the per-row membership histogram.
__global__ void __launch_bounds__(256)
k_gather_segments(SegPart part, const int64_t *__restrict__ colptr, uint32_t *__restrict__ rows,
                  uint32_t *__restrict__ cols, int32_t *__restrict__ rowcnt)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= part.n_segments) return;
    const int32_t m = part.seg_count[k];
    if (m <= 0) return;
    const int32_t seed = part.seg_seed[k];
    const int64_t src = part.seg_offset[k];
    const int64_t dst = colptr[seed];
    for (int i = lane_id(); i < m; i += 32) {
        const int32_t x = part.members[src + i];
        rows[dst + i] = (uint32_t)x;
        cols[dst + i] = (uint32_t)seed;
        atomicAdd(&rowcnt[x], 1);
    }
}

// Row-sharded assembly: members of each segment that fall in the row range [lo, hi).
__global__ void __launch_bounds__(256)
k_count_in_range(SegPart part, int32_t lo, int32_t hi, int32_t *__restrict__ cnt_by_node)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= part.n_segments) return;
    const int32_t m = part.seg_count[k];
    if (m <= 0) return;
    const int64_t src = part.seg_offset[k];
    int c = 0;
    for (int i = lane_id(); i < m; i += 32) {
        const int32_t x = part.members[src + i];
        c += (x >= lo && x < hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if (lane_id() == 0 && c > 0) cnt_by_node[part.seg_seed[k]] = c;
}

// Same as k_gather_segments but keeps only rows in [lo, hi), stored relative to lo, in
// segment order (ballot compaction).
__global__ void __launch_bounds__(256)
k_gather_segments_range(SegPart part, int32_t lo, int32_t hi, const int64_t *__restrict__ colptr,
                        uint32_t *__restrict__ rows, uint32_t *__restrict__ cols, int32_t *__restrict__ rowcnt)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= part.n_segments) return;
    const int32_t m = part.seg_count[k];
    if (m <= 0) return;
    const int32_t seed = part.seg_seed[k];
    const int64_t src = part.seg_offset[k];
    int64_t dst = colptr[seed];
    const unsigned lt = lanemask_lt();
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane_id();
        int32_t x = -1;
        if (i < m) x = part.members[src + i];
        const bool in = x >= lo && x < hi;
        const unsigned mask = __ballot_sync(kFull, in);
        if (in) {
            const int64_t d = dst + __popc(mask & lt);
            rows[d] = (uint32_t)(x - lo);
            cols[d] = (uint32_t)seed;
            atomicAdd(&rowcnt[x - lo], 1);
        }
        dst += __popc(mask);
    }
}

// Length of each output row: |N(x)| + (1 unless x has a self loop) + memberships.
__global__ void __launch_bounds__(256)
k_row_lengths(int64_t nr, int64_t lo, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
              const int32_t *__restrict__ rowcnt, int32_t *__restrict__ base_len,
              int32_t *__restrict__ diag_pos, int64_t *__restrict__ total_len)
{
    const int64_t xl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xl >= nr) return;
    const int64_t x = xl + lo;
    const int64_t b = indptr[x], e = indptr[x + 1];
    // lower_bound of x in the sorted row
    int64_t lb = b, ub = e;
    while (lb < ub) {
        const int64_t mid = (lb + ub) >> 1;
        if (indices[mid] < x) lb = mid + 1;
        else ub = mid;
    }
    const bool self_loop = lb < e && indices[lb] == x;
    const int32_t bl = (int32_t)(e - b) + (self_loop ? 0 : 1);
    base_len[xl] = bl;
    diag_pos[xl] = self_loop ? -1 : (int32_t)(lb - b);  // where the identity entry is inserted
    total_len[xl] = (int64_t)bl + rowcnt[xl];
}

// Base block, one warp per row.
__global__ void __launch_bounds__(256)
k_fill_base(int64_t nr, int64_t lo, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
            const int32_t *__restrict__ diag_pos, const int64_t *__restrict__ out_indptr,
            int32_t *__restrict__ out_indices, double *__restrict__ out_data)
{
    const int64_t xl = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (xl >= nr) return;
    const int64_t x = xl + lo;
    const int64_t b = indptr[x], e = indptr[x + 1];
    const int64_t o = out_indptr[xl];
    const int32_t dp = diag_pos[xl];
    for (int64_t j = b + lane_id(); j < e; j += 32) {
        const int32_t c = indices[j];
        const int64_t t = j - b;
        const int64_t pos = o + t + ((dp >= 0 && t >= dp) ? 1 : 0);
        out_indices[pos] = c;
        out_data[pos] = (c == x) ? 2.0 : 1.0;  // I + ones(pattern): arcte.py:676-679
    }
    if (lane_id() == 0 && dp >= 0) {
        out_indices[o + dp] = (int32_t)x;
        out_data[o + dp] = 1.0;
    }
}

// Local block, one thread per (row, col) pair of the row-sorted list.
__global__ void __launch_bounds__(256)
k_fill_local(int64_t L, int64_t n, const uint32_t *__restrict__ rows, const uint32_t *__restrict__ cols,
             const int64_t *__restrict__ rowstart, const int32_t *__restrict__ base_len,
             const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices,
             double *__restrict__ out_data)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const uint32_t x = rows[i];
    const int64_t pos = out_indptr[x] + base_len[x] + (i - rowstart[x]);
    out_indices[pos] = (int32_t)(n + cols[i]);  // arcte.py:376: column = seed node id
    out_data[pos] = 1.0;
}

static inline unsigned grid_for(int64_t items, int block) { return (unsigned)((items + block - 1) / block); }
static int bit_length(uint64_t v)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

int assemble_parts(arcte_cuda_ctx *c, int n_parts, const SegPart *parts, int64_t row_lo, int64_t row_hi)
{
    if (!c->have_graph) { set_error("assemble: no graph resident"); return ARCTE_E_ARG; }
    const int64_t n = c->n;
    if (row_lo < 0 || row_hi > n || row_lo > row_hi) { set_error("assemble: bad row range"); return ARCTE_E_ARG; }
    const int64_t nr = row_hi - row_lo;  // rows of this block
    const bool all_rows = (row_lo == 0 && row_hi == n);
    cudaStream_t st = c->stream;
    int64_t *launches = &c->stats.launches;
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev0, st));
    // ARCTE_CUDA_DEBUG: host clock at the phase boundaries (the extra synchronisations exist only then)
    static const bool dbg = getenv("ARCTE_CUDA_DEBUG") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        if (!dbg) return;
        cudaStreamSynchronize(st);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[arcte] assemble: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };

    DevBuf &cnt_by_node = c->scratch[8];
    DevBuf &colptr = c->scratch[9];
    DevBuf &rowcnt = c->scratch[10];
    DevBuf &base_len = c->scratch[11];
    DevBuf &diag_pos = c->scratch[12];
    DevBuf &total_len = c->scratch[13];
    DevBuf &rowstart = c->scratch[14];
    ARCTE_TRY(dev_reserve(cnt_by_node, sizeof(int32_t) * (size_t)n));
    ARCTE_TRY(dev_reserve(colptr, sizeof(int64_t) * (size_t)(n + 1)));
    const size_t nr1 = (size_t)(nr > 0 ? nr : 1);
    ARCTE_TRY(dev_reserve(rowcnt, sizeof(int32_t) * nr1));
    ARCTE_TRY(dev_reserve(base_len, sizeof(int32_t) * nr1));
    ARCTE_TRY(dev_reserve(diag_pos, sizeof(int32_t) * nr1));
    ARCTE_TRY(dev_reserve(total_len, sizeof(int64_t) * nr1));
    ARCTE_TRY(dev_reserve(rowstart, sizeof(int64_t) * (nr1 + 1)));
    ARCTE_TRY(dev_reserve(c->out_indptr, sizeof(int64_t) * (nr1 + 1)));

    // 1. community size per seed node id -> column pointers of the local block
    ARCTE_CUDA_TRY(cudaMemsetAsync(cnt_by_node.p, 0, sizeof(int32_t) * (size_t)n, st));
    ARCTE_CUDA_TRY(cudaMemsetAsync(rowcnt.p, 0, sizeof(int32_t) * nr1, st));
    for (int p = 0; p < n_parts; ++p) {
        if (parts[p].n_segments == 0) continue;
        if (all_rows)
            k_scatter_counts<<<grid_for(parts[p].n_segments, 256), 256, 0, st>>>(parts[p], cnt_by_node.as<int32_t>());
        else
            k_count_in_range<<<grid_for(parts[p].n_segments * 32, 256), 256, 0, st>>>(
                parts[p], (int32_t)row_lo, (int32_t)row_hi, cnt_by_node.as<int32_t>());
        ++*launches;
    }
    ARCTE_TRY(exclusive_scan_i32(cnt_by_node.as<int32_t>(), colptr.as<int64_t>(), n, c->scratch[6], st, launches));
    int64_t L = 0;
    ARCTE_CUDA_TRY(cudaMemcpyAsync(&L, colptr.as<int64_t>() + n, sizeof(L), cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    phase("sizes + scan");

    // 2. pairs in column order, row histogram, stable sort by row
    const size_t L1 = (size_t)(L > 0 ? L : 1);
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(uint32_t) * L1));
    ARCTE_TRY(dev_reserve(c->scratch[1], sizeof(uint32_t) * L1));
    ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(uint32_t) * L1));
    ARCTE_TRY(dev_reserve(c->scratch[3], sizeof(uint32_t) * L1));
    for (int p = 0; p < n_parts; ++p) {
        if (parts[p].n_segments == 0) continue;
        if (all_rows)
            k_gather_segments<<<grid_for(parts[p].n_segments * 32, 256), 256, 0, st>>>(
                parts[p], colptr.as<int64_t>(), c->scratch[0].as<uint32_t>(), c->scratch[2].as<uint32_t>(),
                rowcnt.as<int32_t>());
        else
            k_gather_segments_range<<<grid_for(parts[p].n_segments * 32, 256), 256, 0, st>>>(
                parts[p], (int32_t)row_lo, (int32_t)row_hi, colptr.as<int64_t>(), c->scratch[0].as<uint32_t>(),
                c->scratch[2].as<uint32_t>(), rowcnt.as<int32_t>());
        ++*launches;
    }
    phase("scratch + gather segments");
    bool second = false;
    ARCTE_TRY(radix_sort_pairs(c->scratch[0].as<uint32_t>(), c->scratch[2].p, c->scratch[1].as<uint32_t>(),
                               c->scratch[3].p, L, bit_length((uint64_t)(nr > 0 ? nr - 1 : 0)), 4,
                               c->scratch[4], c->scratch[5], c->scratch[6], st, &second, launches));
    const uint32_t *rows = second ? c->scratch[1].as<uint32_t>() : c->scratch[0].as<uint32_t>();
    const uint32_t *cols = second ? c->scratch[3].as<uint32_t>() : c->scratch[2].as<uint32_t>();

    phase("radix sort by row");
    // 3. row lengths -> output row pointers
    k_row_lengths<<<grid_for(nr1, 256), 256, 0, st>>>(nr, row_lo, c->indptr.as<int64_t>(), c->indices.as<int32_t>(),
                                                    rowcnt.as<int32_t>(), base_len.as<int32_t>(),
                                                    diag_pos.as<int32_t>(), total_len.as<int64_t>());
    ++*launches;
    ARCTE_TRY(exclusive_scan_i64(total_len.as<int64_t>(), c->out_indptr.as<int64_t>(), nr, c->scratch[6], st, launches));
    ARCTE_TRY(exclusive_scan_i32(rowcnt.as<int32_t>(), rowstart.as<int64_t>(), nr, c->scratch[6], st, launches));
    int64_t nnz_out = 0;
    ARCTE_CUDA_TRY(cudaMemcpyAsync(&nnz_out, c->out_indptr.as<int64_t>() + nr, sizeof(nnz_out),
                                   cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    ARCTE_TRY(dev_reserve(c->out_indices, sizeof(int32_t) * (size_t)(nnz_out > 0 ? nnz_out : 1)));
    ARCTE_TRY(dev_reserve(c->out_data, sizeof(double) * (size_t)(nnz_out > 0 ? nnz_out : 1)));

    phase("row pointers + output allocation");
    // 4. fill
    k_fill_base<<<grid_for(nr1 * 32, 256), 256, 0, st>>>(nr, row_lo, c->indptr.as<int64_t>(), c->indices.as<int32_t>(),
                                                       diag_pos.as<int32_t>(), c->out_indptr.as<int64_t>(),
                                                       c->out_indices.as<int32_t>(), c->out_data.as<double>());
    ++*launches;
    if (L > 0) {
        k_fill_local<<<grid_for(L, 256), 256, 0, st>>>(L, n, rows, cols, rowstart.as<int64_t>(),
                                                       base_len.as<int32_t>(), c->out_indptr.as<int64_t>(),
                                                       c->out_indices.as<int32_t>(), c->out_data.as<double>());
        ++*launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev1, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    phase("fill base + local");
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.ms_assemble = ms;
    c->out_nnz = nnz_out;
    c->out_rows = nr;
    c->have_features = true;
    return ARCTE_OK;
}

}  // namespace arcte
