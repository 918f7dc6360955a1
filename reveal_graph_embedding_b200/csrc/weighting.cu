// weighting.cu -- the steps that follow the ARCTE path in every experiment of the reference
// (SURVEY.md section 8f rows 1 and 4): column normalisation and the chi2 / peak-SNR
// community weighting, on the device.
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   normalize_columns             embedding/common.py:49-67              (Python loop over all columns)
//   chi2_contingency_matrix       embedding/community_weighting.py:11-45
//   peak_snr_weight_aggregation   embedding/community_weighting.py:48-84 (Python loops over classes / columns)
//   community_weighting           embedding/community_weighting.py:87-125 (Python loop over all columns,
//                                 eliminate_zeros, sklearn l2 row normalisation)
//
// Arithmetic: every fp64 operation is one IEEE rounding in the reference's order (numpy's
// pairwise tree for np.var / np.mean, a left-to-right sum of squares for sklearn's row norm);
// counts are accumulated as integers or integer-valued doubles, so atomics cannot change a
// bit.  The only operations that are not bit-reproducible against numpy are the two log()
// calls (1 ulp, like epsilon-effective).
//
// All kernels are HBM streaming passes over the CSR (12-20 B per stored entry) or over the
// K x F contingency matrix (8-16 B per cell); none is GEMM-shaped.
#include "common.cuh"
#include "primitives.cuh"

namespace arcte {

static inline unsigned wgrid(int64_t items, int block) { return (unsigned)((items + block - 1) / block); }

// ---- document frequencies -----------------------------------------------------------------
// features.getcol(j).data.size (common.py:60, community_weighting.py:95,106): stored entries per
// column, explicit zeros included.
__global__ void k_col_hist(int64_t nnz, const int32_t *__restrict__ indices, int32_t *__restrict__ df)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&df[indices[k]], 1);
}

// common.py:61-63: divisor sqrt(log(df)) for columns with df > 1; 0 marks "leave alone".
__global__ void k_norm_scale(int64_t n_cols, const int32_t *__restrict__ df, double *__restrict__ scale)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_cols) scale[j] = df[j] > 1 ? sqrt(log((double)df[j])) : 0.0;
}

__global__ void k_apply_div(int64_t nnz, const int32_t *__restrict__ indices, const double *__restrict__ in,
                            const double *__restrict__ scale, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const double s = scale[indices[k]];
        const double v = in[k];
        out[k] = s != 0.0 ? __ddiv_rn(v, s) : v;
    }
}

// community_weighting.py:96-103: multiplier per column; 1.0 (an exact identity) where df <= 1.
__global__ void k_reinforcement(int64_t n_cols, const int32_t *__restrict__ df, const double *__restrict__ weights,
                                double *__restrict__ mul)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    double m = 1.0;
    if (df[j] > 1) m = weights[j] == 0.0 ? 0.0 : log(__dadd_rn(1.0, weights[j]));
    mul[j] = m;
}

// Entries that survive eliminate_zeros() (community_weighting.py:118-119), per row.  One warp per row.
__global__ void k_weighted_row_counts(int64_t n_rows, const int64_t *__restrict__ indptr,
                                      const int32_t *__restrict__ indices, const double *__restrict__ data,
                                      const double *__restrict__ mul, int32_t *__restrict__ row_nnz)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int lane = lane_id();
    int cnt = 0;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32)
        cnt += __dmul_rn(data[k], mul[indices[k]]) != 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if (lane == 0) row_nnz[row] = cnt;
}

// Scale, drop zeros, l2-normalise (sklearn _inplace_csr_row_normalize_l2: sum += x*x left to
// right, sqrt, x /= norm; rows with sum == 0 untouched).  One warp per row; the lanes load and
// square 32 entries at a time, the running sum is carried through them in order with shuffles
// so the additions happen exactly in the reference's sequence (a dropped zero adds +0.0, which
// changes nothing).
__global__ void k_weighted_row_fill(int64_t n_rows, const int64_t *__restrict__ indptr,
                                    const int32_t *__restrict__ indices, const double *__restrict__ data,
                                    const double *__restrict__ mul, const int64_t *__restrict__ out_indptr,
                                    int32_t *__restrict__ out_indices, double *__restrict__ out_data)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int64_t b = indptr[row], e = indptr[row + 1];
    const int64_t ob = out_indptr[row];
    int64_t o = ob;
    double sum = 0.0;
    for (int64_t k0 = b; k0 < e; k0 += 32) {
        const int64_t k = k0 + lane;
        int j = 0;
        double v = 0.0;
        if (k < e) {
            j = indices[k];
            v = __dmul_rn(data[k], mul[j]);
        }
        const bool keep = k < e && v != 0.0;
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) {
            const int64_t p = o + __popc(m & lt);
            out_indices[p] = j;
            out_data[p] = v;
        }
        o += __popc(m);
        const double sq = __dmul_rn(v, v);
        const int lim = (int)((e - k0) < 32 ? (e - k0) : 32);
        for (int l = 0; l < lim; ++l) sum = __dadd_rn(sum, __shfl_sync(kFull, sq, l));
    }
    if (sum != 0.0) {
        const double norm = sqrt(sum);
        __syncwarp();
        for (int64_t p = ob + lane; p < o; p += 32) out_data[p] = __ddiv_rn(out_data[p], norm);
    }
}

// ---- chi2 contingency -----------------------------------------------------------------------
// observed = Y.T @ pattern(X) (community_weighting.py:23), feature_count = pattern(X).sum(0) (:27),
// class sums for Y.mean(0) (:28).  One warp per training row.  All addends are integer-valued,
// so the fp64 atomics are exact and order-independent.
__global__ void k_chi2_accumulate(int64_t n_rows, int64_t n_cols, const int64_t *__restrict__ x_indptr,
                                  const int32_t *__restrict__ x_indices, const int64_t *__restrict__ y_indptr,
                                  const int32_t *__restrict__ y_indices, const double *__restrict__ y_data,
                                  double *__restrict__ observed, double *__restrict__ feature_count,
                                  double *__restrict__ class_sum)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int lane = lane_id();
    const int64_t xb = x_indptr[row], xe = x_indptr[row + 1];
    const int64_t yb = y_indptr[row], ye = y_indptr[row + 1];
    for (int64_t k = xb + lane; k < xe; k += 32) atomicAdd(&feature_count[x_indices[k]], 1.0);
    for (int64_t q = yb; q < ye; ++q) {
        const int64_t c = y_indices[q];
        const double y = y_data[q];
        if (lane == 0) atomicAdd(&class_sum[c], y);
        double *__restrict__ orow = observed + c * n_cols;
        for (int64_t k = xb + lane; k < xe; k += 32) atomicAdd(&orow[x_indices[k]], y);
    }
}

// (observed - expected)^2 / expected, expected == 0 -> 1 (community_weighting.py:29-42).
__global__ void k_chi2_finish(int64_t K, int64_t F, int64_t n_rows, double *__restrict__ cm,
                              const double *__restrict__ feature_count, const double *__restrict__ class_sum)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double fc = feature_count[f];
    for (int64_t c = 0; c < K; ++c) {
        const double prob = __ddiv_rn(class_sum[c], (double)n_rows);
        double expected = __dmul_rn(prob, fc);
        double d = __dadd_rn(cm[c * F + f], -expected);
        d = __dmul_rn(d, d);
        if (expected == 0.0) expected = 1.0;
        cm[c * F + f] = __ddiv_rn(d, expected);
    }
}

// ---- peak SNR -----------------------------------------------------------------------------
__global__ void k_nan_to_zero(int64_t n, double *__restrict__ a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (isnan(a[i])) a[i] = 0.0;  // community_weighting.py:49
}

// np.var of every class row (community_weighting.py:52-54): numpy's pairwise tree twice --
// mean = sum/F, then sum((a-mean)^2)/F.  One CTA of 1024 threads per class: the top ten levels
// of numpy's split tree are spread over the threads, every thread streams its own contiguous
// subtree, the partial sums are folded back in tree order through shared memory.
__global__ void __launch_bounds__(1024, 1)
k_row_variance(int64_t K, int64_t F, const double *__restrict__ cm, double *__restrict__ variance)
{
    __shared__ double red[1024];
    const int64_t row = blockIdx.x;
    if (row >= K) return;
    const double *__restrict__ a = cm + row * F;
    auto at = [a](int64_t i) { return a[i]; };
    const double mean = __ddiv_rn(pairwise_sum_block1024(at, 0, F, red), (double)F);
    auto at_sq = [a, mean](int64_t i) {
        const double x = __dadd_rn(a[i], -mean);
        return __dmul_rn(x, x);
    };
    const double var = __ddiv_rn(pairwise_sum_block1024(at_sq, 0, F, red), (double)F);
    if (threadIdx.x == 0) variance[row] = var;
}

// community_weighting.py:55-67: noise = sqrt(mean(variance)); per column the spread of its
// positive entries over the noise.
__global__ void k_psnr_weights(int64_t K, int64_t F, const double *__restrict__ cm,
                               const double *__restrict__ variance, double *__restrict__ weights)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    auto at = [variance](int64_t i) { return variance[i]; };
    const double noise = sqrt(__ddiv_rn(pairwise_sum(at, 0, K), (double)K));
    int cnt = 0;
    double mx = 0.0, mn = 0.0;
    for (int64_t c = 0; c < K; ++c) {
        const double v = cm[c * F + f];
        if (v > 0.0) {
            if (cnt == 0) { mx = v; mn = v; }
            else { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
            ++cnt;
        }
    }
    double w = 0.0;
    if (cnt > 1) w = __ddiv_rn(__dadd_rn(mx, -mn), noise);
    else if (cnt == 1) w = __ddiv_rn(mx, noise);
    weights[f] = w;
}

// ---- host drivers ---------------------------------------------------------------------------
static int check_csr_args(const char *who, int64_t n_rows, int64_t n_cols, const int64_t *indptr, const void *indices)
{
    if (n_rows < 0 || n_cols < 0 || !indptr || n_cols >= (int64_t(1) << 31)) {
        set_error(std::string(who) + ": bad shape or null indptr");
        return ARCTE_E_ARG;
    }
    if (indptr[0] != 0 || indptr[n_rows] < 0 || (indptr[n_rows] > 0 && !indices)) {
        set_error(std::string(who) + ": indptr[0] must be 0 and indices must be given");
        return ARCTE_E_ARG;
    }
    return ARCTE_OK;
}

static int device_col_hist(arcte_cuda_ctx *c, int64_t n_cols, int64_t nnz, const int32_t *dev_indices, DevBuf &df)
{
    ARCTE_TRY(dev_reserve(df, sizeof(int32_t) * (size_t)(n_cols > 0 ? n_cols : 1)));
    ARCTE_CUDA_TRY(cudaMemsetAsync(df.p, 0, sizeof(int32_t) * (size_t)(n_cols > 0 ? n_cols : 1), c->stream));
    if (nnz > 0) {
        unsigned g = wgrid(nnz, 256);
        const unsigned cap = (unsigned)c->sm_count * 16;
        if (g > cap) g = cap;
        k_col_hist<<<g, 256, 0, c->stream>>>(nnz, dev_indices, df.as<int32_t>());
        ++c->stats.launches;
    }
    return ARCTE_OK;
}

// data_out = data_in with every column of document frequency > 1 divided by sqrt(log(df)).
// dev_data_in may equal dev_data_out.
int normalize_columns_device(arcte_cuda_ctx *c, int64_t n_cols, int64_t nnz, const int32_t *dev_indices,
                             const double *dev_data_in, double *dev_data_out)
{
    ARCTE_TRY(device_col_hist(c, n_cols, nnz, dev_indices, c->scratch[8]));
    ARCTE_TRY(dev_reserve(c->scratch[9], sizeof(double) * (size_t)(n_cols > 0 ? n_cols : 1)));
    if (n_cols > 0) {
        k_norm_scale<<<wgrid(n_cols, 256), 256, 0, c->stream>>>(n_cols, c->scratch[8].as<int32_t>(),
                                                               c->scratch[9].as<double>());
        ++c->stats.launches;
    }
    if (nnz > 0) {
        unsigned g = wgrid(nnz, 256);
        const unsigned cap = (unsigned)c->sm_count * 16;
        if (g > cap) g = cap;
        k_apply_div<<<g, 256, 0, c->stream>>>(nnz, dev_indices, dev_data_in, c->scratch[9].as<double>(), dev_data_out);
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

// K x F contingency matrix in `cm` (device, overwritten) from device CSR inputs.
static int chi2_device(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *x_indptr,
                       const int32_t *x_indices, int64_t K, const int64_t *y_indptr, const int32_t *y_indices,
                       const double *y_data, double *cm)
{
    cudaStream_t st = c->stream;
    ARCTE_TRY(dev_reserve(c->scratch[8], sizeof(double) * (size_t)(n_cols > 0 ? n_cols : 1)));
    ARCTE_TRY(dev_reserve(c->scratch[9], sizeof(double) * (size_t)(K > 0 ? K : 1)));
    ARCTE_CUDA_TRY(cudaMemsetAsync(cm, 0, sizeof(double) * (size_t)K * (size_t)n_cols, st));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->scratch[8].p, 0, sizeof(double) * (size_t)(n_cols > 0 ? n_cols : 1), st));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->scratch[9].p, 0, sizeof(double) * (size_t)(K > 0 ? K : 1), st));
    if (n_rows > 0) {
        k_chi2_accumulate<<<wgrid(n_rows * 32, 256), 256, 0, st>>>(n_rows, n_cols, x_indptr, x_indices, y_indptr,
                                                                   y_indices, y_data, cm, c->scratch[8].as<double>(),
                                                                   c->scratch[9].as<double>());
        ++c->stats.launches;
    }
    if (n_cols > 0 && K > 0) {
        k_chi2_finish<<<wgrid(n_cols, 256), 256, 0, st>>>(K, n_cols, n_rows, cm, c->scratch[8].as<double>(),
                                                          c->scratch[9].as<double>());
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

// weights[F] from the device-resident K x F matrix (nan -> 0 in place).
static int peak_snr_device(arcte_cuda_ctx *c, int64_t K, int64_t F, double *cm, double *weights)
{
    cudaStream_t st = c->stream;
    if (K <= 0 || F <= 0) return ARCTE_OK;
    ARCTE_TRY(dev_reserve(c->scratch[10], sizeof(double) * (size_t)K));
    unsigned g = wgrid(K * F, 256);
    const unsigned cap = (unsigned)c->sm_count * 16;
    if (g > cap) g = cap;
    k_nan_to_zero<<<g, 256, 0, st>>>(K * F, cm);
    k_row_variance<<<(unsigned)K, 1024, 0, st>>>(K, F, cm, c->scratch[10].as<double>());
    k_psnr_weights<<<wgrid(F, 256), 256, 0, st>>>(K, F, cm, c->scratch[10].as<double>(), weights);
    c->stats.launches += 3;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

static int upload(arcte_cuda_ctx *c, DevBuf &b, const void *host, size_t bytes)
{
    ARCTE_TRY(dev_reserve(b, bytes));
    if (bytes > 0) ARCTE_TRY(copy_from_host(c, b.p, host, bytes));   // pageable scipy arrays at PCIe rate (hostcopy.cu)
    return ARCTE_OK;
}

// community_weighting for one device-resident CSR: outputs into the given buffers (grown as
// needed), *kept = entries that survive eliminate_zeros().  weights: one double per column.
static int community_weighting_device(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                      const int64_t *indptr, const int32_t *indices, const double *data,
                                      const double *weights, DevBuf &out_indptr, DevBuf &out_indices,
                                      DevBuf &out_data, int64_t *kept)
{
    cudaStream_t st = c->stream;
    ARCTE_TRY(device_col_hist(c, n_cols, nnz, indices, c->scratch[8]));
    ARCTE_TRY(dev_reserve(c->scratch[9], sizeof(double) * (size_t)n_cols));
    k_reinforcement<<<wgrid(n_cols, 256), 256, 0, st>>>(n_cols, c->scratch[8].as<int32_t>(), weights,
                                                        c->scratch[9].as<double>());
    ARCTE_TRY(dev_reserve(c->scratch[10], sizeof(int32_t) * (size_t)n_rows));
    k_weighted_row_counts<<<wgrid(n_rows * 32, 256), 256, 0, st>>>(n_rows, indptr, indices, data,
                                                                   c->scratch[9].as<double>(),
                                                                   c->scratch[10].as<int32_t>());
    c->stats.launches += 2;
    ARCTE_TRY(dev_reserve(out_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ARCTE_TRY(exclusive_scan_i32(c->scratch[10].as<int32_t>(), out_indptr.as<int64_t>(), n_rows, c->scratch[0], st,
                                 &c->stats.launches));
    ARCTE_TRY(dev_reserve(out_indices, sizeof(int32_t) * (size_t)nnz));
    ARCTE_TRY(dev_reserve(out_data, sizeof(double) * (size_t)nnz));
    k_weighted_row_fill<<<wgrid(n_rows * 32, 256), 256, 0, st>>>(n_rows, indptr, indices, data,
                                                                 c->scratch[9].as<double>(), out_indptr.as<int64_t>(),
                                                                 out_indices.as<int32_t>(), out_data.as<double>());
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    ARCTE_CUDA_TRY(cudaMemcpyAsync(kept, out_indptr.as<int64_t>() + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    return ARCTE_OK;
}

// X[rows, :] of a device-resident CSR (experiments/utility.py:94-97) into the context's gather
// buffers: row lengths -> exclusive scan -> one warp copies each row.
__global__ void k_gather_lengths(int64_t n, const int64_t *__restrict__ rows, const int64_t *__restrict__ indptr,
                                 int32_t *__restrict__ len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = (int32_t)(indptr[rows[i] + 1] - indptr[rows[i]]);
}
__global__ void k_gather_rows(int64_t n, const int64_t *__restrict__ rows, const int64_t *__restrict__ indptr,
                              const int32_t *__restrict__ indices, const double *__restrict__ data,
                              const int64_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices,
                              double *__restrict__ out_data)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int64_t b = indptr[rows[i]], e = indptr[rows[i] + 1], o = out_indptr[i];
    for (int64_t k = b + lane_id(); k < e; k += 32) {
        out_indices[o + (k - b)] = indices[k];
        out_data[o + (k - b)] = data[k];
    }
}
static int gather_rows_device(arcte_cuda_ctx *c, int64_t n, const int64_t *host_rows, const int64_t *indptr,
                              const int32_t *indices, const double *data, int64_t *nnz_out)
{
    cudaStream_t st = c->stream;
    *nnz_out = 0;
    ARCTE_TRY(dev_reserve(c->fg_indptr, sizeof(int64_t) * (size_t)(n + 1)));
    if (n == 0) {
        ARCTE_CUDA_TRY(cudaMemsetAsync(c->fg_indptr.p, 0, sizeof(int64_t), st));
        return ARCTE_OK;
    }
    ARCTE_TRY(upload(c, c->fg_rows, host_rows, sizeof(int64_t) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->scratch[10], sizeof(int32_t) * (size_t)n));
    k_gather_lengths<<<wgrid(n, 256), 256, 0, st>>>(n, c->fg_rows.as<int64_t>(), indptr, c->scratch[10].as<int32_t>());
    ++c->stats.launches;
    ARCTE_TRY(exclusive_scan_i32(c->scratch[10].as<int32_t>(), c->fg_indptr.as<int64_t>(), n, c->scratch[0], st,
                                 &c->stats.launches));
    int64_t nnz = 0;
    ARCTE_CUDA_TRY(cudaMemcpyAsync(&nnz, c->fg_indptr.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    ARCTE_TRY(dev_reserve(c->fg_indices, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1)));
    ARCTE_TRY(dev_reserve(c->fg_data, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    if (nnz > 0) {
        k_gather_rows<<<wgrid(n * 32, 256), 256, 0, st>>>(n, c->fg_rows.as<int64_t>(), indptr, indices, data,
                                                          c->fg_indptr.as<int64_t>(), c->fg_indices.as<int32_t>(),
                                                          c->fg_data.as<double>());
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    *nnz_out = nnz;
    return ARCTE_OK;
}

}  // namespace arcte

using namespace arcte;

#define CHECK_CTX(ctx)                                   \
    do {                                                 \
        if (!(ctx)) {                                    \
            set_error("null context");                   \
            return ARCTE_E_ARG;                          \
        }                                                \
        ARCTE_CUDA_TRY(cudaSetDevice((ctx)->device));    \
    } while (0)

extern "C" {

int arcte_cuda_normalize_columns(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                                 const int32_t *host_indices, const double *host_data_in, double *host_data_out)
{
    CHECK_CTX(c);
    ARCTE_TRY(check_csr_args("normalize_columns", n_rows, n_cols, host_indptr, host_indices));
    const int64_t nnz = host_indptr[n_rows];
    if (nnz == 0) return ARCTE_OK;
    if (!host_data_in || !host_data_out) { set_error("normalize_columns: null data"); return ARCTE_E_ARG; }
    c->stats.launches = 0;
    ARCTE_TRY(upload(c, c->scratch[11], host_indices, sizeof(int32_t) * (size_t)nnz));
    ARCTE_TRY(upload(c, c->scratch[12], host_data_in, sizeof(double) * (size_t)nnz));
    ARCTE_TRY(normalize_columns_device(c, n_cols, nnz, c->scratch[11].as<int32_t>(), c->scratch[12].as<double>(),
                                       c->scratch[12].as<double>()));
    ARCTE_TRY(copy_to_host(c, host_data_out, c->scratch[12].p, sizeof(double) * (size_t)nnz));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_normalize_features(arcte_cuda_ctx *c)
{
    CHECK_CTX(c);
    if (!c->have_features) { set_error("normalize_features: nothing assembled"); return ARCTE_E_ARG; }
    if (c->out_rows != c->n) {
        set_error("normalize_features: only a row block is resident; document frequencies need every row");
        return ARCTE_E_ARG;
    }
    c->stats.launches = 0;
    ARCTE_TRY(normalize_columns_device(c, 2 * c->n, c->out_nnz, c->out_indices.as<int32_t>(),
                                       c->out_data.as<double>(), c->out_data.as<double>()));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

static int upload_xy(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *x_indptr,
                     const int32_t *x_indices, int64_t K, const int64_t *y_indptr, const int32_t *y_indices,
                     const double *y_data)
{
    ARCTE_TRY(check_csr_args("chi2: X", n_rows, n_cols, x_indptr, x_indices));
    ARCTE_TRY(check_csr_args("chi2: Y", n_rows, K, y_indptr, y_indices));
    if (y_indptr[n_rows] > 0 && !y_data) { set_error("chi2: null label data"); return ARCTE_E_ARG; }
    ARCTE_TRY(upload(c, c->scratch[11], x_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ARCTE_TRY(upload(c, c->scratch[12], x_indices, sizeof(int32_t) * (size_t)x_indptr[n_rows]));
    ARCTE_TRY(upload(c, c->scratch[13], y_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ARCTE_TRY(upload(c, c->scratch[14], y_indices, sizeof(int32_t) * (size_t)y_indptr[n_rows]));
    ARCTE_TRY(upload(c, c->scratch[15], y_data, sizeof(double) * (size_t)y_indptr[n_rows]));
    return ARCTE_OK;
}

int arcte_cuda_chi2_contingency(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *host_x_indptr,
                                const int32_t *host_x_indices, int64_t n_classes, const int64_t *host_y_indptr,
                                const int32_t *host_y_indices, const double *host_y_data, double *host_out)
{
    CHECK_CTX(c);
    if (n_classes <= 0 || !host_out) { set_error("chi2: bad class count or null output"); return ARCTE_E_ARG; }
    c->stats.launches = 0;
    ARCTE_TRY(upload_xy(c, n_rows, n_cols, host_x_indptr, host_x_indices, n_classes, host_y_indptr, host_y_indices,
                        host_y_data));
    if (n_cols == 0) return ARCTE_OK;
    ARCTE_TRY(dev_reserve(c->scratch[7], sizeof(double) * (size_t)n_classes * (size_t)n_cols));
    ARCTE_TRY(chi2_device(c, n_rows, n_cols, c->scratch[11].as<int64_t>(), c->scratch[12].as<int32_t>(), n_classes,
                          c->scratch[13].as<int64_t>(), c->scratch[14].as<int32_t>(), c->scratch[15].as<double>(),
                          c->scratch[7].as<double>()));
    ARCTE_TRY(copy_to_host(c, host_out, c->scratch[7].p, sizeof(double) * (size_t)n_classes * (size_t)n_cols));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_peak_snr(arcte_cuda_ctx *c, int64_t n_classes, int64_t n_cols, double *host_cm_inout,
                        double *host_weights_out)
{
    CHECK_CTX(c);
    if (n_classes <= 0 || n_cols < 0 || !host_cm_inout || !host_weights_out) {
        set_error("peak_snr: bad arguments");
        return ARCTE_E_ARG;
    }
    if (n_cols == 0) return ARCTE_OK;
    c->stats.launches = 0;
    const size_t bytes = sizeof(double) * (size_t)n_classes * (size_t)n_cols;
    ARCTE_TRY(upload(c, c->scratch[7], host_cm_inout, bytes));
    ARCTE_TRY(dev_reserve(c->scratch[6], sizeof(double) * (size_t)n_cols));
    ARCTE_TRY(peak_snr_device(c, n_classes, n_cols, c->scratch[7].as<double>(), c->scratch[6].as<double>()));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(host_cm_inout, c->scratch[7].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    ARCTE_TRY(copy_to_host(c, host_weights_out, c->scratch[6].p, sizeof(double) * (size_t)n_cols));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_chi2_psnr_weights(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *host_x_indptr,
                                 const int32_t *host_x_indices, int64_t n_classes, const int64_t *host_y_indptr,
                                 const int32_t *host_y_indices, const double *host_y_data, double *host_weights_out)
{
    CHECK_CTX(c);
    if (n_classes <= 0 || !host_weights_out) { set_error("chi2_psnr: bad class count or null output"); return ARCTE_E_ARG; }
    c->stats.launches = 0;
    ARCTE_TRY(upload_xy(c, n_rows, n_cols, host_x_indptr, host_x_indices, n_classes, host_y_indptr, host_y_indices,
                        host_y_data));
    if (n_cols == 0) return ARCTE_OK;
    ARCTE_TRY(dev_reserve(c->scratch[7], sizeof(double) * (size_t)n_classes * (size_t)n_cols));
    ARCTE_TRY(dev_reserve(c->scratch[6], sizeof(double) * (size_t)n_cols));
    ARCTE_TRY(chi2_device(c, n_rows, n_cols, c->scratch[11].as<int64_t>(), c->scratch[12].as<int32_t>(), n_classes,
                          c->scratch[13].as<int64_t>(), c->scratch[14].as<int32_t>(), c->scratch[15].as<double>(),
                          c->scratch[7].as<double>()));
    ARCTE_TRY(peak_snr_device(c, n_classes, n_cols, c->scratch[7].as<double>(), c->scratch[6].as<double>()));
    ARCTE_TRY(copy_to_host(c, host_weights_out, c->scratch[6].p, sizeof(double) * (size_t)n_cols));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_community_weighting(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                                   const int32_t *host_indices, const double *host_data,
                                   const double *host_weights, int64_t *host_out_indptr, int32_t *host_out_indices,
                                   double *host_out_data, int64_t *out_nnz)
{
    CHECK_CTX(c);
    ARCTE_TRY(check_csr_args("community_weighting", n_rows, n_cols, host_indptr, host_indices));
    if (!host_out_indptr || !out_nnz || (n_cols > 0 && !host_weights)) {
        set_error("community_weighting: null argument");
        return ARCTE_E_ARG;
    }
    const int64_t nnz = host_indptr[n_rows];
    *out_nnz = 0;
    if (nnz > 0 && (!host_data || !host_out_indices || !host_out_data)) {
        set_error("community_weighting: null data");
        return ARCTE_E_ARG;
    }
    cudaStream_t st = c->stream;
    c->stats.launches = 0;
    if (n_rows == 0 || nnz == 0) {
        for (int64_t i = 0; i <= n_rows; ++i) host_out_indptr[i] = 0;
        return ARCTE_OK;
    }
    ARCTE_TRY(upload(c, c->scratch[11], host_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ARCTE_TRY(upload(c, c->scratch[12], host_indices, sizeof(int32_t) * (size_t)nnz));
    ARCTE_TRY(upload(c, c->scratch[13], host_data, sizeof(double) * (size_t)nnz));
    ARCTE_TRY(upload(c, c->scratch[14], host_weights, sizeof(double) * (size_t)n_cols));
    int64_t kept = 0;
    ARCTE_TRY(community_weighting_device(c, n_rows, n_cols, nnz, c->scratch[11].as<int64_t>(),
                                         c->scratch[12].as<int32_t>(), c->scratch[13].as<double>(),
                                         c->scratch[14].as<double>(), c->scratch[15], c->scratch[6], c->scratch[7], &kept));
    ARCTE_TRY(copy_to_host(c, host_out_indptr, c->scratch[15].p, sizeof(int64_t) * (size_t)(n_rows + 1)));
    if (kept > 0) {
        ARCTE_TRY(copy_to_host(c, host_out_indices, c->scratch[6].p, sizeof(int32_t) * (size_t)kept));
        ARCTE_TRY(copy_to_host(c, host_out_data, c->scratch[7].p, sizeof(double) * (size_t)kept));
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    *out_nnz = kept;
    return ARCTE_OK;
}

// ---- feature matrix resident in HBM across the folds of an experiment --------------------------
int arcte_cuda_store_features(arcte_cuda_ctx *c, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                              const int32_t *host_indices, const double *host_data)
{
    CHECK_CTX(c);
    ARCTE_TRY(check_csr_args("store_features", n_rows, n_cols, host_indptr, host_indices));
    const int64_t nnz = host_indptr[n_rows];
    if (nnz > 0 && !host_data) { set_error("store_features: null data"); return ARCTE_E_ARG; }
    c->fs_valid = c->fo_valid = false;
    ARCTE_TRY(upload(c, c->fs_indptr, host_indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ARCTE_TRY(upload(c, c->fs_indices, host_indices, sizeof(int32_t) * (size_t)nnz));
    ARCTE_TRY(upload(c, c->fs_data, host_data, sizeof(double) * (size_t)nnz));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->fs_rows = n_rows; c->fs_cols = n_cols; c->fs_nnz = nnz;
    c->fs_valid = true;
    return ARCTE_OK;
}

int arcte_cuda_store_assembled(arcte_cuda_ctx *c)
{
    CHECK_CTX(c);
    if (!c->have_features || c->out_rows != c->n) {
        set_error("store_assembled: no complete feature matrix is resident (call assemble for all rows first)");
        return ARCTE_E_ARG;
    }
    c->fs_valid = c->fo_valid = false;
    const size_t nnz = (size_t)c->out_nnz;
    ARCTE_TRY(dev_reserve(c->fs_indptr, sizeof(int64_t) * (size_t)(c->n + 1)));
    ARCTE_TRY(dev_reserve(c->fs_indices, sizeof(int32_t) * (nnz ? nnz : 1)));
    ARCTE_TRY(dev_reserve(c->fs_data, sizeof(double) * (nnz ? nnz : 1)));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->fs_indptr.p, c->out_indptr.p, sizeof(int64_t) * (size_t)(c->n + 1),
                                   cudaMemcpyDeviceToDevice, c->stream));
    if (nnz) {
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->fs_indices.p, c->out_indices.p, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, c->stream));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->fs_data.p, c->out_data.p, sizeof(double) * nnz, cudaMemcpyDeviceToDevice, c->stream));
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->fs_rows = c->n; c->fs_cols = 2 * c->n; c->fs_nnz = c->out_nnz;
    c->fs_valid = true;
    return ARCTE_OK;
}

int arcte_cuda_weighted_fold(arcte_cuda_ctx *c, int64_t n_train, const int64_t *host_train_rows, int64_t n_test,
                             const int64_t *host_test_rows, int64_t n_classes, const int64_t *host_y_indptr,
                             const int32_t *host_y_indices, const double *host_y_data, int64_t *train_nnz,
                             int64_t *test_nnz)
{
    CHECK_CTX(c);
    if (!c->fs_valid) { set_error("weighted_fold: no feature matrix stored (store_features / store_assembled)"); return ARCTE_E_ARG; }
    if (n_train < 0 || n_test < 0 || n_classes <= 0 || (n_train > 0 && !host_train_rows) ||
        (n_test > 0 && !host_test_rows) || !host_y_indptr) {
        set_error("weighted_fold: bad arguments");
        return ARCTE_E_ARG;
    }
    for (int64_t i = 0; i < n_train; ++i)
        if (host_train_rows[i] < 0 || host_train_rows[i] >= c->fs_rows) { set_error("weighted_fold: training row out of range"); return ARCTE_E_ARG; }
    for (int64_t i = 0; i < n_test; ++i)
        if (host_test_rows[i] < 0 || host_test_rows[i] >= c->fs_rows) { set_error("weighted_fold: test row out of range"); return ARCTE_E_ARG; }
    ARCTE_TRY(check_csr_args("weighted_fold: Y", n_train, n_classes, host_y_indptr, host_y_indices));
    if (host_y_indptr[n_train] > 0 && !host_y_data) { set_error("weighted_fold: null label data"); return ARCTE_E_ARG; }
    cudaStream_t st = c->stream;
    c->stats.launches = 0;
    c->fo_valid = false;
    const int64_t F = c->fs_cols;
    const int64_t *fsp = c->fs_indptr.as<int64_t>();
    const int32_t *fsi = c->fs_indices.as<int32_t>();
    const double *fsd = c->fs_data.as<double>();
    // ---- training block: gather, chi2 + peak SNR weights (kept on the device), weighting ----
    int64_t g_nnz = 0;
    ARCTE_TRY(gather_rows_device(c, n_train, host_train_rows, fsp, fsi, fsd, &g_nnz));
    ARCTE_TRY(upload(c, c->scratch[13], host_y_indptr, sizeof(int64_t) * (size_t)(n_train + 1)));
    ARCTE_TRY(upload(c, c->scratch[14], host_y_indices, sizeof(int32_t) * (size_t)host_y_indptr[n_train]));
    ARCTE_TRY(upload(c, c->scratch[15], host_y_data, sizeof(double) * (size_t)host_y_indptr[n_train]));
    ARCTE_TRY(dev_reserve(c->scratch[7], sizeof(double) * (size_t)n_classes * (size_t)(F > 0 ? F : 1)));
    ARCTE_TRY(dev_reserve(c->scratch[6], sizeof(double) * (size_t)(F > 0 ? F : 1)));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->scratch[6].p, 0, sizeof(double) * (size_t)(F > 0 ? F : 1), st));
    if (F > 0) {
        ARCTE_TRY(chi2_device(c, n_train, F, c->fg_indptr.as<int64_t>(), c->fg_indices.as<int32_t>(), n_classes,
                              c->scratch[13].as<int64_t>(), c->scratch[14].as<int32_t>(), c->scratch[15].as<double>(),
                              c->scratch[7].as<double>()));
        ARCTE_TRY(peak_snr_device(c, n_classes, F, c->scratch[7].as<double>(), c->scratch[6].as<double>()));
    }
    for (int which = 0; which < 2; ++which) {
        const int64_t rows = which == 0 ? n_train : n_test;
        if (which == 1) ARCTE_TRY(gather_rows_device(c, n_test, host_test_rows, fsp, fsi, fsd, &g_nnz));
        int64_t kept = 0;
        ARCTE_TRY(dev_reserve(c->fo_indptr[which], sizeof(int64_t) * (size_t)(rows + 1)));
        if (rows == 0 || g_nnz == 0) {
            ARCTE_CUDA_TRY(cudaMemsetAsync(c->fo_indptr[which].p, 0, sizeof(int64_t) * (size_t)(rows + 1), st));
        } else {
            ARCTE_TRY(community_weighting_device(c, rows, F, g_nnz, c->fg_indptr.as<int64_t>(),
                                                 c->fg_indices.as<int32_t>(), c->fg_data.as<double>(),
                                                 c->scratch[6].as<double>(), c->fo_indptr[which], c->fo_indices[which],
                                                 c->fo_data[which], &kept));
        }
        c->fo_rows[which] = rows;
        c->fo_nnz[which] = kept;
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    c->fo_valid = true;
    if (train_nnz) *train_nnz = c->fo_nnz[0];
    if (test_nnz) *test_nnz = c->fo_nnz[1];
    return ARCTE_OK;
}

int arcte_cuda_get_fold(arcte_cuda_ctx *c, int which, int64_t *host_indptr, int32_t *host_indices, double *host_data)
{
    CHECK_CTX(c);
    if (!c->fo_valid || which < 0 || which > 1 || !host_indptr) { set_error("get_fold: call weighted_fold first"); return ARCTE_E_ARG; }
    const int64_t rows = c->fo_rows[which], nnz = c->fo_nnz[which];
    ARCTE_TRY(copy_to_host(c, host_indptr, c->fo_indptr[which].p, sizeof(int64_t) * (size_t)(rows + 1)));
    if (nnz > 0) {
        if (!host_indices || !host_data) { set_error("get_fold: null output"); return ARCTE_E_ARG; }
        ARCTE_TRY(copy_to_host(c, host_indices, c->fo_indices[which].p, sizeof(int32_t) * (size_t)nnz));
        ARCTE_TRY(copy_to_host(c, host_data, c->fo_data[which].p, sizeof(double) * (size_t)nnz));
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

}  // extern "C"
