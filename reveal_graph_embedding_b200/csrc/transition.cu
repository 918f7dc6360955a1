// transition.cu -- K1 (degrees + row normalisation), K2a (seed selection/ordering) and
// K2b (epsilon-effective per seed).
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   K1   get_natural_random_walk_matrix        eps_randomwalk/transition.py:43-68
//   K2a  seed list in arcte()                  embedding/arcte/arcte.py:610-617
//   K2b  calculate_epsilon_effective           embedding/arcte/arcte.py:26-50
//
// Bit-exactness notes.  The degree vectors are scipy.sparse sums whose rounding depends
// on the order of the additions; both orders are reproduced exactly ("K1"):
//
//   out_degree[i] = data[first] + numpy_pairwise_sum(data[first+1 : end])   (add.reduceat)
//   in_degree[c]  = ((0 + a_{r1,c}) + a_{r2,c}) + ...  rows ascending       (CSC mat-vec)
// The second needs the column's entries in row order, i.e. a transposition; it is done
// with the stable radix sort (key = column, value = weight) so no atomics on doubles and
// no run-to-run variation.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "primitives.cuh"

namespace arcte {

constexpr int64_t kHeavyRow = 128;  // numpy's pairwise leaf size: longer sums are split over a warp

// ---- K1a: out-degree, one thread per row ----------------------------------------------
__global__ void __launch_bounds__(256)
k_row_degree(int64_t n, const int64_t *__restrict__ indptr, const double *__restrict__ adj,
             double *__restrict__ d_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t b = indptr[i], e = indptr[i + 1];
    if (e - b - 1 > kHeavyRow) return;  // long rows: k_row_degree_heavy
    double sum = 0.0;
    if (e > b) {
        sum = adj[b];
        if (e - b > 1) {
            auto at = [adj](int64_t j) { return adj[j]; };
            sum = __dadd_rn(sum, pairwise_sum(at, b + 1, e - b - 1));
        }
    }
    if (sum == 0.0) sum = 1.0;  // transition.py:58
    d_out[i] = sum;
}

// Same for rows whose tail is longer than one pairwise leaf: one warp per row.
__global__ void __launch_bounds__(256)
k_row_degree_heavy(int64_t n, const int64_t *__restrict__ indptr, const double *__restrict__ adj,
                   double *__restrict__ d_out)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int64_t b = indptr[i], e = indptr[i + 1];
    if (e - b - 1 <= kHeavyRow) return;
    auto at = [adj](int64_t j) { return adj[j]; };
    double sum = __dadd_rn(adj[b], pairwise_sum_warp(at, b + 1, e - b - 1));
    if (sum == 0.0) sum = 1.0;
    if (lane_id() == 0) d_out[i] = sum;
}

// ---- K1b: binarised column counts + sort keys ------------------------------------------
__global__ void __launch_bounds__(256)
k_col_count(int64_t nnz, const int32_t *__restrict__ indices, int32_t *__restrict__ colcnt,
            uint32_t *__restrict__ keys)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int32_t c = indices[j];
    keys[j] = (uint32_t)c;
    atomicAdd(&colcnt[c], 1);
}

// ---- K1c: in-degree, one thread per column over the column-sorted weights ---------------
__global__ void __launch_bounds__(256)
k_col_degree(int64_t n, const int64_t *__restrict__ cscptr, const double *__restrict__ sorted_w,
             double *__restrict__ d_in)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double y = 0.0;
    for (int64_t j = cscptr[c]; j < cscptr[c + 1]; ++j) y = __dadd_rn(y, sorted_w[j]);
    d_in[c] = y;
}

// ---- K1d: W = D_out^-1 A, one warp per row ------------------------------------------------
__global__ void __launch_bounds__(256)
k_normalise(int64_t n, const int64_t *__restrict__ indptr, const double *__restrict__ adj,
            const double *__restrict__ d_out, double *__restrict__ w)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t b = indptr[row], e = indptr[row + 1];
    const double d = d_out[row];
    for (int64_t j = b + lane_id(); j < e; j += 32) w[j] = __ddiv_rn(adj[j], d);  // transition.py:61-63
}

// ---- K1e: per-node record {d_in, row begin, row length} for the push kernel ----------------
__global__ void __launch_bounds__(256)
k_node_info(int64_t n, const int64_t *__restrict__ indptr, const double *__restrict__ d_in,
            NodeInfo *__restrict__ info)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NodeInfo r;
    r.d_in = d_in[i];
    r.begin = (uint32_t)indptr[i];
    r.len = (uint32_t)(indptr[i + 1] - indptr[i]);
    info[i] = r;
}

// ---- K1f: per-edge record {w_uv, d_in[v]} for the push kernel: the in-degree the enqueue test
// needs (similarity.py:194 / :214) arrives with the coalesced read of the row instead of through
// a second random gather per neighbour ----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_edge_records(int64_t nnz, const int32_t *__restrict__ indices, const double *__restrict__ w,
               const double *__restrict__ d_in, double2 *__restrict__ wd)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride)
        wd[j] = make_double2(w[j], d_in[indices[j]]);
}

// In-degree of the target of every stored entry, read coalesced next to the column index by the batched
// engine: no random gather of the node record per neighbour touch (similarity.py:194 / :214 need d_in[v]).
__global__ void __launch_bounds__(256)
k_edge_din(int64_t nnz, const int32_t *__restrict__ indices, const double *__restrict__ d_in, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) out[j] = d_in[indices[j]];
}

// ---- K2a: seed keys -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_count_stats(int64_t n, const int32_t *__restrict__ colcnt, int64_t *__restrict__ out /*[2]: max, n_seeds*/)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t c = i < n ? colcnt[i] : 0;
    int is_seed = c > 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c = max(c, __shfl_xor_sync(kFull, c, o));
        is_seed += __shfl_xor_sync(kFull, is_seed, o);
    }
    if (lane_id() == 0) {
        atomicMax((unsigned long long *)&out[0], (unsigned long long)c);
        if (is_seed) atomicAdd((unsigned long long *)&out[1], (unsigned long long)is_seed);
    }
}

__global__ void __launch_bounds__(256)
k_seed_keys(int64_t n, const int32_t *__restrict__ colcnt, int32_t maxcnt, uint32_t *__restrict__ keys,
            uint32_t *__restrict__ ids)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = colcnt[i];
    // seeds (count > 1, arcte.py:617) sort by count descending; everything else goes last
    keys[i] = c > 1 ? (uint32_t)(maxcnt - c) : (uint32_t)maxcnt;
    ids[i] = (uint32_t)i;
}

// ---- K2b: epsilon-effective; one thread per seed, one warp per seed with > 128 neighbours ----
__device__ __forceinline__ double eps_effective_finish(double epsilon, double ds, double mean, double dmin,
                                                       double dmax)
{
    // arcte.py:35
    double eff = __ddiv_rn(__dmul_rn(epsilon, log(__dadd_rn(1.0, ds))), log(__dadd_rn(1.0, mean)));
    const double hi = __ddiv_rn(1.0, __dmul_rn(ds, dmin));  // arcte.py:39
    const double lo = __ddiv_rn(1.0, __dmul_rn(ds, dmax));  // arcte.py:40
    if (eff > hi) eff = hi;                                       // arcte.py:45-46
    else if (eff < lo) eff = __ddiv_rn(__dadd_rn(lo, eff), 2.0);  // arcte.py:47-48
    return eff;
}

__global__ void __launch_bounds__(128)
k_eps_effective(int64_t n_seeds, const int32_t *__restrict__ seeds, double epsilon,
                const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const double *__restrict__ d_out, double *__restrict__ eps_out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seeds) return;
    const int32_t seed = seeds[k];
    const int64_t b = indptr[seed], e = indptr[seed + 1];
    const int64_t deg = e - b;
    if (deg > kHeavyRow) return;  // k_eps_effective_heavy
    auto at = [indices, d_out](int64_t j) { return d_out[indices[j]]; };
    const double mean = __ddiv_rn(pairwise_sum(at, b, deg), (double)deg);  // arcte.py:32
    double dmin = INFINITY, dmax = -INFINITY;
    for (int64_t j = b; j < e; ++j) {
        const double d = d_out[indices[j]];
        dmin = fmin(dmin, d);
        dmax = fmax(dmax, d);
    }
    eps_out[k] = eps_effective_finish(epsilon, d_out[seed], mean, dmin, dmax);
}

__global__ void __launch_bounds__(256)
k_eps_effective_heavy(int64_t n_seeds, const int32_t *__restrict__ seeds, double epsilon,
                      const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      const double *__restrict__ d_out, double *__restrict__ eps_out)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_seeds) return;
    const int32_t seed = seeds[k];
    const int64_t b = indptr[seed], e = indptr[seed + 1];
    const int64_t deg = e - b;
    if (deg <= kHeavyRow) return;
    auto at = [indices, d_out](int64_t j) { return d_out[indices[j]]; };
    const double mean = __ddiv_rn(pairwise_sum_warp(at, b, deg), (double)deg);  // arcte.py:32
    double dmin = INFINITY, dmax = -INFINITY;
    for (int64_t j = b + lane_id(); j < e; j += 32) {
        const double d = d_out[indices[j]];
        dmin = fmin(dmin, d);
        dmax = fmax(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = fmin(dmin, __shfl_xor_sync(kFull, dmin, o));
        dmax = fmax(dmax, __shfl_xor_sync(kFull, dmax, o));
    }
    if (lane_id() == 0) eps_out[k] = eps_effective_finish(epsilon, d_out[seed], mean, dmin, dmax);
}

static int bit_length(uint64_t v)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

static inline unsigned grid_for(int64_t items, int block) { return (unsigned)((items + block - 1) / block); }

// Row-uniform transition weights (every unweighted graph: w[u][:] = 1/d_out[u]): the batched engine then
// reads one weight per pushed row instead of one per stored entry.  One warp per row.
__global__ void k_row_uniform(int64_t n, const NodeInfo *__restrict__ info, const double *__restrict__ w,
                              double *__restrict__ row_w, unsigned long long *__restrict__ any_nonuniform)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const NodeInfo iu = info[row];
    const double w0 = iu.len ? w[iu.begin] : 0.0;
    bool same = true;
    for (unsigned j = lane; j < iu.len; j += 32)
        same = same && (__double_as_longlong(w[iu.begin + j]) == __double_as_longlong(w0));
    same = __all_sync(kFull, same);
    if (lane == 0) {
        row_w[row] = w0;
        if (!same) atomicOr(any_nonuniform, 1ull);
        // unit adjacency weights: the row sum is the row length and every transition weight is 1/len
        // (transition.py:61-63), which a push kernel can recompute instead of reading the weight array
        if (iu.len && (!same || __double_as_longlong(w0) != __double_as_longlong(__ddiv_rn(1.0, (double)iu.len))))
            atomicOr(any_nonuniform + 1, 1ull);
    }
}

// ---- K2c: walk labels ------------------------------------------------------------------------
// The push kernels keep one dense state array per walk in flight and touch it at random; what bounds them is
// the number of distinct DRAM lines a walk touches (profiles/r2_engines.md).  A walk's support is, to first
// order, a union of complete neighbourhoods of the hubs it pushes, so the walks run in a private label space in
// which the nodes that share their two highest-count neighbours are consecutive: node v sorts by (rank of its
// best neighbour, rank of its second-best neighbour, v), rank = position in the count-descending node list of
// K2a.  Only the column indices the push kernels read are renamed -- every row keeps its position and its stored
// order, so the queue discipline, every sum and the order of the emitted members are what they were
// (similarity.py:194-216 enqueue in CSR order) -- and members are renamed back when they are emitted.
__global__ void __launch_bounds__(256)
k_rank_from_order(int64_t n, const int32_t *__restrict__ order, uint32_t *__restrict__ rank)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[order[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
k_hub_keys(int64_t n, const NodeInfo *__restrict__ info, const int32_t *__restrict__ indices,
           const uint32_t *__restrict__ rank, uint32_t *__restrict__ key1, uint32_t *__restrict__ key2,
           uint32_t *__restrict__ ids)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const NodeInfo iu = info[row];
    uint32_t m1 = (uint32_t)n, m2 = (uint32_t)n;   // smallest and second-smallest neighbour rank
    for (unsigned j = lane_id(); j < iu.len; j += 32) {
        const uint32_t r = rank[indices[iu.begin + j]];
        if (r < m1) { m2 = m1; m1 = r; }
        else if (r < m2 && r != m1) m2 = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t a1 = __shfl_xor_sync(kFull, m1, o), a2 = __shfl_xor_sync(kFull, m2, o);
        const uint32_t lo = min(m1, a1), hi = max(m1, a1);
        m2 = min(hi == lo ? (uint32_t)n : hi, min(m2, a2));
        m1 = lo;
    }
    if (lane_id() == 0) {
        key1[row] = m1;
        key2[row] = m2;
        ids[row] = (uint32_t)row;
    }
}

__global__ void __launch_bounds__(256)
k_gather_u32(int64_t n, const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// order[i] = node that gets walk label i
__global__ void __launch_bounds__(256)
k_walk_nodes(int64_t n, const uint32_t *__restrict__ order, const NodeInfo *__restrict__ info,
             const double *__restrict__ row_w, int32_t *__restrict__ to_walk, int32_t *__restrict__ from_walk,
             NodeInfo *__restrict__ walk_info, double *__restrict__ walk_row_w)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = order[i];
    to_walk[v] = (int32_t)i;
    from_walk[i] = (int32_t)v;
    walk_info[i] = info[v];
    walk_row_w[i] = row_w[v];
}

__global__ void __launch_bounds__(256)
k_walk_indices(int64_t nnz, const int32_t *__restrict__ indices, const int32_t *__restrict__ to_walk,
               int32_t *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) out[j] = to_walk[indices[j]];
}

// Needs node_info, row_w and the count-descending node list (c->seeds, all n entries) of select_seeds.
static int build_walk_labels(arcte_cuda_ctx *c)
{
    c->walk_labels_valid = false;
    const char *env = getenv("ARCTE_CUDA_WALK_LABELS");
    if (env && !strcmp(env, "0")) return ARCTE_OK;
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    int64_t *launches = &c->stats.launches;
    const size_t m = (size_t)n + 1;
    ARCTE_TRY(dev_reserve(c->walk_info, sizeof(NodeInfo) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->walk_row_w, sizeof(double) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->walk_indices, sizeof(int32_t) * (size_t)(c->nnz > 0 ? c->nnz : 1)));
    ARCTE_TRY(dev_reserve(c->to_walk, sizeof(int32_t) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->from_walk, sizeof(int32_t) * (size_t)n));
    for (int k : {0, 1, 2, 3, 8, 9}) ARCTE_TRY(dev_reserve(c->scratch[k], sizeof(uint32_t) * m));
    uint32_t *rank = c->scratch[8].as<uint32_t>(), *key1 = c->scratch[9].as<uint32_t>();
    k_rank_from_order<<<grid_for(n, 256), 256, 0, st>>>(n, c->seeds.as<int32_t>(), rank);
    k_hub_keys<<<grid_for(n * 32, 256), 256, 0, st>>>(n, c->node_info.as<NodeInfo>(), c->indices.as<int32_t>(), rank, key1,
                                                     c->scratch[0].as<uint32_t>(), c->scratch[2].as<uint32_t>());
    *launches += 2;
    const int bits = bit_length((uint64_t)n);
    bool second = false;
    ARCTE_TRY(radix_sort_pairs(c->scratch[0].as<uint32_t>(), c->scratch[2].p, c->scratch[1].as<uint32_t>(), c->scratch[3].p,
                               n, bits, 4, c->scratch[4], c->scratch[5], c->scratch[6], st, &second, launches));
    // nodes by second-best neighbour; stable sort by the best neighbour on top gives the lexicographic order
    uint32_t *ids_a = second ? c->scratch[3].as<uint32_t>() : c->scratch[2].as<uint32_t>();
    uint32_t *ids_b = second ? c->scratch[2].as<uint32_t>() : c->scratch[3].as<uint32_t>();
    uint32_t *keys_a = second ? c->scratch[1].as<uint32_t>() : c->scratch[0].as<uint32_t>();
    uint32_t *keys_b = second ? c->scratch[0].as<uint32_t>() : c->scratch[1].as<uint32_t>();
    k_gather_u32<<<grid_for(n, 256), 256, 0, st>>>(n, key1, ids_a, keys_a);
    ++*launches;
    ARCTE_TRY(radix_sort_pairs(keys_a, ids_a, keys_b, ids_b, n, bits, 4, c->scratch[4], c->scratch[5], c->scratch[6], st,
                               &second, launches));
    const uint32_t *order = second ? ids_b : ids_a;
    k_walk_nodes<<<grid_for(n, 256), 256, 0, st>>>(n, order, c->node_info.as<NodeInfo>(), c->row_w.as<double>(),
                                                  c->to_walk.as<int32_t>(), c->from_walk.as<int32_t>(),
                                                  c->walk_info.as<NodeInfo>(), c->walk_row_w.as<double>());
    ++*launches;
    if (c->nnz > 0) {
        unsigned g = grid_for(c->nnz, 256);
        if (g > (unsigned)c->sm_count * 16) g = (unsigned)c->sm_count * 16;
        k_walk_indices<<<g, 256, 0, st>>>(c->nnz, c->indices.as<int32_t>(), c->to_walk.as<int32_t>(),
                                          c->walk_indices.as<int32_t>());
        ++*launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    c->walk_labels_valid = true;
    return ARCTE_OK;
}

// ---- K2a: seeds, count-descending (ties: ascending node id); needs colcnt ----
int select_seeds(arcte_cuda_ctx *c)
{
    const int64_t n = c->n;
    cudaStream_t st = c->stream;
    int64_t *launches = &c->stats.launches;
    // the node records depend on d_in, which is final by now on both upload paths
    ARCTE_TRY(dev_reserve(c->node_info, sizeof(NodeInfo) * (size_t)n));
    k_node_info<<<grid_for(n, 256), 256, 0, st>>>(n, c->indptr.as<int64_t>(), c->d_in.as<double>(),
                                                  c->node_info.as<NodeInfo>());
    ++*launches;
#if ARCTE_EDGE_RECORDS
    ARCTE_TRY(dev_reserve(c->edge_wd, sizeof(double2) * (size_t)(c->nnz > 0 ? c->nnz : 1)));
    if (c->nnz > 0) {
        unsigned g = grid_for(c->nnz, 256);
        if (g > (unsigned)c->sm_count * 16) g = (unsigned)c->sm_count * 16;
        k_edge_records<<<g, 256, 0, st>>>(c->nnz, c->indices.as<int32_t>(), c->w.as<double>(), c->d_in.as<double>(),
                                          c->edge_wd.as<double2>());
        ++*launches;
    }
#endif
    ARCTE_TRY(dev_reserve(c->edge_din, sizeof(double) * (size_t)(c->nnz > 0 ? c->nnz : 1)));
    if (c->nnz > 0) {
        unsigned g = grid_for(c->nnz, 256);
        if (g > (unsigned)c->sm_count * 16) g = (unsigned)c->sm_count * 16;
        k_edge_din<<<g, 256, 0, st>>>(c->nnz, c->indices.as<int32_t>(), c->d_in.as<double>(), c->edge_din.as<double>());
        ++*launches;
    }
    const size_t m = (size_t)n + 1;
    ARCTE_TRY(dev_reserve(c->seeds, sizeof(int32_t) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[1], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[3], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[7], sizeof(int64_t) * 8));
    ARCTE_TRY(dev_reserve(c->counters, sizeof(int64_t) * PC_COUNT));
    ARCTE_TRY(dev_reserve(c->row_w, sizeof(double) * (size_t)n));
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev0, st));
    int64_t *tmp2 = c->scratch[7].as<int64_t>();
    ARCTE_CUDA_TRY(cudaMemsetAsync(tmp2, 0, 5 * sizeof(int64_t), st));
    k_count_stats<<<grid_for(n, 256), 256, 0, st>>>(n, c->colcnt.as<int32_t>(), tmp2);
    ++*launches;
    k_row_uniform<<<grid_for(n * 32, 256), 256, 0, st>>>(n, c->node_info.as<NodeInfo>(), c->w.as<double>(),
                                                        c->row_w.as<double>(), (unsigned long long *)(tmp2 + 2));
    ++*launches;
    int64_t host2[5];
    ARCTE_CUDA_TRY(cudaMemcpyAsync(host2, tmp2, sizeof(host2), cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    const int32_t maxcnt = (int32_t)host2[0];
    c->n_seeds = host2[1];
    c->uniform_rows = host2[2] == 0 && !getenv("ARCTE_CUDA_NO_UNIFORM");
    c->unit_rows = host2[3] == 0 && !getenv("ARCTE_CUDA_NO_UNIT_ROWS");
    c->row_w_valid = true;
    k_seed_keys<<<grid_for(n, 256), 256, 0, st>>>(n, c->colcnt.as<int32_t>(), maxcnt,
                                                  c->scratch[0].as<uint32_t>(),
                                                  (uint32_t *)c->scratch[2].p);
    ++*launches;
    bool second = false;
    ARCTE_TRY(radix_sort_pairs(c->scratch[0].as<uint32_t>(), c->scratch[2].p,
                               c->scratch[1].as<uint32_t>(), c->scratch[3].p, n,
                               bit_length((uint64_t)maxcnt), 4, c->scratch[4], c->scratch[5],
                               c->scratch[6], st, &second, launches));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->seeds.p, second ? c->scratch[3].p : c->scratch[2].p,
                                   sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    ARCTE_TRY(build_walk_labels(c));
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev1, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.ms_seeds = ms;
    c->stats.n_seeds_total = c->n_seeds;
    c->have_segments = false;
    c->have_features = false;
    return ARCTE_OK;
}

// Column counts only (used when the caller supplies W and the degrees itself).
int count_columns(arcte_cuda_ctx *c)
{
    ARCTE_TRY(dev_reserve(c->colcnt, sizeof(int32_t) * (size_t)c->n));
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(uint32_t) * (size_t)((c->nnz > c->n ? c->nnz : c->n) + 1)));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->colcnt.p, 0, sizeof(int32_t) * (size_t)c->n, c->stream));
    if (c->nnz > 0) {
        k_col_count<<<grid_for(c->nnz, 256), 256, 0, c->stream>>>(c->nnz, c->indices.as<int32_t>(),
                                                                  c->colcnt.as<int32_t>(),
                                                                  c->scratch[0].as<uint32_t>());
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

int build_transition(arcte_cuda_ctx *c)
{
    if (!c->have_graph) { set_error("build_transition: no graph resident"); return ARCTE_E_ARG; }
    const int64_t n = c->n, nnz = c->nnz;
    cudaStream_t st = c->stream;
    int64_t *launches = &c->stats.launches;
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev0, st));

    ARCTE_TRY(dev_reserve(c->w, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    ARCTE_TRY(dev_reserve(c->d_out, sizeof(double) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->d_in, sizeof(double) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->colcnt, sizeof(int32_t) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->seeds, sizeof(int32_t) * (size_t)n));

    // scratch: [0] keys a, [1] keys b, [2] vals a, [3] vals b, [4] hist, [5] scan, [6] scan2, [7] cscptr
    const size_t m = (size_t)(nnz > n ? nnz : n) + 1;
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[1], sizeof(uint32_t) * m));
    ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(double) * m));
    ARCTE_TRY(dev_reserve(c->scratch[3], sizeof(double) * m));
    ARCTE_TRY(dev_reserve(c->scratch[7], sizeof(int64_t) * (size_t)(n + 2)));

    k_row_degree<<<grid_for(n, 256), 256, 0, st>>>(n, c->indptr.as<int64_t>(), c->adj.as<double>(),
                                                   c->d_out.as<double>());
    k_row_degree_heavy<<<grid_for(n * 32, 256), 256, 0, st>>>(n, c->indptr.as<int64_t>(), c->adj.as<double>(),
                                                            c->d_out.as<double>());
    *launches += 2;
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->colcnt.p, 0, sizeof(int32_t) * (size_t)n, st));
    if (nnz > 0) {
        k_col_count<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, c->indices.as<int32_t>(),
                                                        c->colcnt.as<int32_t>(),
                                                        c->scratch[0].as<uint32_t>());
        ++*launches;
    }
    ARCTE_TRY(exclusive_scan_i32(c->colcnt.as<int32_t>(), c->scratch[7].as<int64_t>(), n,
                                 c->scratch[6], st, launches));
    // column-major order of the weights: stable sort by column keeps rows ascending
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[2].p, c->adj.p, sizeof(double) * (size_t)nnz,
                                   cudaMemcpyDeviceToDevice, st));
    bool second = false;
    ARCTE_TRY(radix_sort_pairs(c->scratch[0].as<uint32_t>(), c->scratch[2].p,
                               c->scratch[1].as<uint32_t>(), c->scratch[3].p, nnz,
                               bit_length((uint64_t)(n > 0 ? n - 1 : 0)), 8, c->scratch[4],
                               c->scratch[5], c->scratch[6], st, &second, launches));
    const double *sorted_w = second ? c->scratch[3].as<double>() : c->scratch[2].as<double>();
    k_col_degree<<<grid_for(n, 256), 256, 0, st>>>(n, c->scratch[7].as<int64_t>(), sorted_w,
                                                   c->d_in.as<double>());
    ++*launches;
    k_normalise<<<grid_for(n * 32, 256), 256, 0, st>>>(n, c->indptr.as<int64_t>(), c->adj.as<double>(),
                                                       c->d_out.as<double>(), c->w.as<double>());
    ++*launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev1, st));

    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.ms_transition = ms;
    c->have_transition = true;
    return select_seeds(c);
}

// eps_out[k] for seeds[k], both on the device.
int compute_eps_effective(arcte_cuda_ctx *c, double epsilon, const int32_t *dev_seeds,
                          int64_t n_seeds, double *dev_eps_out)
{
    if (n_seeds == 0) return ARCTE_OK;
    k_eps_effective<<<grid_for(n_seeds, 128), 128, 0, c->stream>>>(
        n_seeds, dev_seeds, epsilon, c->indptr.as<int64_t>(), c->indices.as<int32_t>(),
        c->d_out.as<double>(), dev_eps_out);
    k_eps_effective_heavy<<<grid_for(n_seeds * 32, 256), 256, 0, c->stream>>>(
        n_seeds, dev_seeds, epsilon, c->indptr.as<int64_t>(), c->indices.as<int32_t>(),
        c->d_out.as<double>(), dev_eps_out);
    c->stats.launches += 2;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

}  // namespace arcte
