// common.cuh -- context layout, error handling and small device helpers shared by
// the translation units of libarcte_cuda.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>

#include "../../include/arcte_cuda.h"

// Experiment switch (default off, measured neutral on both bench shapes: profiles/README.md):
// 1 = the push kernel reads {weight, in-degree of the target} as one coalesced 16-byte record per
// stored entry instead of the weight array plus a random gather of the target's node record.
#ifndef ARCTE_EDGE_RECORDS
#define ARCTE_EDGE_RECORDS 0
#endif

namespace arcte {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const std::string &msg);

#define ARCTE_CUDA_TRY(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::arcte::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                               __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
            return ARCTE_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ARCTE_TRY(expr)                  \
    do {                                 \
        int _rc = (expr);                \
        if (_rc != ARCTE_OK) return _rc; \
    } while (0)

// A device allocation that remembers its size so buffers can be grown lazily.
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    bool borrowed = false;  // a view into another allocation (the graph arena): never freed or regrown on its own
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};
int dev_reserve(DevBuf &b, size_t bytes);  // grow-only, contents not preserved
void dev_free(DevBuf &b);

// Immutable per-node record read on every pop and every neighbour touch: one 16-byte load
// gives the in-degree and the CSR row of the node.
struct __align__(16) NodeInfo {
    double d_in;
    uint32_t begin;  // indptr[v]
    uint32_t len;    // indptr[v+1] - indptr[v]
};

// Per-seed walk state of one warp ("slot"): interleaved {s, r} pairs over all n
// nodes (always all-zero between seeds), the touched list and the FIFO ring.
struct SlotPool {
    DevBuf sr;       // double2 [n_slots][n]
    DevBuf touched;  // int32   [n_slots][n]
    DevBuf queue;    // int32   [n_slots][queue_cap]  (compact engine: int2 {node, compact index})
    DevBuf cmap;     // uint32  [n_slots][map_stride]  compact engine: epoch-tagged index map
    DevBuf cepoch;   // uint32  [n_slots]              compact engine: epoch of the last walk of each slot
    bool compact = false;     // sr holds compact pairs (not the all-zero dense arrays the dense FIFO engine relies on)
    int64_t map_stride = 0;
    int64_t ccap = 0;         // compact engine: pairs / touched entries per slot (dense engines: n)
    DevBuf frontier; // int32   [frontier_slots][2][n]  (frontier schedule only)
    DevBuf fval;     // double  [frontier_slots][n]     (frontier schedule only)
    int64_t frontier_slots = 0;
    int64_t n_slots = 0;      // state/touched rows allocated
    int64_t queue_slots = 0;  // FIFO rings allocated (== n_slots except after a retry grew the rings)
    int64_t queue_cap = 0;
    int64_t n = 0;
};

// Device-side counters of the fused push kernel (one int64 each, atomically added).
enum PushCounter {
    PC_PUSHES = 0, PC_EDGES, PC_ENQUEUES, PC_MAXQ, PC_SUPPORT, PC_TOUCHED, PC_SEEDDEG, PC_MEMBERS,
    PC_EMITTED, PC_OVERFLOW_SEEDS, PC_QOVERFLOW, PC_MEMBER_CURSOR, PC_WORK_CURSOR,
    PC_T_START, PC_T_END, PC_T_BUSY, PC_WORK_CURSOR2, PC_ROUNDS, PC_TOVERFLOW,
    PC_PROF0, PC_PROF1, PC_PROF2, PC_PROF3, PC_PROF4, PC_PROF5, PC_PROF6, PC_PROF7, PC_PROF8, PC_PROF9,  // kernel experiments
    PC_PROF10, PC_PROF11, PC_PROF12, PC_PROF13, PC_PROF14, PC_PROF15, PC_PROF16, PC_PROF17, PC_PROF18, PC_PROF19,
    PC_COUNT
};

// Slot pool of the batched hash engine (push_batched.cu): per slot two table halves of `cap` 32-byte
// entries, a member staging list of `cap` ints and a FIFO ring.
struct BatchedPool {
    DevBuf tbl;     // hash: TableEntry [n_slots][2][cap]; direct: TableEntry [n_slots][cap], cap = n
    DevBuf stage;   // int32 [n_slots][cap]  hash: member staging; direct: touched list, then members
    DevBuf clean;   // int32 [n_slots][2]  hash: entries of each half known all-EMPTY; direct: [0] = last epoch used
    DevBuf queue;   // int32 [n_slots][queue_cap]
    int64_t n_slots = 0, cap = 0, plan_cap = 0, cap_cfg = 0, queue_cap = 0, queue_slots = 0;
    int mode = -1;  // engine the pool is laid out for
};

// Pinned staging ring of the streamed host copies (hostcopy.cu): two 4 MB slots, a stream and two events per
// worker thread.
constexpr int kMaxCopyThreads = 16;
struct HostRing {
    void *pinned[kMaxCopyThreads] = {};   // per copy thread: two slots of kSlotBytes, allocated once, never moved
    int n_threads = 0;
    cudaStream_t streams[kMaxCopyThreads] = {};
    cudaEvent_t events[2 * kMaxCopyThreads] = {};
    std::mutex mutex;                     // growth
};
int copy_to_host(arcte_cuda_ctx *c, void *dst, const void *src, size_t bytes);
int copy_from_host(arcte_cuda_ctx *c, void *dst, const void *src, size_t bytes);
int fetch_features(arcte_cuda_ctx *c, int64_t *host_indptr, int32_t *host_indices, double *host_data, int values_are_ones,
                   int n_threads, long pid = 0);
int host_write_remote(long pid, void *remote_dst, const void *local_src, size_t bytes);
void host_advise_huge(void *p, size_t bytes);
int host_ones_alloc(size_t count, void **out, size_t *mapped_bytes);
void host_ones_free(void *p, size_t mapped_bytes);
void free_ring(arcte_cuda_ctx *c);
void comm_free(arcte_cuda_ctx *c);

}  // namespace arcte

struct arcte_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // caller-visible step timer
    cudaEvent_t pk0 = nullptr, pk1 = nullptr;  // around the push-kernel launches of one extraction
    cudaEvent_t xev0 = nullptr, xev1 = nullptr;  // around the multi-GPU exchange

    // multi-GPU exchange (exchange.cu): NCCL communicator (ncclComm_t), the shard the last extract walked
    void *comm = nullptr;
    int comm_world = 0, comm_rank = 0;
    int shard_rank = 0, shard_count = 1;
    arcte::DevBuf xbuf[5];
    arcte::DevBuf l2_flush;
    arcte::HostRing ring;

    // tuning
    int warps_per_sm = 0;      // 0 = default
    int64_t queue_cap_cfg = 0; // 0 = default
    int mem_percent = 0;       // 0 = default
    int64_t member_cap_cfg = 0; // 0 = default
    int schedule = 0;          // ARCTE_SCHEDULE_FIFO (exact) or ARCTE_SCHEDULE_FRONTIER
    int engine = -1;           // ARCTE_ENGINE_* of the FIFO schedule
    int64_t tbl_cap_cfg = 0;   // 0 = default
    int fr_heavy_permille = -1, fr_heavy_threads = 0, fr_heavy_ctas = 0, fr_light_threads = 0, fr_light_ctas = 0;

    // graph (adjacency + transition), all resident
    int64_t n = 0, nnz = 0;
    arcte::DevBuf indptr;   // int64 [n+1]
    arcte::DevBuf indices;  // int32 [nnz]
    arcte::DevBuf adj;      // double [nnz]  adjacency weights
    arcte::DevBuf w;        // double [nnz]  transition probabilities
    arcte::DevBuf d_out;    // double [n]
    arcte::DevBuf d_in;     // double [n]
    arcte::DevBuf colcnt;   // int32 [n]  binarised column counts
    arcte::DevBuf node_info; // NodeInfo [n]
    arcte::DevBuf edge_wd;   // double2 [nnz]  {transition weight, in-degree of the target} per stored entry
    arcte::DevBuf edge_din;  // double [nnz]  in-degree of the target of every stored entry (batched engine)
    bool have_graph = false, have_transition = false;
    // node_info | row_w | indices | w live back to back in one allocation so that ONE L2 access-policy
    // window (persisting) can cover the arrays every walk re-reads
    arcte::DevBuf graph_arena;
    size_t l2_persist_bytes = 0, l2_window_max = 0;

    // seeds
    arcte::DevBuf seeds;    // int32 [n]  count-descending, first n_seeds valid
    int64_t n_seeds = 0;

    // extraction results (segments of this context's shard)
    arcte::DevBuf work_seed;   // int32 [S]  seed node per segment
    arcte::DevBuf work_eps;    // double [S]
    arcte::DevBuf seg_count;   // int32 [S]
    arcte::DevBuf seg_offset;  // int64 [S]
    arcte::DevBuf members;     // int32 [member_cap]
    arcte::DevBuf retry_list;  // int32 [S]  positions whose FIFO overflowed
    int64_t n_segments = 0, n_members = 0, member_cap = 0;
    bool have_segments = false;

    unsigned long long *centrality_acc = nullptr;  // set by arcte_cuda_centrality for the duration of its extraction
    arcte::SlotPool slots;
    arcte::BatchedPool bpool;
    arcte::DevBuf row_w;       // double [n]  the repeated transition weight of each row (uniform_rows)
    bool row_w_valid = false, uniform_rows = false;
    // walk labels (transition.cu, K2c): the label space the FIFO-schedule push kernels run in
    arcte::DevBuf to_walk, from_walk;     // int32 [n]  node -> walk label and back
    arcte::DevBuf walk_info;              // NodeInfo [n]  node records by walk label
    arcte::DevBuf walk_row_w;             // double [n]
    arcte::DevBuf walk_indices;           // int32 [nnz]  column indices in walk labels, rows where and as they were
    arcte::DevBuf work_seed_w;            // int32 [S]  the shard's seeds in walk labels
    arcte::DevBuf work_order;             // int32 [S]  work-list positions by ascending epsilon-effective (longest walks first)
    bool walk_labels_valid = false;
    bool unit_rows = false;               // every transition weight of row u is exactly 1/len(u)
    arcte::DevBuf counters;  // int64 [PC_COUNT]

    // assembly output
    arcte::DevBuf out_indptr;   // int64 [n+1]
    arcte::DevBuf out_indices;  // int32 [nnz_out]
    arcte::DevBuf out_data;     // double [nnz_out]
    int64_t out_nnz = 0;
    int64_t out_rows = 0;       // rows of the assembled block (n unless row-sharded)
    bool have_features = false;

    // feature matrix kept resident for the experiment loop (weighting.cu): the stored matrix, the
    // gathered row block of the fold being processed and the two weighted outputs
    arcte::DevBuf fs_indptr, fs_indices, fs_data;
    int64_t fs_rows = 0, fs_cols = 0, fs_nnz = 0;
    bool fs_valid = false;
    arcte::DevBuf fg_indptr, fg_indices, fg_data, fg_rows;
    arcte::DevBuf fo_indptr[2], fo_indices[2], fo_data[2];
    int64_t fo_rows[2] = {0, 0}, fo_nnz[2] = {0, 0};
    bool fo_valid = false;

    // staging for segment parts that live on other GPUs (kept across calls: cudaMalloc/cudaFree
    // serialise the whole process)
    arcte::DevBuf peer_stage[64];

    // scratch shared by primitives
    arcte::DevBuf scratch[16];

    arcte_cuda_stats stats{};
};

namespace arcte {

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// numpy's DOUBLE_pairwise_sum over n values produced by `at(i)`: identical tree and
// therefore identical rounding to np.add.reduce on a contiguous float64 array
// (blocks of <=128 elements with 8 strided accumulators, halves split on multiples of 8).
template <typename F> __device__ __forceinline__ double pairwise_leaf(F at, int64_t off, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, at(off + i));
        return res;
    }
    double r0 = at(off + 0), r1 = at(off + 1), r2 = at(off + 2), r3 = at(off + 3);
    double r4 = at(off + 4), r5 = at(off + 5), r6 = at(off + 6), r7 = at(off + 7);
    int64_t i;
    const int64_t lim = n - (n % 8);
    for (i = 8; i < lim; i += 8) {
        r0 = __dadd_rn(r0, at(off + i + 0));
        r1 = __dadd_rn(r1, at(off + i + 1));
        r2 = __dadd_rn(r2, at(off + i + 2));
        r3 = __dadd_rn(r3, at(off + i + 3));
        r4 = __dadd_rn(r4, at(off + i + 4));
        r5 = __dadd_rn(r5, at(off + i + 5));
        r6 = __dadd_rn(r6, at(off + i + 6));
        r7 = __dadd_rn(r7, at(off + i + 7));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, at(off + i));
    return res;
}

// The recursion of the numpy routine (n > 128: split at n/2 rounded down to a multiple of
// 8, left + right) unrolled into an explicit post-order walk: no device-side recursion, so
// the stack need is static.  Depth <= 32 covers any n < 2^38.
template <typename F> __device__ double pairwise_sum(F at, int64_t off, int64_t n)
{
    if (n <= 128) return pairwise_leaf(at, off, n);
    constexpr int kDepth = 32;
    int64_t f_off[kDepth], f_n[kDepth];
    double f_left[kDepth];
    signed char f_phase[kDepth];
    int sp = 0;
    f_off[0] = off; f_n[0] = n; f_phase[0] = 0;
    double ret = 0.0;
    while (sp >= 0) {
        const int64_t o = f_off[sp], m = f_n[sp];
        if (m <= 128) {
            ret = pairwise_leaf(at, o, m);
            --sp;
            continue;
        }
        int64_t n2 = m / 2;
        n2 -= n2 % 8;
        if (f_phase[sp] == 0) {          // descend left
            f_phase[sp] = 1;
            ++sp;
            f_off[sp] = o; f_n[sp] = n2; f_phase[sp] = 0;
        } else if (f_phase[sp] == 1) {   // left done, descend right
            f_left[sp] = ret;
            f_phase[sp] = 2;
            ++sp;
            f_off[sp] = o + n2; f_n[sp] = m - n2; f_phase[sp] = 0;
        } else {                         // both done
            ret = __dadd_rn(f_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

// The same sum computed by a whole warp for long inputs, bit-identical to pairwise_sum():
// the top (up to) five levels of numpy's split tree are unrolled over the lanes -- lane bits,
// most significant first, choose left/right -- each lane sums its own subtree with the
// sequential routine, and the partial sums are folded back up in tree order with shuffles.
// All 32 lanes must call it with the same (off, n); every lane returns the total.
template <typename F> __device__ double pairwise_sum_warp(F at, int64_t off, int64_t n)
{
    const int lane = threadIdx.x & 31;
    int64_t o = off, m = n;
    int depth = 0;
#pragma unroll
    for (int lvl = 0; lvl < 5; ++lvl) {
        if (m <= 128) break;
        int64_t n2 = m / 2;
        n2 -= n2 % 8;
        if ((lane >> (4 - lvl)) & 1) {
            o += n2;
            m -= n2;
        } else {
            m = n2;
        }
        depth = lvl + 1;
    }
    // lanes sharing a node: identical top `depth` bits; the one with zero low bits computes it
    const bool owner = (lane & ((1 << (5 - depth)) - 1)) == 0;
    double val = owner ? pairwise_sum(at, o, m) : 0.0;
#pragma unroll
    for (int lvl = 4; lvl >= 0; --lvl) {
        const double other = __shfl_xor_sync(kFull, val, 1 << (4 - lvl));
        const bool left_holder = depth > lvl && (lane & ((1 << (5 - lvl)) - 1)) == 0;
        if (left_holder) val = __dadd_rn(val, other);  // pairwise_sum(left) + pairwise_sum(right)
    }
    return __shfl_sync(kFull, val, 0);
}

// The same sum computed by a whole CTA of 1024 threads (10 tree levels over the thread index,
// most significant bit first), bit-identical to pairwise_sum().  `red` is 1024 doubles of shared
// memory.  Every thread must call it with the same (off, n); every thread returns the total.
template <typename F> __device__ double pairwise_sum_block1024(F at, int64_t off, int64_t n, double *red)
{
    const int t = threadIdx.x;
    int64_t o = off, m = n;
    int depth = 0;
#pragma unroll
    for (int lvl = 0; lvl < 10; ++lvl) {
        if (m <= 128) break;
        int64_t n2 = m / 2;
        n2 -= n2 % 8;
        if ((t >> (9 - lvl)) & 1) {
            o += n2;
            m -= n2;
        } else {
            m = n2;
        }
        depth = lvl + 1;
    }
    const bool owner = (t & ((1 << (10 - depth)) - 1)) == 0;
    double val = owner ? pairwise_sum(at, o, m) : 0.0;
    for (int lvl = 9; lvl >= 0; --lvl) {
        red[t] = val;
        __syncthreads();
        const bool left_holder = depth > lvl && (t & ((1 << (10 - lvl)) - 1)) == 0;
        if (left_holder) val = __dadd_rn(val, red[t ^ (1 << (9 - lvl))]);  // left + right
        __syncthreads();
    }
    if (t == 0) red[0] = val;
    __syncthreads();
    const double total = red[0];
    __syncthreads();
    return total;
}

}  // namespace arcte
