// push.cuh -- parameters shared by the two walk schedules of the push engine:
//   push.cu           exact FIFO replay of the reference's queue discipline (default)
//   push_batched.cu   the same exact FIFO order, several queue entries per warp iteration, walk state
//                     in a compact per-walk hash table (or dense), shared-memory staging (default)
//   push_frontier.cu  synchronous frontier rounds on fixed-point state (opt-in, tolerance parity)
#pragma once

#include "common.cuh"

// Compact engine (push_compact.cu): the kernel is compiled for two occupancy points, 6 resident CTAs of 8 warps per
// SM (40 registers, 48 walks per SM) and 8 (32 registers, 64 walks per SM).  Measured on one B200
// (profiles/r2_occupancy_ab.md): graphs of short rows are latency-bound and gain from more walks in flight (YouTube
// shape, 5.3 entries per row: 591 -> 572 ms at 64), graphs of long rows lose (Flickr shape, 146 per row: 109 -> 114 ms).
// The slot pool follows the choice (plan_slots); the launch picks the instantiation that matches the pool.
#ifndef ARCTE_COMPACT_MIN_BLOCKS
#define ARCTE_COMPACT_MIN_BLOCKS 6
#endif
#ifndef ARCTE_COMPACT_MIN_BLOCKS_HI
#define ARCTE_COMPACT_MIN_BLOCKS_HI 8
#endif

namespace arcte {

// Walk-state entry of the batched engines: ONE 32-byte sector per touched node, read and written with
// single 256-bit accesses.  d_in rides along so that neither the enqueue test of a re-touch nor the
// threshold sweep has to gather the node record again.  `key` is the node id in a hash table
// (kEmptyKey when free) and the walk's epoch in the direct-mapped layout (entry v belongs to node v and
// is valid only when its epoch is the current walk's: nothing is ever reset).
struct __align__(32) TableEntry {
    double s, r, d_in;
    int32_t key;   // node id, kEmptyKey when free
    int32_t aux;
};
constexpr int32_t kEmptyKey = -1;

struct PushParams {
    int64_t n;
    const NodeInfo *info;      // {d_in, row begin, row length} per node
    const int32_t *indices;
    const double *w;
    const double2 *wd;         // {w, d_in of the target} per stored entry (FIFO schedule)
    // work list
    const int32_t *work_seed;  // [n_work_total] seed node per position
    const double *work_eps;    // [n_work_total]
    const int32_t *work_ids;   // positions to run (retry pass) or nullptr = 0..n_work-1
    int64_t n_work;
    int retry_pass;
    // slots
    double2 *sr;
    int32_t *touched;
    int32_t *queue;
    int64_t queue_cap;  // power of two
    int64_t n_slots;
    // outputs
    int32_t *seg_count;
    int64_t *seg_offset;
    int32_t *members;
    int64_t member_cap;
    int32_t *retry_list;
    unsigned long long *counters;
    // rule constants, computed on the host exactly as Python evaluates them
    double rho;            // rho
    double one_minus_rho;  // (1-rho)
    double lazy_b;         // (1-rho)*(1-lazy)
    double lazy_c;         // (1-rho)*lazy
    int debug_keep;        // operator seam: stop after the walk, leave s/r in slot 0
    // frontier schedule only (push_frontier.cu)
    int32_t *frontier;     // [n_slots][2][n] current / next frontier
    double *fval;          // [n_slots][n] residual mass taken from each frontier node this round
    double scale;          // fixed-point unit: 2^F
    double inv_scale;      // 2^-F
    int64_t work_lo;       // first work-list position of this launch (work_ids == nullptr)
    int cursor;            // PushCounter index of the work cursor this launch pulls from
    // batched engine (push_batched.cu)
    int64_t touched_stride;    // ints per slot in `touched` (dense: n; hash: the member staging list)
    TableEntry *tbl;           // [n_slots][2][tbl_cap_max] open-addressing tables (two halves: grow = move)
    int64_t tbl_cap_max;       // entries per half, power of two
    int32_t *tbl_clean;        // [n_slots][2] entries of each half known to be all-EMPTY from index 0
    double *dbg_s, *dbg_r;     // operator seam: dense s / r of the single walked seed (pre-zeroed)
    unsigned long long *centrality;   // [n] fixed-point sum over seeds of s/d_in (arcte_and_centrality), or nullptr
    double cent_scale;                // 2^38
    const double *edge_din;    // [nnz] in-degree of the target of every stored entry
    int uniform_rows;          // 1: every row of w holds one repeated value, kept in row_w
    const double *row_w;       // [n] that value (uniform_rows)
    const int32_t *from_walk;  // [n] walk label -> node id for everything that leaves the kernel, or nullptr (labels are node ids)
    // compact-state FIFO engine (push_compact.cu): `sr` holds the pairs by compact index, `queue` holds int2 {node, index}
    uint32_t *cmap;            // [n_slots][map_stride] (epoch << idx_bits) | compact index of the node in the walk of that epoch
    uint32_t *cepoch;          // [n_slots] epoch of the last walk of the slot
    int64_t map_stride;        // entries per slot in cmap (n rounded up to a multiple of 4)
    int idx_bits;              // bits of the compact index in a map entry
    int64_t ccap;              // pairs (and touched entries) per slot: a walk that touches more nodes is re-run with ccap = n
    int unit_rows;             // 1: every transition weight of row u is exactly 1/len(u): recomputed, the weight array is not read
};


// Implemented in push_batched.cu.
int batched_plan(arcte_cuda_ctx *c, int engine, int64_t n_work, int64_t *n_slots, int64_t *queue_cap);
int batched_ensure(arcte_cuda_ctx *c, int engine, int64_t n_slots, int64_t queue_cap);
void batched_fill_params(arcte_cuda_ctx *c, int engine, PushParams &P);
int batched_launch(arcte_cuda_ctx *c, int engine, const PushParams &P);

// Implemented in push_compact.cu.
int compact_launch(arcte_cuda_ctx *c, int rule, const PushParams &P);
int compact_scatter(arcte_cuda_ctx *c, const PushParams &P, int64_t nt, double *s_dev, double *r_dev);

// Implemented in push_frontier.cu.
int frontier_plan_slots(arcte_cuda_ctx *c, int64_t n_work, int64_t *n_slots);
int frontier_ensure_slots(arcte_cuda_ctx *c, int64_t n_slots);
int frontier_launch(arcte_cuda_ctx *c, PushParams P, int64_t n_work, bool retry_pass);
double frontier_scale(double rho);

#ifdef __CUDACC__
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// State pairs are gathered at random over gigabytes: keep them out of L1 (L2-only loads and
// stores) so L1 stays with what is re-read -- the FIFO ring, the touched list, node records,
// CSR rows and the few spilled registers.
//
// L2 eviction hints (experiment switches, see profiles/README.md):
//   ARCTE_HINT_STATE  state pairs are loaded/stored with an L2 evict_first policy (they stream
//                     through L2: the next use of a sector is milliseconds away);
//   ARCTE_HINT_GRAPH  node records, column indices and transition weights are loaded with an
//                     L2 evict_last policy (90 MB on the YouTube shape, re-read by every walk).
// The policies are created once per thread (createpolicy is a register-only instruction).
#ifndef ARCTE_HINT_STATE
#define ARCTE_HINT_STATE 0
#endif
#ifndef ARCTE_HINT_GRAPH
#define ARCTE_HINT_GRAPH 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#if ARCTE_HINT_STATE
__device__ __forceinline__ double2 ld_state(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(l2_policy_evict_first()) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(double2 *p, double2 v)
{
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;"
                 :: "l"(p), "d"(v.x), "d"(v.y), "l"(l2_policy_evict_first()) : "memory");
}
#elif !defined(ARCTE_STATE_L1)
__device__ __forceinline__ double2 ld_state(const double2 *p) { return __ldcg(p); }
__device__ __forceinline__ void st_state(double2 *p, double2 v) { __stcg(p, v); }
#else
__device__ __forceinline__ double2 ld_state(const double2 *p) { return *p; }
__device__ __forceinline__ void st_state(double2 *p, double2 v) { *p = v; }
#endif
#if ARCTE_HINT_GRAPH
__device__ __forceinline__ NodeInfo ld_info(const NodeInfo *p)
{
    NodeInfo v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                 : "=l"(*reinterpret_cast<unsigned long long *>(&v.d_in)),
                   "=l"(*reinterpret_cast<unsigned long long *>(&v.begin))
                 : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
__device__ __forceinline__ double ld_info_din(const NodeInfo *p)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(&p->d_in), "l"(l2_policy_evict_last()));
    return v;
}
__device__ __forceinline__ int ld_index(const int32_t *p)
{
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
__device__ __forceinline__ double ld_weight(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(l2_policy_evict_last()));
    return v;
}
#else
__device__ __forceinline__ NodeInfo ld_info(const NodeInfo *p) { return *p; }
__device__ __forceinline__ double ld_info_din(const NodeInfo *p) { return p->d_in; }
__device__ __forceinline__ int ld_index(const int32_t *p) { return *p; }
__device__ __forceinline__ double ld_weight(const double *p) { return *p; }
#endif

// Per-warp statistics live in shared memory (no registers held across the walk).
enum WarpStat { WS_PUSHES = 0, WS_EDGES, WS_ENQ, WS_MAXQ, WS_SUPPORT, WS_TOUCHED, WS_SEEDDEG, WS_MEMBERS,
                WS_EMITTED, WS_T_BEGIN, WS_COUNT };


// `x / d >= t` as the reference evaluates it (one IEEE division, similarity.py:204, :214; arcte.py:363-367)
// without paying for the fp64 division when a single-precision quotient already decides: the float quotient
// is within 4e-7 of the true one (normal operands), the guard band is 1e-5, and everything inside the band,
// denormal or out of float range goes through the exact division.  Same truth value, always.
struct Threshold {
    double t;
    float lo, hi;
};
__device__ __forceinline__ Threshold make_threshold(double t)
{
    Threshold th;
    th.t = t;
    th.lo = __double2float_rd(t * (1.0 - 1e-5));
    th.hi = __double2float_ru(t * (1.0 + 1e-5));
    return th;
}
__device__ __forceinline__ bool quot_ge(double x, double d, const Threshold &th)
{
#ifndef ARCTE_NO_QUOT_FILTER
    const float xf = __double2float_rn(x), df = __double2float_rn(d);
    if (xf >= 1.17549435e-38f && df >= 1.17549435e-38f) {
        const float q = __fdividef(xf, df);
        if (q > th.hi) return true;
        if (q < th.lo && q > 0.0f) return false;
    }
#endif
    return __ddiv_rn(x, d) >= th.t;
}

#endif  // __CUDACC__

}  // namespace arcte
