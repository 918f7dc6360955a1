// push.cu -- K3 + K4: the batched epsilon-push engine fused with the per-seed threshold
// and the stream compaction of community members.
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   drivers   fast_approximate_cumulative_pagerank_difference   eps_randomwalk/similarity.py:149-222
//             fast_approximate_personalized_pagerank            eps_randomwalk/similarity.py:11-63
//             lazy_approximate_personalized_pagerank            eps_randomwalk/similarity.py:66-146
//   rules     cumulative_pagerank_difference_limit_push         eps_randomwalk/push.py:41-64
//             pagerank_limit_push / pagerank_lazy_push          eps_randomwalk/push.py:4-38
//   worker    arcte_worker (threshold + membership)             embedding/arcte/arcte.py:328-376
//             guarded variants                                  embedding/arcte/arcte.py:121-152, 233-264
//
// Execution model.  One warp owns one seed at a time ("slot"): a dense, always-zero-
// between-seeds array of interleaved {s, r} pairs over all n nodes in HBM, a list of the
// nodes it touched (sparse reset instead of the reference's O(n) clears) and a FIFO
// ring.  The warp replays the reference's queue discipline EXACTLY -- same pop order,
// same CSR-order enqueue of the neighbours that pass r/d_in >= eps, duplicates kept --
// with the 32 lanes spread over the pushed node's neighbour list.  Every fp64 operation
// is a single IEEE rounding (__dmul_rn/__dadd_rn/__ddiv_rn, no FMA contraction), so s and
// r are bit-identical to numpy's and the thresholded support is identical, not merely
// within the push error bound.  Thousands of slots are in flight per GPU; seeds are
// pulled from a degree-descending work list through one atomic counter.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "primitives.cuh"
#include "push.cuh"

namespace arcte {


// Walk state of the warp for the current seed (all lanes hold the same values).
struct Walk {
    unsigned head, tail;   // FIFO positions (monotone; ring index = pos & mask)
    int nt;                // touched count
};

// Experiment switches of the one-entry-per-iteration kernel, both measured as losses on the bench shapes and off
// (profiles/r2_fifo_variants.md):
//   ARCTE_FIFO_UNIFORM   rows of one repeated transition weight: one product per push, no weight loads
//                        (the extra gather of the row weight costs more than the coalesced weight loads: +5 %)
//   ARCTE_FIFO_EDGE_DIN  in-degree of the target read coalesced with the row instead of gathered per neighbour
//                        (the gathers mostly hit L2, the extra 8 bytes per stored entry do not: +4 %)
#ifndef ARCTE_FIFO_UNIFORM
#define ARCTE_FIFO_UNIFORM 0
#endif
#ifndef ARCTE_FIFO_EDGE_DIN
#define ARCTE_FIFO_EDGE_DIN 0
#endif
#ifndef ARCTE_PUSH_UNROLL
#define ARCTE_PUSH_UNROLL 1
#endif
constexpr int kPushUnroll = ARCTE_PUSH_UNROLL;  // neighbour chunks (of 32) in flight per warp

// One push of node u (row [begin, begin+len), state pair su just read), then -- when
// `scan` -- the enqueue scan over the same neighbours (similarity.py:194-196 / :214-216).
// Returns false when the FIFO ring would overflow (the caller aborts and retries the seed).
// ws: this warp's statistics row in shared memory (updated by lane 0 only).
template <int RULE>
__device__ __forceinline__ bool push_node(const PushParams &P, double2 *__restrict__ sr,
                                          int32_t *__restrict__ touched, int32_t *__restrict__ queue,
                                          Walk &wk, unsigned long long *ws, int u, double2 su, unsigned begin,
                                          unsigned len, const Threshold &eps, bool scan, int lane, unsigned lt)
{
    double c;
    if (RULE == ARCTE_RULE_ABSORBING) {
        c = __dmul_rn(P.one_minus_rho, su.y);            // push.py:57
        if (lane == 0) st_state(&sr[u], make_double2(su.x, 0.0));  // push.py:60
    } else if (RULE == ARCTE_RULE_PAGERANK) {
        const double a = __dmul_rn(P.rho, su.y);          // push.py:9
        c = __dmul_rn(P.one_minus_rho, su.y);             // push.py:10
        if (lane == 0) st_state(&sr[u], make_double2(__dadd_rn(su.x, a), 0.0));  // push.py:13-14
    } else {
        const double a = __dmul_rn(P.rho, su.y);          // push.py:29
        c = __dmul_rn(P.lazy_b, su.y);                    // push.py:30
        const double keep = __dmul_rn(P.lazy_c, su.y);    // push.py:31
        if (lane == 0) st_state(&sr[u], make_double2(__dadd_rn(su.x, a), keep));  // push.py:34-35
    }
    if (lane == 0) {
        ws[WS_PUSHES] += 1;
        ws[WS_EDGES] += len;
    }
    __syncwarp();
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const int32_t *__restrict__ idx = P.indices + begin;
#if ARCTE_EDGE_RECORDS
    const double2 *__restrict__ wd = P.wd + begin;
#else
    const double *__restrict__ wgt = P.w + begin;
#endif
    // every row of an unweighted graph holds one repeated transition weight (transition.cu: k_row_uniform):
    // one product per push instead of one weight load per stored entry
#if ARCTE_FIFO_UNIFORM
    const bool uni = P.uniform_rows != 0;
    const double p_uni = uni ? __dmul_rn(c, P.row_w[u]) : 0.0;
#else
    const bool uni = false;
    const double p_uni = 0.0;
#endif
#if ARCTE_FIFO_EDGE_DIN
    const double *__restrict__ edin = P.edge_din + begin;   // in-degree of every entry's target, coalesced
#endif
    for (unsigned base = 0; base < len; base += 32 * kPushUnroll) {
        // phase 1: neighbour ids, transition weights and target in-degrees of up to kPushUnroll chunks
        int v[kPushUnroll];
        double p[kPushUnroll];
        double dv[kPushUnroll];
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k) {
            const unsigned j = base + k * 32 + lane;
            v[k] = -1;
            if (j < len) {
                v[k] = ld_index(idx + j);
#if ARCTE_EDGE_RECORDS
                const double2 e = wd[j];  // one coalesced 16-byte read: weight and the target's in-degree
                p[k] = __dmul_rn(c, e.x);
                dv[k] = e.y;
#else
                p[k] = uni ? p_uni : __dmul_rn(c, ld_weight(wgt + j));
#if ARCTE_FIFO_EDGE_DIN
                dv[k] = edin[j];
#endif
#endif
            }
        }
        // phase 2: their state pairs (independent gathers, all in flight)
        double2 o[kPushUnroll];
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k) {
            if (v[k] >= 0) {
                o[k] = ld_state(&sr[v[k]]);
#if !ARCTE_EDGE_RECORDS && !ARCTE_FIFO_EDGE_DIN
                dv[k] = ld_info_din(&P.info[v[k]]);
#endif
            }
        }
        // phase 3: update and store (neighbours of one node are distinct: no ordering needed)
        unsigned f_new = 0, f_enq = 0;  // per-lane flag bits, one per chunk
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k) {
            if (v[k] >= 0) {
                double2 nw;
                if (RULE == ARCTE_RULE_ABSORBING) {
                    nw.x = __dadd_rn(o[k].x, p[k]);  // push.py:63
                    nw.y = __dadd_rn(o[k].y, p[k]);  // push.py:64
                } else {
                    nw.x = o[k].x;
                    nw.y = __dadd_rn(o[k].y, p[k]);  // push.py:17 / :38
                }
                st_state(&sr[v[k]], nw);
                if ((o[k].x == 0.0 && o[k].y == 0.0) && (nw.x != 0.0 || nw.y != 0.0)) f_new |= 1u << k;
                if (scan && quot_ge(nw.y, dv[k], eps)) f_enq |= 1u << k;  // similarity.py:194 / :214
            }
        }
        // phase 4: ordered appends (CSR order = chunk order, then lane order).  Every node
        // touched above is recorded BEFORE the ring can report overflow, so an aborted walk
        // can always be undone through the touched list.
#pragma unroll
        for (int k = 0; k < kPushUnroll; ++k) {
            if (base + k * 32 >= len) break;  // warp-uniform
            const bool is_new = (f_new >> k) & 1u;
            const unsigned m_new = __ballot_sync(kFull, is_new);
            if (is_new) touched[wk.nt + __popc(m_new & lt)] = v[k];
            wk.nt += __popc(m_new);
        }
        if (scan) {
#pragma unroll
            for (int k = 0; k < kPushUnroll; ++k) {
                if (base + k * 32 >= len) break;  // warp-uniform
                const bool enq = (f_enq >> k) & 1u;
                const unsigned m_enq = __ballot_sync(kFull, enq);
                const unsigned cnt = __popc(m_enq);
                if (cnt) {
                    if (wk.tail - wk.head + cnt > (unsigned)P.queue_cap) return false;
                    if (enq) queue[(wk.tail + __popc(m_enq & lt)) & qmask] = v[k];
                    wk.tail += cnt;
                    if (lane == 0) ws[WS_ENQ] += cnt;
                }
            }
        }
    }
    if (lane == 0 && wk.tail - wk.head > ws[WS_MAXQ]) ws[WS_MAXQ] = wk.tail - wk.head;
    __syncwarp();
    return true;
}

#ifndef ARCTE_EPI_UNROLL
#define ARCTE_EPI_UNROLL 2
#endif
constexpr int kEpiUnroll = ARCTE_EPI_UNROLL;  // touched entries per lane in flight in the threshold sweep

// K4 of one finished walk: threshold, membership, sparse reset of the slot and the per-warp totals.
template <int RULE>
__device__ __forceinline__ void threshold_and_emit(const PushParams &P, double2 *__restrict__ sr,
                                                   int32_t *__restrict__ touched, const Walk &wk,
                                                   unsigned long long *ws, unsigned long long *wtot, int pos, int seed,
                                                   int lane, unsigned lt)
{
    // ---------------- K4: threshold + membership (arcte.py:352-376) ----------------
    const NodeInfo si = P.info[seed];
    const int base_size = (int)si.len + 1;  // np.append(adjacent_nodes[n], n), arcte.py:358
    bool emit = true;
    if (RULE != ARCTE_RULE_ABSORBING) {
        // arcte.py:129-133 / :241-245: intersect1d(base, support).size >= base.size
        int inside = 0;
        for (unsigned j = lane; j < si.len; j += 32) {
            const int v = P.indices[si.begin + j];
            inside += (v != seed && ld_state(&sr[v]).x != 0.0);
        }
        inside = warp_sum_i(inside) + (ld_state(&sr[seed]).x != 0.0 ? 1 : 0);
        emit = inside >= base_size;
    }
    double tau_v = 0.0;
    if (emit) {
        double q = __ddiv_rn(ld_state(&sr[seed]).x, si.d_in);
        for (unsigned j0 = 0; j0 < si.len; j0 += 64) {
            int v[2];
            double2 o[2];
            double d[2];
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                const unsigned j = j0 + k2 * 32 + lane;
                v[k2] = j < si.len ? P.indices[si.begin + j] : -1;
            }
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2)
                if (v[k2] >= 0) {
                    o[k2] = ld_state(&sr[v[k2]]);
                    d[k2] = ld_info_din(&P.info[v[k2]]);
                }
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2)
                if (v[k2] >= 0) q = fmin(q, __ddiv_rn(o[k2].x, d[k2]));  // arcte.py:355-356
        }
        tau_v = warp_min(q);  // arcte.py:359-360
    }
    const Threshold tau = make_threshold(tau_v);
    // One sweep over the touched list: count the support, keep (compacted in place, in
    // list order) the nodes with s/d_in >= tau -- arcte.py:363-367, searchsorted 'left' --
    // and zero the state (the sparse form of s[:]=0; r[:]=0, arcte.py:337-338).
    int m = 0, support = 0;
    for (int i0 = 0; i0 < wk.nt; i0 += 32 * kEpiUnroll) {
        int x[kEpiUnroll];
        double sx[kEpiUnroll], dx[kEpiUnroll];
#pragma unroll
        for (int k2 = 0; k2 < kEpiUnroll; ++k2) {
            const int i = i0 + k2 * 32 + lane;
            x[k2] = i < wk.nt ? touched[i] : -1;
        }
#pragma unroll
        for (int k2 = 0; k2 < kEpiUnroll; ++k2) {
            if (x[k2] >= 0) {
                sx[k2] = ld_state(&sr[x[k2]]).x;
                dx[k2] = ld_info_din(&P.info[x[k2]]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < kEpiUnroll; ++k2) {
            if (i0 + k2 * 32 >= wk.nt) break;  // warp-uniform
            bool in_sup = false, pass = false;
            if (x[k2] >= 0) {
                st_state(&sr[x[k2]], make_double2(0.0, 0.0));
                in_sup = sx[k2] != 0.0;
                pass = emit && in_sup && quot_ge(sx[k2], dx[k2], tau);
                // arcte.pyx:164-191: centrality += s / d_in over the support of every seed, in 2^-38 fixed
                // point so that the sum does not depend on the order the seeds finish in
                if (P.centrality && in_sup)
                    atomicAdd(&P.centrality[P.from_walk ? P.from_walk[x[k2]] : x[k2]], __double2ull_rn(__dmul_rn(__ddiv_rn(sx[k2], dx[k2]), P.cent_scale)));
            }
            support += __popc(__ballot_sync(kFull, in_sup));
            if (emit) {
                const unsigned mp = __ballot_sync(kFull, pass);
                if (pass) touched[m + __popc(mp & lt)] = x[k2];
                m += __popc(mp);
            }
        }
    }
    __syncwarp();
    emit = emit && (m > base_size);  // arcte.py:370
    bool write = false;
    if (emit) {
        int64_t off;
        if (P.retry_pass && P.seg_count[pos] > 0) {
            off = P.seg_offset[pos];  // offset was assigned in the pass that overflowed
        } else {
            unsigned long long o = 0;
            if (lane == 0) o = atomicAdd(&P.counters[PC_MEMBER_CURSOR], (unsigned long long)m);
            off = (int64_t)__shfl_sync(kFull, o, 0);
        }
        write = off + m <= P.member_cap;
        if (write)
            for (int i = lane; i < m; i += 32)
                P.members[off + i] = P.from_walk ? P.from_walk[touched[i]] : touched[i];  // arcte.py:372-376
        if (lane == 0) {
            P.seg_count[pos] = m;
            P.seg_offset[pos] = off;
            if (!write) {
                const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                P.retry_list[r] = pos;
            }
        }
    } else if (lane == 0) {
        P.seg_count[pos] = 0;
        P.seg_offset[pos] = 0;
    }
    __syncwarp();

    if (lane == 0 && (!emit || write)) {  // a seed whose members did not fit is re-run and counted then
        wtot[WS_PUSHES] += ws[WS_PUSHES];
        wtot[WS_EDGES] += ws[WS_EDGES];
        wtot[WS_ENQ] += ws[WS_ENQ];
        if (ws[WS_MAXQ] > wtot[WS_MAXQ]) wtot[WS_MAXQ] = ws[WS_MAXQ];
        wtot[WS_SUPPORT] += support;
        wtot[WS_TOUCHED] += wk.nt;
        wtot[WS_SEEDDEG] += si.len;
        if (emit) {
            wtot[WS_MEMBERS] += m;
            wtot[WS_EMITTED] += 1;
        }
    }
    __syncwarp();
}

#ifndef ARCTE_PUSH_MIN_BLOCKS
#define ARCTE_PUSH_MIN_BLOCKS 4
#endif

template <int RULE>
__global__ void __launch_bounds__(256, ARCTE_PUSH_MIN_BLOCKS)
k_push_threshold(const PushParams P)
{
    // per-warp statistics: [0] totals of the finished seeds, [1] the seed being walked
    __shared__ unsigned long long wstat[8][2][WS_COUNT];
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= P.n_slots) return;
    double2 *__restrict__ sr = P.sr + slot * P.n;
    int32_t *__restrict__ touched = P.touched + slot * P.n;
    int32_t *__restrict__ queue = P.queue + slot * P.queue_cap;
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    unsigned long long *wtot = wstat[threadIdx.x >> 5][0];
    unsigned long long *ws = wstat[threadIdx.x >> 5][1];
    if (lane < WS_COUNT) wtot[lane] = 0ull;
    __syncwarp();

    if (lane == 0) {
        unsigned long long t_begin;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        wtot[WS_T_BEGIN] = t_begin;
    }

    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&P.counters[PC_WORK_CURSOR], 1ull);
        k = __shfl_sync(kFull, k, 0);
        if ((int64_t)k >= P.n_work) break;
        const int pos = P.work_ids ? P.work_ids[k] : (int)k;
        const int seed = P.work_seed[pos];
        const Threshold eps = make_threshold(P.work_eps[pos]);

        Walk wk;
        wk.head = wk.tail = 0;
        wk.nt = 1;
        if (lane < WS_COUNT) ws[lane] = 0ull;

        // similarity.py:176-177 (absorbing) / :26, :84 (pagerank variants)
        double2 su = make_double2(RULE == ARCTE_RULE_ABSORBING ? 1.0 : 0.0, 1.0);
        if (lane == 0) {
            st_state(&sr[seed], su);
            touched[0] = seed;
        }
        __syncwarp();

        // The walk.  Iteration 0 is the seed's unconditional push ("Do one push for free",
        // similarity.py:183-196); every later iteration pops the FIFO head and pushes it if
        // its residual still passes (similarity.py:199-216).  (Fetching the next entry's node
        // record one pop ahead was measured and dropped: the launch is bound by random DRAM
        // sectors, not by this dependent load -- profiles/README.md.)
        int u = seed;
        NodeInfo iu = ld_info(&P.info[seed]);
        bool first = true, ok = true;
        for (;;) {
            if (first || quot_ge(su.y, iu.d_in, eps)) {  // similarity.py:204
                ok = push_node<RULE>(P, sr, touched, queue, wk, ws, u, su, iu.begin, iu.len, eps, true, lane, lt);
                if (!ok) break;
            }
            first = false;
            if (RULE == ARCTE_RULE_LAZY) {  // similarity.py:106-114 / :134-142: repeated pushes, no scan
                su = ld_state(&sr[u]);
                while (quot_ge(su.y, iu.d_in, eps)) {
                    push_node<RULE>(P, sr, touched, queue, wk, ws, u, su, iu.begin, iu.len, eps, false, lane, lt);
                    su = ld_state(&sr[u]);
                }
            }
            if (wk.head == wk.tail) break;
            u = queue[wk.head & qmask];
            iu = ld_info(&P.info[u]);
            wk.head += 1;
            su = ld_state(&sr[u]);
        }

        if (P.debug_keep) {
            if (lane == 0) {
                P.counters[PC_PUSHES] = ws[WS_PUSHES];
                P.counters[PC_TOUCHED] = (unsigned long long)wk.nt;
                P.counters[PC_OVERFLOW_SEEDS] = ok ? 0ull : 1ull;
            }
            return;
        }

        if (!ok) {
            // FIFO ring too small: undo and hand the seed to the retry pass
            __syncwarp();   // the touched list was appended to by other lanes just before the abort
            for (int i = lane; i < wk.nt; i += 32) st_state(&sr[touched[i]], make_double2(0.0, 0.0));
            if (lane == 0) {
                const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                P.retry_list[r] = pos;
                atomicAdd(&P.counters[PC_QOVERFLOW], 1ull);
                P.seg_count[pos] = -1;
            }
            __syncwarp();
            continue;
        }

        threshold_and_emit<RULE>(P, sr, touched, wk, ws, wtot, pos, seed, lane, lt);
    }

    if (lane == 0 && !P.debug_keep) {
        atomicAdd(&P.counters[PC_PUSHES], wtot[WS_PUSHES]);
        atomicAdd(&P.counters[PC_EDGES], wtot[WS_EDGES]);
        atomicAdd(&P.counters[PC_ENQUEUES], wtot[WS_ENQ]);
        atomicMax(&P.counters[PC_MAXQ], wtot[WS_MAXQ]);
        atomicAdd(&P.counters[PC_SUPPORT], wtot[WS_SUPPORT]);
        atomicAdd(&P.counters[PC_TOUCHED], wtot[WS_TOUCHED]);
        atomicAdd(&P.counters[PC_SEEDDEG], wtot[WS_SEEDDEG]);
        atomicAdd(&P.counters[PC_MEMBERS], wtot[WS_MEMBERS]);
        atomicAdd(&P.counters[PC_EMITTED], wtot[WS_EMITTED]);
        // occupancy of the launch: when this warp started/finished and how long it was busy
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicMin(&P.counters[PC_T_START], wtot[WS_T_BEGIN]);
        atomicMax(&P.counters[PC_T_END], t_end);
        atomicAdd(&P.counters[PC_T_BUSY], t_end - wtot[WS_T_BEGIN]);
    }
}

// ---- the pipelined walk loop ----------------------------------------------------------------------------
// Same queue discipline, same arithmetic, same results as k_push_threshold; what changes is WHEN the loads are
// issued.  One pop of the loop above is a chain of four dependent round trips -- ring entry, {node record, state
// pair}, the row, the neighbours' state pairs -- and with the ~7 neighbours of a typical pushed node a warp has a
// handful of sectors in flight at any time: the launch ran at the LATENCY of that chain (about 6,300 cycles per
// push, profiles/r2_pipelined_fifo.md), not at any throughput limit.  Here
//   * the next 32 ring entries and their node records sit in a shared-memory window per warp (one coalesced ring
//     read and one record gather per 32 pops; their state pairs and rows are prefetched into L2 on the way);
//   * while the neighbours of pop k are being gathered, the state pair of pop k+1 and the first 32 entries of its
//     row are already on their way (speculatively: a pop that fails the threshold wastes the row read);
//   * a pair that was loaded early is patched in registers when push k rewrites it (pop k+1 a neighbour of
//     pop k, or the same node again), so the value tested is exactly the one the reference tests;
//   * inside a long row the next 32 column indices and weights are loaded before the current ones are consumed.
// Per push the chain is one gather deep.
#ifndef ARCTE_PIPE_MIN_BLOCKS
#define ARCTE_PIPE_MIN_BLOCKS 3
#endif
constexpr int kPipeWarps = 8;   // warps per CTA

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

struct PipeWindow {
    int32_t u[kPipeWarps][32];
    NodeInfo info[kPipeWarps][32];
};

template <int RULE>
__device__ __forceinline__ bool push_node_pipe(const PushParams &P, double2 *__restrict__ sr,
                                               int32_t *__restrict__ touched, int32_t *__restrict__ queue, Walk &wk,
                                               unsigned long long *ws, int u, double2 su, unsigned begin, unsigned len,
                                               const Threshold *eps_sh, int lane, unsigned lt, bool pre, int pv, double pw,
                                               int nu, double2 *patch, bool &patched)
{
    double c;
    double2 su_new;
    if (RULE == ARCTE_RULE_ABSORBING) {
        c = __dmul_rn(P.one_minus_rho, su.y);                       // push.py:57
        su_new = make_double2(su.x, 0.0);                           // push.py:60
    } else {
        const double a = __dmul_rn(P.rho, su.y);                    // push.py:9
        c = __dmul_rn(P.one_minus_rho, su.y);                       // push.py:10
        su_new = make_double2(__dadd_rn(su.x, a), 0.0);             // push.py:13-14
    }
    // The pair of the next pop is in flight (loaded early): nothing here may wait for it, so a value this push
    // gives it goes to `patch` (shared memory) and replaces the loaded one when the pop is taken.
    if (lane == 0) {
        st_state(&sr[u], su_new);
        ws[WS_PUSHES] += 1;
        ws[WS_EDGES] += len;
        if (nu == u) *patch = su_new;   // the same node is popped again next
    }
    patched = nu == u;
    __syncwarp();
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const int32_t *__restrict__ idx = P.indices + begin;
    const double *__restrict__ wgt = P.w + begin;
    int v = -1;
    double wv = 0.0;
    if (pre) {
        v = pv;
        wv = pw;
    } else if ((unsigned)lane < len) {
        v = ld_index(idx + lane);
        wv = ld_weight(wgt + lane);
    }
    for (unsigned base = 0; base < len; base += 32) {
        // the neighbours' state pairs and in-degrees: independent gathers, all in flight
        double2 o = make_double2(0.0, 0.0);
        double dv = 1.0;
        if (v >= 0) {
            o = ld_state(&sr[v]);
            dv = ld_info_din(&P.info[v]);
        }
        // the next 32 entries of a long row, before the gathers are consumed
        int v2 = -1;
        double wv2 = 0.0;
        const unsigned j2 = base + 32 + lane;
        if (j2 < len) {
            v2 = ld_index(idx + j2);
            wv2 = ld_weight(wgt + j2);
        }
        bool is_new = false, enq = false;
        double2 nw = o;
        if (v >= 0) {
            const double p = __dmul_rn(c, wv);
            if (RULE == ARCTE_RULE_ABSORBING) nw.x = __dadd_rn(o.x, p);   // push.py:63
            nw.y = __dadd_rn(o.y, p);                                     // push.py:64 / :17
            st_state(&sr[v], nw);
            is_new = (o.x == 0.0 && o.y == 0.0) && (nw.x != 0.0 || nw.y != 0.0);
            enq = quot_ge(nw.y, dv, *eps_sh);                             // similarity.py:194 / :214
        }
        if (nu >= 0) {   // the pair that was loaded early is rewritten by this push: take the value just stored
            const unsigned hit = __ballot_sync(kFull, v == nu);
            if (v == nu) *patch = nw;
            patched = patched || hit != 0u;
        }
        // ordered appends (CSR order = lane order).  Every node touched above is recorded BEFORE the ring can
        // report overflow, so an aborted walk can always be undone through the touched list.
        const unsigned m_new = __ballot_sync(kFull, is_new);
        if (is_new) touched[wk.nt + __popc(m_new & lt)] = v;
        wk.nt += __popc(m_new);
        const unsigned m_enq = __ballot_sync(kFull, enq);
        const unsigned cnt = __popc(m_enq);
        if (cnt) {
            if (wk.tail - wk.head + cnt > (unsigned)P.queue_cap) return false;
            if (enq) queue[(wk.tail + __popc(m_enq & lt)) & qmask] = v;
            wk.tail += cnt;
            if (lane == 0) ws[WS_ENQ] += cnt;
        }
        v = v2;
        wv = wv2;
    }
    if (lane == 0 && wk.tail - wk.head > ws[WS_MAXQ]) ws[WS_MAXQ] = wk.tail - wk.head;
    __syncwarp();
    return true;
}

template <int RULE>
__global__ void __launch_bounds__(32 * kPipeWarps, ARCTE_PIPE_MIN_BLOCKS)
k_push_pipelined(const PushParams P)
{
    __shared__ unsigned long long wstat[kPipeWarps][2][WS_COUNT];
    __shared__ PipeWindow win;
    __shared__ double2 patch_buf[kPipeWarps];
    __shared__ Threshold eps_buf[kPipeWarps];   // walk-constant, read where it is used: registers are for the loads in flight
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int wid = threadIdx.x >> 5;
    const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= P.n_slots) return;
    double2 *__restrict__ sr = P.sr + slot * P.n;
    int32_t *__restrict__ touched = P.touched + slot * P.n;
    int32_t *__restrict__ queue = P.queue + slot * P.queue_cap;
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    unsigned long long *wtot = wstat[wid][0];
    unsigned long long *ws = wstat[wid][1];
    int32_t *win_u = win.u[wid];
    NodeInfo *win_i = win.info[wid];
    if (lane < WS_COUNT) wtot[lane] = 0ull;
    __syncwarp();
    if (lane == 0) {
        unsigned long long t_begin;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        wtot[WS_T_BEGIN] = t_begin;
    }

    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(&P.counters[PC_WORK_CURSOR], 1ull);
        k = __shfl_sync(kFull, k, 0);
        if ((int64_t)k >= P.n_work) break;
        const int pos = P.work_ids ? P.work_ids[k] : (int)k;
        const int seed = P.work_seed[pos];
        if (lane == 0) eps_buf[wid] = make_threshold(P.work_eps[pos]);
        const Threshold *eps_sh = &eps_buf[wid];

        Walk wk;
        wk.head = wk.tail = 0;
        wk.nt = 1;
        if (lane < WS_COUNT) ws[lane] = 0ull;
        double2 su = make_double2(RULE == ARCTE_RULE_ABSORBING ? 1.0 : 0.0, 1.0);   // similarity.py:176-177 / :26
        if (lane == 0) {
            st_state(&sr[seed], su);
            touched[0] = seed;
        }
        __syncwarp();

        int u = seed;
        unsigned u_begin, u_len;
        double u_din;
        {
            const NodeInfo iu = ld_info(&P.info[seed]);
            u_begin = iu.begin;
            u_len = iu.len;
            u_din = iu.d_in;
        }
        bool first = true, ok = true, pre = false;
        int pv = -1;
        double pw = 0.0;
        unsigned wbase = 0, wn = 0;   // the window holds ring positions [wbase, wbase + wn)
        // Ring positions [pos, min(pos + 32, tail)) -> window: ids, node records; their state pairs and rows -> L2.
        auto refill = [&](unsigned at) {
            __syncwarp();
            wbase = at;
            wn = min(32u, wk.tail - at);
            if ((unsigned)lane < wn) {
                const int x = queue[(at + lane) & qmask];
                const NodeInfo ix = ld_info(&P.info[x]);
                prefetch_l2(&sr[x]);
                win_u[lane] = x;
                win_i[lane] = ix;
                prefetch_l2(P.indices + ix.begin);
                prefetch_l2(P.w + ix.begin);
            }
            __syncwarp();
        };
        for (;;) {
            const bool do_push = first || quot_ge(su.y, u_din, *eps_sh);   // similarity.py:183 / :204
            const bool have_next = wk.head != wk.tail;
            int nu = -1, nv = -1;
            double2 nsu = su;
            double nwt = 0.0;
            if (have_next) {
                if (wk.head - wbase >= wn) refill(wk.head);
                nu = win_u[wk.head - wbase];
                nsu = ld_state(&sr[nu]);
                const NodeInfo ni = win_i[wk.head - wbase];
                if ((unsigned)lane < ni.len) {
                    nv = ld_index(P.indices + ni.begin + lane);
                    nwt = ld_weight(P.w + ni.begin + lane);
                }
            }
            bool patched = false;
            if (do_push) {
                ok = push_node_pipe<RULE>(P, sr, touched, queue, wk, ws, u, su, u_begin, u_len, eps_sh, lane, lt, pre, pv,
                                          pw, nu, &patch_buf[wid], patched);
                if (!ok) break;
            }
            if (patched) nsu = patch_buf[wid];   // (push_node_pipe ends with __syncwarp)
            first = false;
            bool npre = have_next;
            if (!have_next) {
                if (wk.head == wk.tail) break;
                refill(wk.head);   // the ring was empty before this push: nothing could be fetched ahead
                nu = win_u[0];
                nsu = ld_state(&sr[nu]);
                npre = false;
            }
            {   // the record of the pop being taken: from the window, not from registers held across the push
                const NodeInfo ni = win_i[wk.head - wbase];
                u_begin = ni.begin;
                u_len = ni.len;
                u_din = ni.d_in;
            }
            wk.head += 1;
            u = nu;
            su = nsu;
            pre = npre;
            pv = nv;
            pw = nwt;
        }

        if (P.debug_keep) {
            if (lane == 0) {
                P.counters[PC_PUSHES] = ws[WS_PUSHES];
                P.counters[PC_TOUCHED] = (unsigned long long)wk.nt;
                P.counters[PC_OVERFLOW_SEEDS] = ok ? 0ull : 1ull;
            }
            return;
        }
        if (!ok) {
            // FIFO ring too small: undo and hand the seed to the retry pass
            __syncwarp();
            for (int i = lane; i < wk.nt; i += 32) st_state(&sr[touched[i]], make_double2(0.0, 0.0));
            if (lane == 0) {
                const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                P.retry_list[r] = pos;
                atomicAdd(&P.counters[PC_QOVERFLOW], 1ull);
                P.seg_count[pos] = -1;
            }
            __syncwarp();
            continue;
        }
        threshold_and_emit<RULE>(P, sr, touched, wk, ws, wtot, pos, seed, lane, lt);
    }

    if (lane == 0 && !P.debug_keep) {
        atomicAdd(&P.counters[PC_PUSHES], wtot[WS_PUSHES]);
        atomicAdd(&P.counters[PC_EDGES], wtot[WS_EDGES]);
        atomicAdd(&P.counters[PC_ENQUEUES], wtot[WS_ENQ]);
        atomicMax(&P.counters[PC_MAXQ], wtot[WS_MAXQ]);
        atomicAdd(&P.counters[PC_SUPPORT], wtot[WS_SUPPORT]);
        atomicAdd(&P.counters[PC_TOUCHED], wtot[WS_TOUCHED]);
        atomicAdd(&P.counters[PC_SEEDDEG], wtot[WS_SEEDDEG]);
        atomicAdd(&P.counters[PC_MEMBERS], wtot[WS_MEMBERS]);
        atomicAdd(&P.counters[PC_EMITTED], wtot[WS_EMITTED]);
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicMin(&P.counters[PC_T_START], wtot[WS_T_BEGIN]);
        atomicMax(&P.counters[PC_T_END], t_end);
        atomicAdd(&P.counters[PC_T_BUSY], t_end - wtot[WS_T_BEGIN]);
    }
}

// ---- small helper kernels ---------------------------------------------------------------
__global__ void k_build_work(int64_t n_work, int shard_rank, int shard_count,
                             const int32_t *__restrict__ seeds, int32_t *__restrict__ work_seed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_work) work_seed[i] = seeds[shard_rank + i * shard_count];  // arcte.py:19-23
}

__global__ void k_to_walk_labels(int64_t n_work, const int32_t *__restrict__ work_seed, const int32_t *__restrict__ to_walk,
                                 int32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_work) out[i] = to_walk[work_seed[i]];
}

// The graph as the FIFO-schedule kernels see it: in walk labels when they were built (transition.cu, K2c).
static void fill_graph_params(const arcte_cuda_ctx *c, PushParams &P, bool walk_labels)
{
    P.info = walk_labels ? c->walk_info.as<NodeInfo>() : c->node_info.as<NodeInfo>();
    P.indices = walk_labels ? c->walk_indices.as<int32_t>() : c->indices.as<int32_t>();
    P.row_w = walk_labels ? c->walk_row_w.as<double>() : c->row_w.as<double>();
    P.from_walk = walk_labels ? c->from_walk.as<int32_t>() : nullptr;
    P.unit_rows = c->unit_rows ? 1 : 0;
}

constexpr int kCompactWarpsPerSm = 8 * ARCTE_COMPACT_MIN_BLOCKS;       // 6 CTAs of 8 warps at 40 registers (push.cuh)
constexpr int kCompactWarpsPerSmHi = 8 * ARCTE_COMPACT_MIN_BLOCKS_HI;  // 8 CTAs at 32 registers: graphs of short rows
constexpr int64_t kShortRowMean = 16;                                   // stored entries per row below which a graph takes the latter

static int bits_for(int64_t v)   // bits needed for values 0 .. v-1 (at least 1)
{
    int b = 1;
    while ((int64_t(1) << b) < v) ++b;
    return b;
}

static void fill_compact_params(const arcte_cuda_ctx *c, PushParams &P)
{
    P.cmap = c->slots.cmap.as<uint32_t>();
    P.cepoch = c->slots.cepoch.as<uint32_t>();
    P.map_stride = c->slots.map_stride;
    P.idx_bits = bits_for(c->n);
    P.ccap = c->slots.ccap;
    P.sr = c->slots.sr.as<double2>();
    P.touched = c->slots.touched.as<int32_t>();
    P.queue = c->slots.queue.as<int32_t>();
    P.queue_cap = c->slots.queue_cap;
    const char *env = getenv("ARCTE_CUDA_COMPACT_EPOCH_BITS");   // tests: force the epoch field to wrap early
    if (env && atoi(env) >= 2 && 32 - atoi(env) >= P.idx_bits) P.idx_bits = 32 - atoi(env);
}

// sort key of a walk: the high word of its (positive) epsilon-effective, ascending = longest walks first
__global__ void k_eps_keys(int64_t n_work, const double *__restrict__ work_eps, uint32_t *__restrict__ keys,
                           uint32_t *__restrict__ ids)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    keys[i] = (uint32_t)((unsigned long long)__double_as_longlong(work_eps[i]) >> 32);
    ids[i] = (uint32_t)i;
}

__global__ void k_gather_eps(int64_t n_work, int shard_rank, int shard_count,
                             const double *__restrict__ eps_global, double *__restrict__ work_eps)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_work) work_eps[i] = eps_global[shard_rank + i * shard_count];
}

__global__ void k_split_and_reset(int64_t n, int64_t n_touched, double2 *__restrict__ sr,
                                  const int32_t *__restrict__ touched, double *__restrict__ s_out,
                                  double *__restrict__ r_out, int phase, double inv_scale,
                                  const int32_t *__restrict__ from_walk)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 0) {
        if (i < n) {
            const int64_t o = from_walk ? from_walk[i] : i;
            if (inv_scale == 0.0) {
                const double2 v = sr[i];
                s_out[o] = v.x;
                r_out[o] = v.y;
            } else {  // frontier schedule: unsigned 64-bit fixed point
                const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(sr)[i];
                s_out[o] = __dmul_rn(__ull2double_rn(v.x), inv_scale);
                r_out[o] = __dmul_rn(__ull2double_rn(v.y), inv_scale);
            }
        }
    } else if (i < n_touched) {
        sr[touched[i]] = make_double2(0.0, 0.0);
    }
}

static inline unsigned grid_for(int64_t items, int block) { return (unsigned)((items + block - 1) / block); }

int compute_eps_effective(arcte_cuda_ctx *c, double epsilon, const int32_t *dev_seeds,
                          int64_t n_seeds, double *dev_eps_out);

// Allocate (or re-shape) the slot pool.  Dense engine: states are zeroed once here and every seed leaves its slot
// all-zero again, so the pool is reused across extractions as is.  Compact engine (push_compact.cu): `sr` holds the
// pairs by compact index and is never cleared; the index map starts all-zero with epoch 0 and the FIFO rings hold
// 8-byte entries.
static int ensure_slots(arcte_cuda_ctx *c, int64_t want_slots, int64_t queue_cap, bool compact, int64_t ccap = 0)
{
    SlotPool &sp = c->slots;
    if (!compact || ccap <= 0 || ccap > c->n) ccap = c->n;
    const bool same_state = (sp.n == c->n && sp.n_slots >= want_slots && sp.compact == compact && sp.ccap == ccap);
    const size_t qe = compact ? sizeof(int2) : sizeof(int32_t);
    if (!same_state) {
        dev_free(sp.sr);
        dev_free(sp.touched);
        dev_free(sp.queue);
        dev_free(sp.cmap);
        dev_free(sp.cepoch);
        sp.n_slots = sp.queue_slots = 0;
        ARCTE_TRY(dev_reserve(sp.sr, sizeof(double2) * (size_t)want_slots * (size_t)ccap));
        ARCTE_TRY(dev_reserve(sp.touched, sizeof(int32_t) * (size_t)want_slots * (size_t)ccap));
        sp.ccap = ccap;
        if (compact) {
            sp.map_stride = (c->n + 3) / 4 * 4;
            ARCTE_TRY(dev_reserve(sp.cmap, sizeof(uint32_t) * (size_t)want_slots * (size_t)sp.map_stride));
            ARCTE_TRY(dev_reserve(sp.cepoch, sizeof(uint32_t) * (size_t)want_slots));
            ARCTE_CUDA_TRY(cudaMemsetAsync(sp.cmap.p, 0, sizeof(uint32_t) * (size_t)want_slots * (size_t)sp.map_stride, c->stream));
            ARCTE_CUDA_TRY(cudaMemsetAsync(sp.cepoch.p, 0, sizeof(uint32_t) * (size_t)want_slots, c->stream));
        } else {
            ARCTE_CUDA_TRY(cudaMemsetAsync(sp.sr.p, 0, sizeof(double2) * (size_t)want_slots * (size_t)c->n, c->stream));
        }
        sp.n = c->n;
        sp.n_slots = want_slots;
        sp.compact = compact;
    }
    if (sp.queue_cap != queue_cap || sp.queue_slots < want_slots) {
        dev_free(sp.queue);
        ARCTE_TRY(dev_reserve(sp.queue, qe * (size_t)sp.n_slots * (size_t)queue_cap));
        sp.queue_cap = queue_cap;
        sp.queue_slots = sp.n_slots;
    }
    return ARCTE_OK;
}

static int64_t pow2_at_least(int64_t v)
{
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void fill_rule_constants(PushParams &P, double rho)
{
    const double lazy = 0.5;  // similarity.py:75 default, the only value the reference uses
    P.rho = rho;
    P.one_minus_rho = 1.0 - rho;
    P.lazy_b = (1.0 - rho) * (1.0 - lazy);
    P.lazy_c = (1.0 - rho) * lazy;
}

static int launch_push(arcte_cuda_ctx *c, int rule, const PushParams &P)
{
    if (P.cmap) return compact_launch(c, rule, P);
    const unsigned grid = grid_for(P.n_slots * 32, 256);
    // experiment switch, off: the pipelined loop needs 80 registers (24 walks per SM) and ends up slower than the
    // plain loop at 32 walks per SM (profiles/r2_pipelined_fifo.md)
    static const bool pipelined = getenv("ARCTE_CUDA_PIPELINED") && !strcmp(getenv("ARCTE_CUDA_PIPELINED"), "1");
    switch (rule) {
    case ARCTE_RULE_ABSORBING:
        if (pipelined) k_push_pipelined<ARCTE_RULE_ABSORBING><<<grid, 32 * kPipeWarps, 0, c->stream>>>(P);
        else k_push_threshold<ARCTE_RULE_ABSORBING><<<grid, 256, 0, c->stream>>>(P);
        break;
    case ARCTE_RULE_PAGERANK:
        if (pipelined) k_push_pipelined<ARCTE_RULE_PAGERANK><<<grid, 32 * kPipeWarps, 0, c->stream>>>(P);
        else k_push_threshold<ARCTE_RULE_PAGERANK><<<grid, 256, 0, c->stream>>>(P);
        break;
    case ARCTE_RULE_LAZY: k_push_threshold<ARCTE_RULE_LAZY><<<grid, 256, 0, c->stream>>>(P); break;
    default: set_error("unknown push rule"); return ARCTE_E_ARG;
    }
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

// Slot-pool geometry for this graph: how many walks can be in flight.
// Compact engine: pairs per slot.  The supports of the bench shapes stay below 10^5 nodes (the reference scales
// epsilon with the seed's degree, arcte.py:26-50), so a quarter of a million pairs per walk is plenty; the rare walk
// that needs more is re-run with room for all n nodes (extract_shard's retry pass).
static int64_t compact_cap(const arcte_cuda_ctx *c, bool full)
{
    int64_t cap = int64_t(1) << 18;
    const char *env = getenv("ARCTE_CUDA_COMPACT_CAP");
    if (env && atoll(env) > 0) cap = atoll(env);
    return (full || cap > c->n) ? c->n : cap;
}

static int plan_slots(arcte_cuda_ctx *c, int64_t n_work, int64_t *n_slots, int64_t *queue_cap, bool compact,
                      bool full_cap = false)
{
    const bool short_rows = c->nnz < kShortRowMean * c->n;
    const int wps = c->warps_per_sm > 0 ? c->warps_per_sm
                                        : (compact ? (short_rows ? kCompactWarpsPerSmHi : kCompactWarpsPerSm) : 32);
    const int64_t ccap = compact ? compact_cap(c, full_cap) : c->n;
    int64_t want = (int64_t)c->sm_count * wps;
    want = ((want + 7) / 8) * 8;
    int64_t qcap = c->queue_cap_cfg > 0 ? c->queue_cap_cfg : (c->n < 65536 ? c->n : 65536);
    if (c->queue_cap_cfg <= 0 && qcap < 8192) qcap = 8192;
    if (qcap < 64) qcap = 64;
    qcap = pow2_at_least(qcap);
    if (want > ((n_work + 7) / 8) * 8) want = ((n_work + 7) / 8) * 8;
    if (want < 8) want = 8;
    // keep whatever is already allocated if it is big enough (avoids re-zeroing)
    if (c->slots.n == c->n && c->slots.n_slots >= want && c->slots.queue_slots >= want &&
        c->slots.queue_cap == qcap && c->slots.compact == compact && c->slots.ccap == ccap) {
        *n_slots = want;
        *queue_cap = qcap;
        return ARCTE_OK;
    }
    size_t free_b = 0, total_b = 0;
    ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->slots.sr.bytes + c->slots.touched.bytes + c->slots.queue.bytes + c->slots.cmap.bytes;
    const int pct = c->mem_percent > 0 ? c->mem_percent : 60;
    const double budget = (double)free_b * pct / 100.0;
    const double per_slot = 20.0 * ccap + (compact ? 4.0 * c->n + 8.0 * qcap : 4.0 * qcap);
    int64_t fit = (int64_t)(budget / per_slot);
    fit = (fit / 8) * 8;
    if (fit < 8) {
        set_error("not enough device memory for 8 walk states of this graph");
        return ARCTE_E_NOMEM;
    }
    if (want > fit) want = fit;
    *n_slots = want;
    *queue_cap = qcap;
    return ARCTE_OK;
}

// The pools of the dense FIFO / frontier engines and of the batched engines are each sized for most of the
// free memory: when an extraction switches engine, the pool of the other family goes first.
static void release_other_pools(arcte_cuda_ctx *c, bool batched, int engine)
{
    if (batched) {
        if (c->bpool.mode == engine) return;   // already laid out for this engine
        dev_free(c->slots.sr);
        dev_free(c->slots.touched);
        dev_free(c->slots.queue);
        dev_free(c->slots.cmap);
        dev_free(c->slots.cepoch);
        dev_free(c->slots.frontier);
        dev_free(c->slots.fval);
        c->slots = SlotPool();
    } else if (c->bpool.mode >= 0 && !(c->slots.n == c->n && c->slots.n_slots > 0)) {
        dev_free(c->bpool.tbl);
        dev_free(c->bpool.stage);
        dev_free(c->bpool.clean);
        dev_free(c->bpool.queue);
        c->bpool = BatchedPool();
    }
}

// Which engine of the FIFO schedule walks this call (include/arcte_cuda.h, ARCTE_ENGINE_*).
static int resolve_engine(const arcte_cuda_ctx *c, int rule)
{
    int e = c->engine;
    if (rule != ARCTE_RULE_ABSORBING) {   // the batched engines implement the absorbing rule only
        const char *env = getenv("ARCTE_CUDA_ENGINE");
        if (e == ARCTE_ENGINE_AUTO && env && !strcmp(env, "compact")) e = ARCTE_ENGINE_FIFO_COMPACT;
        if (e == ARCTE_ENGINE_AUTO && !env && c->nnz >= (int64_t(1) << 22)) e = ARCTE_ENGINE_FIFO_COMPACT;
        return e == ARCTE_ENGINE_FIFO_COMPACT ? e : ARCTE_ENGINE_FIFO_DENSE;
    }
    if (e == ARCTE_ENGINE_AUTO) {
        const char *env = getenv("ARCTE_CUDA_ENGINE");
        if (env && !strcmp(env, "fifo")) e = ARCTE_ENGINE_FIFO_DENSE;
        else if (env && !strcmp(env, "dense")) e = ARCTE_ENGINE_BATCHED_DENSE;
        else if (env && !strcmp(env, "hash")) e = ARCTE_ENGINE_BATCHED_HASH;
        else if (env && !strcmp(env, "compact")) e = ARCTE_ENGINE_FIFO_COMPACT;
    }
    // measured (profiles/r2_engines.md, r2_compact_state.md): graphs with millions of stored entries are bound by
    // DRAM accesses and go to the compact-state engine (YouTube shape 605 vs 706 ms, Flickr shape 110 vs 133 / 158 ms);
    // smaller graphs are bound by their longest walk, where the shorter dependent chain of the dense layouts wins:
    // the batched direct-mapped engine on medium graphs (BA(150000,3) 13.7 vs 15.6 ms), the FIFO engine on tiny ones
    if (e == ARCTE_ENGINE_AUTO) {
        if (c->nnz >= (int64_t(1) << 22)) e = ARCTE_ENGINE_FIFO_COMPACT;
        else e = (c->n >= 4096 && c->n <= (int64_t(1) << 19)) ? ARCTE_ENGINE_BATCHED_DENSE : ARCTE_ENGINE_FIFO_DENSE;
    }
    return e;
}

int extract_shard(arcte_cuda_ctx *c, int rule, double rho, double epsilon, int shard_rank,
                  int shard_count, const double *host_eps_override)
{
    if (!c->have_transition) { set_error("extract: no transition matrix (call set_graph)"); return ARCTE_E_ARG; }
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) {
        set_error("extract: bad shard");
        return ARCTE_E_ARG;
    }
    cudaStream_t st = c->stream;
    const int64_t S = c->n_seeds > shard_rank ? (c->n_seeds - shard_rank + shard_count - 1) / shard_count : 0;
    c->n_segments = S;
    c->n_members = 0;
    c->shard_rank = shard_rank;
    c->shard_count = shard_count;
    c->have_segments = false;
    c->have_features = false;
    arcte_cuda_stats &stt = c->stats;
    stt.n_seeds_shard = S;
    stt.pushes = stt.edge_touches = stt.enqueues = stt.max_queue = stt.support = stt.touched = 0;
    stt.seed_degree = stt.members = stt.emitted = stt.retries = 0;
    stt.ms_push = 0.0;
    stt.alg_bytes_push = 0.0;

    const size_t S1 = (size_t)(S > 0 ? S : 1);
    ARCTE_TRY(dev_reserve(c->work_seed, sizeof(int32_t) * S1));
    ARCTE_TRY(dev_reserve(c->work_eps, sizeof(double) * S1));
    ARCTE_TRY(dev_reserve(c->seg_count, sizeof(int32_t) * S1));
    ARCTE_TRY(dev_reserve(c->seg_offset, sizeof(int64_t) * S1));
    ARCTE_TRY(dev_reserve(c->retry_list, sizeof(int32_t) * S1));
    ARCTE_TRY(dev_reserve(c->counters, sizeof(int64_t) * PC_COUNT));
    if (S == 0) {
        c->have_segments = true;
        return ARCTE_OK;
    }

    // ---- K2b: work list + epsilon-effective ----
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev0, st));
    k_build_work<<<grid_for(S, 256), 256, 0, st>>>(S, shard_rank, shard_count, c->seeds.as<int32_t>(),
                                                   c->work_seed.as<int32_t>());
    ++stt.launches;
    if (host_eps_override) {
        ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(double) * (size_t)c->n_seeds));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[2].p, host_eps_override,
                                       sizeof(double) * (size_t)c->n_seeds, cudaMemcpyHostToDevice, st));
        k_gather_eps<<<grid_for(S, 256), 256, 0, st>>>(S, shard_rank, shard_count,
                                                       c->scratch[2].as<double>(), c->work_eps.as<double>());
        ++stt.launches;
    } else {
        ARCTE_TRY(compute_eps_effective(c, epsilon, c->work_seed.as<int32_t>(), S, c->work_eps.as<double>()));
    }
    // ---- K2c': work order.  A walk's cost follows 1/epsilon-effective closely (rank correlation 0.997 on the bench
    // shape, profiles/r2_work_order.json, tools/work_order_study.py) and NOT the seed's count: a seed of count 2 next to a hub gets a tiny epsilon
    // and one of the longest walks of the run.  Handing the seeds out by ascending epsilon (longest walks first) keeps
    // the end of the launch -- and, with several GPUs, the wait for the slowest rank -- free of stragglers.
    const bool frontier = c->schedule == ARCTE_SCHEDULE_FRONTIER;
    const bool ordered = !frontier && S > 1 && !(getenv("ARCTE_CUDA_WORK_ORDER") && !strcmp(getenv("ARCTE_CUDA_WORK_ORDER"), "0"));
    if (ordered) {
        for (int k = 0; k < 4; ++k) ARCTE_TRY(dev_reserve(c->scratch[k], sizeof(uint32_t) * S1));
        ARCTE_TRY(dev_reserve(c->work_order, sizeof(int32_t) * S1));
        k_eps_keys<<<grid_for(S, 256), 256, 0, st>>>(S, c->work_eps.as<double>(), c->scratch[0].as<uint32_t>(),
                                                     c->scratch[2].as<uint32_t>());
        ++stt.launches;
        bool second = false;
        ARCTE_TRY(radix_sort_pairs(c->scratch[0].as<uint32_t>(), c->scratch[2].p, c->scratch[1].as<uint32_t>(), c->scratch[3].p,
                                   S, 32, 4, c->scratch[4], c->scratch[5], c->scratch[6], st, &second, &stt.launches));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->work_order.p, second ? c->scratch[3].p : c->scratch[2].p, sizeof(int32_t) * (size_t)S,
                                       cudaMemcpyDeviceToDevice, st));
    }
    ARCTE_CUDA_TRY(cudaEventRecord(c->ev1, st));

    // ---- slot pool + member buffer ----
    if (frontier && rule != ARCTE_RULE_ABSORBING) {
        set_error("extract: the frontier schedule implements the absorbing rule (arcte) only");
        return ARCTE_E_ARG;
    }
    int64_t n_slots = 0, qcap = 0;
    int engine = frontier ? ARCTE_ENGINE_FIFO_DENSE : resolve_engine(c, rule);
    if (c->centrality_acc && engine != ARCTE_ENGINE_FIFO_COMPACT) engine = ARCTE_ENGINE_FIFO_DENSE;
    const bool compact = engine == ARCTE_ENGINE_FIFO_COMPACT;
    const bool batched = !frontier && !compact && engine != ARCTE_ENGINE_FIFO_DENSE;
    if (c->centrality_acc && (frontier || rule != ARCTE_RULE_ABSORBING)) {
        set_error("centrality: absorbing rule, FIFO schedule only");
        return ARCTE_E_ARG;
    }
    const bool walk_labels = c->walk_labels_valid && !frontier;   // the frontier schedule keeps the caller's labels
    release_other_pools(c, batched, engine);
    if (frontier) {
        ARCTE_TRY(frontier_plan_slots(c, S, &n_slots));
        ARCTE_TRY(frontier_ensure_slots(c, n_slots));
    } else if (batched) {
        ARCTE_TRY(batched_plan(c, engine, S, &n_slots, &qcap));
        ARCTE_TRY(batched_ensure(c, engine, n_slots, qcap));
    } else {
        ARCTE_TRY(plan_slots(c, S, &n_slots, &qcap, compact));
        ARCTE_TRY(ensure_slots(c, n_slots, qcap, compact, compact_cap(c, false)));
    }
    stt.engine = frontier ? -2 : engine;
    if (c->member_cap == 0) {
        // default: a tenth of what is free now, never more than the n_seeds x n worst case
        size_t free_b = 0, total_b = 0;
        ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        int64_t cap = c->member_cap_cfg;
        if (cap <= 0) {
            cap = (int64_t)(free_b / 10 / sizeof(int32_t));
            const double worst = (double)S * (double)c->n;
            if ((double)cap > worst) cap = (int64_t)worst;
            if (cap < (1 << 20)) cap = 1 << 20;
        }
        ARCTE_TRY(dev_reserve(c->members, sizeof(int32_t) * (size_t)cap));
        c->member_cap = cap;
    }

    PushParams P{};
    P.n = c->n;
    P.w = c->w.as<double>();
    P.wd = c->edge_wd.as<double2>();
    P.edge_din = c->edge_din.as<double>();
    P.uniform_rows = c->uniform_rows ? 1 : 0;
    P.work_seed = c->work_seed.as<int32_t>();
    if (walk_labels) {
        ARCTE_TRY(dev_reserve(c->work_seed_w, sizeof(int32_t) * S1));
        k_to_walk_labels<<<grid_for(S, 256), 256, 0, st>>>(S, c->work_seed.as<int32_t>(), c->to_walk.as<int32_t>(),
                                                           c->work_seed_w.as<int32_t>());
        ++stt.launches;
        P.work_seed = c->work_seed_w.as<int32_t>();
    }
    P.work_eps = c->work_eps.as<double>();
    P.work_ids = ordered ? c->work_order.as<int32_t>() : nullptr;
    P.n_work = S;
    P.retry_pass = 0;
    P.sr = c->slots.sr.as<double2>();
    P.touched = c->slots.touched.as<int32_t>();
    P.queue = c->slots.queue.as<int32_t>();
    P.queue_cap = c->slots.queue_cap;
    P.n_slots = n_slots;
    P.seg_count = c->seg_count.as<int32_t>();
    P.seg_offset = c->seg_offset.as<int64_t>();
    P.members = c->members.as<int32_t>();
    P.member_cap = c->member_cap;
    P.retry_list = c->retry_list.as<int32_t>();
    P.counters = c->counters.as<unsigned long long>();
    P.debug_keep = 0;
    // arcte.py:109: the lazy worker walks with lazy_rho = 0.5 rho / (1 - 0.5 rho)
    fill_rule_constants(P, rule == ARCTE_RULE_LAZY ? (rho * (0.5)) / (1.0 - (0.5 * rho)) : rho);
    if (frontier) {
        P.scale = frontier_scale(rho);
        P.inv_scale = 1.0 / P.scale;
    }
    if (batched) batched_fill_params(c, engine, P);
    fill_graph_params(c, P, walk_labels);
    if (compact) fill_compact_params(c, P);
    P.centrality = c->centrality_acc;
    P.cent_scale = 274877906944.0;   // 2^38

    ARCTE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, sizeof(int64_t) * PC_COUNT, st));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->counters.as<int64_t>() + PC_T_START, 0xff, sizeof(int64_t), st));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->seg_count.p, 0, sizeof(int32_t) * (size_t)S, st));
    const cudaEvent_t p0 = c->pk0, p1 = c->pk1;  // owned by the context: nothing to release on the error paths
    ARCTE_CUDA_TRY(cudaEventRecord(p0, st));
    if (frontier) ARCTE_TRY(frontier_launch(c, P, S, false));
    else if (batched) ARCTE_TRY(batched_launch(c, engine, P));
    else ARCTE_TRY(launch_push(c, rule, P));
    ARCTE_CUDA_TRY(cudaEventRecord(p1, st));

    int64_t hc[PC_COUNT];
    ARCTE_CUDA_TRY(cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    stt.ms_seeds += ms;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, p0, p1));
    stt.ms_push = ms;
    stt.n_slots = n_slots;
    // mean busy time of a walk state / span of the launch (1.0 = no idle tail)
    stt.rounds = 0;
    stt.slot_utilisation = hc[PC_T_END] > hc[PC_T_START]
                               ? (double)hc[PC_T_BUSY] / ((double)n_slots * (double)(hc[PC_T_END] - hc[PC_T_START]))
                               : 0.0;

    // ---- retry passes: seeds whose FIFO ring or member range did not fit ----
    int rounds = 0;
    while (hc[PC_OVERFLOW_SEEDS] > 0) {
        const int64_t n_retry = hc[PC_OVERFLOW_SEEDS];
        stt.retries += n_retry;
        if (++rounds > 12) { set_error("extract: retry passes did not converge"); return ARCTE_E_OVERFLOW; }
        if (compact && hc[PC_TOVERFLOW] > 0 && c->slots.ccap < c->n) {
            // walks that touched more nodes than a slot has pairs: fewer slots with room for all n nodes
            int64_t ls = 0, lq = 0;
            ARCTE_TRY(plan_slots(c, n_retry, &ls, &lq, true, true));
            ARCTE_TRY(ensure_slots(c, ls, lq, true, c->n));
            fill_compact_params(c, P);
        }
        if (batched && rounds == 1) {
            // seeds the batched engine gave up on (ring, table region or member range too small) are re-run
            // by the dense FIFO engine: same results, and its rings grow as far as memory allows
            int64_t ls = 0, lq = 0;
            ARCTE_TRY(plan_slots(c, n_retry, &ls, &lq, false));
            ARCTE_TRY(ensure_slots(c, ls, lq, false));
            P.sr = c->slots.sr.as<double2>();
            P.touched = c->slots.touched.as<int32_t>();
            P.queue = c->slots.queue.as<int32_t>();
            P.queue_cap = c->slots.queue_cap;
        }
        // members: grow to the exact demand seen so far (+ headroom for the seeds still to run)
        if (hc[PC_MEMBER_CURSOR] > c->member_cap || hc[PC_QOVERFLOW] > 0 || hc[PC_TOVERFLOW] > 0) {
            int64_t new_cap = hc[PC_MEMBER_CURSOR] + (hc[PC_QOVERFLOW] + hc[PC_TOVERFLOW] > 0 ? c->member_cap : 0);
            if (new_cap > c->member_cap) {
                DevBuf nb;
                ARCTE_TRY(dev_reserve(nb, sizeof(int32_t) * (size_t)new_cap));
                ARCTE_CUDA_TRY(cudaMemcpyAsync(nb.p, c->members.p, sizeof(int32_t) * (size_t)c->member_cap,
                                               cudaMemcpyDeviceToDevice, st));
                ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
                dev_free(c->members);
                c->members = nb;
                c->member_cap = new_cap;
            }
        }
        int64_t r_slots = ((n_retry + 7) / 8) * 8;
        if (frontier) r_slots = n_slots;
        else if (r_slots > c->slots.queue_slots) r_slots = c->slots.queue_slots;
        int64_t r_qcap = c->slots.queue_cap;
        if (hc[PC_QOVERFLOW] > 0) {
            // a bigger ring for fewer walks: same bytes first, then grow
            r_qcap = c->slots.queue_cap * 8;
            size_t free_b = 0, total_b = 0;
            ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            free_b += c->slots.queue.bytes;
            const double qe = compact ? 8.0 : 4.0;   // bytes per ring entry
            while (r_slots > 8 && (double)r_slots * r_qcap * qe > 0.8 * (double)free_b) r_slots -= 8;
            if ((double)r_slots * r_qcap * qe > 0.8 * (double)free_b) {
                set_error("extract: FIFO ring cannot be grown further (out of device memory)");
                return ARCTE_E_OVERFLOW;
            }
            dev_free(c->slots.queue);
            ARCTE_TRY(dev_reserve(c->slots.queue, (size_t)qe * (size_t)r_slots * (size_t)r_qcap));
            c->slots.queue_cap = r_qcap;
            c->slots.queue_slots = r_slots;  // fewer, larger rings until the next plan_slots
        }
        // retry list -> scratch (the kernel appends to retry_list again)
        ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(int32_t) * (size_t)n_retry));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[0].p, c->retry_list.p, sizeof(int32_t) * (size_t)n_retry,
                                       cudaMemcpyDeviceToDevice, st));
        // keep accumulated counters except the cursors that restart
        int64_t zero = 0;
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->counters.as<int64_t>() + PC_WORK_CURSOR, &zero, sizeof(zero),
                                       cudaMemcpyHostToDevice, st));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->counters.as<int64_t>() + PC_OVERFLOW_SEEDS, &zero, sizeof(zero),
                                       cudaMemcpyHostToDevice, st));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->counters.as<int64_t>() + PC_QOVERFLOW, &zero, sizeof(zero),
                                       cudaMemcpyHostToDevice, st));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->counters.as<int64_t>() + PC_TOVERFLOW, &zero, sizeof(zero),
                                       cudaMemcpyHostToDevice, st));
        P.work_ids = c->scratch[0].as<int32_t>();
        P.n_work = n_retry;
        P.retry_pass = 1;
        P.queue = c->slots.queue.as<int32_t>();
        P.queue_cap = c->slots.queue_cap;
        P.n_slots = r_slots;
        P.members = c->members.as<int32_t>();
        P.member_cap = c->member_cap;
        ARCTE_CUDA_TRY(cudaEventRecord(p0, st));
        if (frontier) ARCTE_TRY(frontier_launch(c, P, n_retry, true));
        else ARCTE_TRY(launch_push(c, rule, P));   // batched engines too: see above
        ARCTE_CUDA_TRY(cudaEventRecord(p1, st));
        ARCTE_CUDA_TRY(cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
        ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, p0, p1));
        stt.ms_push += ms;
    }

    if (getenv("ARCTE_CUDA_PROFILE")) {  // per-phase clock sums of instrumented kernel builds
        fprintf(stderr, "[arcte] profile counters:");
        for (int k = PC_PROF0; k <= PC_PROF19; ++k) fprintf(stderr, " %lld", (long long)hc[k]);
        fprintf(stderr, "\n");
    }
    stt.pushes = hc[PC_PUSHES];
    stt.edge_touches = hc[PC_EDGES];
    stt.enqueues = hc[PC_ENQUEUES];
    stt.max_queue = hc[PC_MAXQ];
    stt.support = hc[PC_SUPPORT];
    stt.touched = hc[PC_TOUCHED];
    stt.seed_degree = hc[PC_SEEDDEG];
    stt.members = hc[PC_MEMBERS];
    stt.emitted = hc[PC_EMITTED];
    stt.rounds = hc[PC_ROUNDS];
    // sum_pushes(24 + 52 deg(u)) + sum_seeds(32 |supp| + 12 deg(seed) + 12 |comm|)
    stt.alg_bytes_push = 24.0 * stt.pushes + 52.0 * stt.edge_touches + 32.0 * stt.support +
                         12.0 * stt.seed_degree + 12.0 * stt.members;
    c->n_members = hc[PC_MEMBER_CURSOR];
    c->have_segments = true;
    return ARCTE_OK;
}

// Operator seam: one seed, dense s and r back to the host.
int push_single(arcte_cuda_ctx *c, int rule, int64_t seed, double rho, double eps_eff, double *host_s,
                double *host_r, int64_t *n_push)
{
    if (!c->have_transition) { set_error("push: no transition matrix (call set_graph)"); return ARCTE_E_ARG; }
    if (seed < 0 || seed >= c->n) { set_error("push: seed out of range"); return ARCTE_E_ARG; }
    cudaStream_t st = c->stream;
    const bool frontier = c->schedule == ARCTE_SCHEDULE_FRONTIER;
    if (frontier && rule != ARCTE_RULE_ABSORBING) {
        set_error("push: the frontier schedule implements the absorbing rule only");
        return ARCTE_E_ARG;
    }
    int engine = frontier ? ARCTE_ENGINE_FIFO_DENSE : resolve_engine(c, rule);
    const bool compact = engine == ARCTE_ENGINE_FIFO_COMPACT;
    ARCTE_TRY(dev_reserve(c->counters, sizeof(int64_t) * PC_COUNT));
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(int32_t) * 4));
    ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(double) * 4));
    ARCTE_TRY(dev_reserve(c->scratch[3], sizeof(double) * 2 * (size_t)c->n));
    const int32_t seed32 = (int32_t)seed;
    const bool walk_labels = c->walk_labels_valid && !frontier;
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[0].p, &seed32, sizeof(seed32), cudaMemcpyHostToDevice, st));
    if (walk_labels) {
        k_to_walk_labels<<<1, 32, 0, st>>>(1, c->scratch[0].as<int32_t>(), c->to_walk.as<int32_t>(), c->scratch[0].as<int32_t>() + 1);
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[2].p, &eps_eff, sizeof(eps_eff), cudaMemcpyHostToDevice, st));
    double *s_dev = c->scratch[3].as<double>();
    double *r_dev = s_dev + c->n;

    int64_t cap = 0;
    bool pools_ready = false, full_cap = false;
    for (int attempt = 0;; ++attempt) {
        const bool batched = !frontier && !compact && engine != ARCTE_ENGINE_FIFO_DENSE;
        if (!pools_ready) {
            int64_t n_slots = 0, qcap = 0;
            if (attempt == 0) release_other_pools(c, batched, engine);
            if (frontier) {
                ARCTE_TRY(frontier_plan_slots(c, 1, &n_slots));
                ARCTE_TRY(frontier_ensure_slots(c, n_slots));
            } else if (batched) {
                ARCTE_TRY(batched_plan(c, engine, 1, &n_slots, &qcap));
                ARCTE_TRY(batched_ensure(c, engine, n_slots, qcap));
            } else {
                ARCTE_TRY(plan_slots(c, 1, &n_slots, &qcap, compact, full_cap));
                ARCTE_TRY(ensure_slots(c, n_slots, qcap, compact, compact_cap(c, full_cap)));
            }
            cap = c->slots.queue_cap;
            pools_ready = true;
        }
        ARCTE_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, sizeof(int64_t) * PC_COUNT, st));
        PushParams P{};
        P.n = c->n;
        P.w = c->w.as<double>();
        P.wd = c->edge_wd.as<double2>();
        P.edge_din = c->edge_din.as<double>();
        P.uniform_rows = c->uniform_rows ? 1 : 0;
        P.work_seed = c->scratch[0].as<int32_t>() + (walk_labels ? 1 : 0);
        P.work_eps = c->scratch[2].as<double>();
        P.n_work = 1;
        P.sr = c->slots.sr.as<double2>();
        P.touched = c->slots.touched.as<int32_t>();
        P.queue = c->slots.queue.as<int32_t>();
        P.queue_cap = cap;
        P.n_slots = 1;
        P.counters = c->counters.as<unsigned long long>();
        P.debug_keep = 1;
        fill_rule_constants(P, rho);
        fill_graph_params(c, P, walk_labels);
        if (compact) fill_compact_params(c, P);
        double inv_scale = 0.0;
        const bool hash = batched;   // both batched engines write the dense vectors themselves and clean up
        if (frontier) {
            P.scale = frontier_scale(rho);
            P.inv_scale = inv_scale = 1.0 / P.scale;
            P.cursor = PC_WORK_CURSOR;
            ARCTE_TRY(frontier_launch(c, P, 1, false));
        } else if (batched) {
            batched_fill_params(c, engine, P);
            fill_graph_params(c, P, walk_labels);
            P.dbg_s = s_dev;
            P.dbg_r = r_dev;
            if (hash) ARCTE_CUDA_TRY(cudaMemsetAsync(s_dev, 0, sizeof(double) * 2 * (size_t)c->n, st));
            ARCTE_TRY(batched_launch(c, engine, P));
        } else {
            ARCTE_TRY(launch_push(c, rule, P));
        }
        int64_t hc[PC_COUNT];
        ARCTE_CUDA_TRY(cudaMemcpyAsync(hc, c->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
        const int64_t nt = hc[PC_TOUCHED];
        if (hc[PC_OVERFLOW_SEEDS] == 0) {
            if (compact) {
                ARCTE_TRY(compact_scatter(c, P, nt, s_dev, r_dev));
            } else if (!hash) {
                k_split_and_reset<<<grid_for(c->n, 256), 256, 0, st>>>(c->n, nt, P.sr, P.touched, s_dev, r_dev, 0, inv_scale, P.from_walk);
                ++c->stats.launches;
            }
            ARCTE_CUDA_TRY(cudaMemcpyAsync(host_s, s_dev, sizeof(double) * (size_t)c->n, cudaMemcpyDeviceToHost, st));
            ARCTE_CUDA_TRY(cudaMemcpyAsync(host_r, r_dev, sizeof(double) * (size_t)c->n, cudaMemcpyDeviceToHost, st));
        }
        if (nt > 0 && !hash && !compact) {   // the hash engine leaves its table empty itself, the compact one needs no reset
            k_split_and_reset<<<grid_for(nt, 256), 256, 0, st>>>(c->n, nt, P.sr, P.touched, s_dev, r_dev, 1, inv_scale, P.from_walk);
            ++c->stats.launches;
        }
        ARCTE_CUDA_TRY(cudaStreamSynchronize(st));
        if (hc[PC_OVERFLOW_SEEDS] == 0) {
            if (n_push) *n_push = hc[PC_PUSHES];
            return ARCTE_OK;
        }
        if (compact && hc[PC_TOVERFLOW] > 0 && !full_cap) {   // more touched nodes than a slot has pairs
            full_cap = true;
            pools_ready = false;
            continue;
        }
        if (batched) {   // ring or table region too small: the dense FIFO engine takes the seed (its ring grows below)
            engine = ARCTE_ENGINE_FIFO_DENSE;
            pools_ready = false;
            continue;
        }
        // ring too small for this seed: one walk only, so give it one big ring
        const int64_t next = cap * 8;
        if (attempt > 10) {
            set_error("push: FIFO ring cannot be grown further for this seed");
            return ARCTE_E_OVERFLOW;
        }
        if ((size_t)next * (compact ? sizeof(int2) : sizeof(int32_t)) > c->slots.queue.bytes) {
            dev_free(c->slots.queue);
            c->slots.queue_slots = 0;
            ARCTE_TRY(dev_reserve(c->slots.queue, (compact ? sizeof(int2) : sizeof(int32_t)) * (size_t)next));
            c->slots.queue_cap = next;
            c->slots.queue_slots = 1;  // the next plan_slots re-creates the per-slot rings
        }
        cap = next;
    }
}

}  // namespace arcte
