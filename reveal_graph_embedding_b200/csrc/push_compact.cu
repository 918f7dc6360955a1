// push_compact.cu -- the FIFO push engine with COMPACT walk state (ARCTE_ENGINE_FIFO_COMPACT).
//
// Same schedule, same arithmetic and the same results as push.cu (one warp replays the reference's queue
// discipline for one seed: similarity.py:149-222, push.py:41-64, threshold and membership arcte.py:328-376);
// what changes is where a walk keeps s and r.
//
// push.cu gives every walk a dense {s, r} array over all n nodes.  A walk touches a few thousand of them, each a
// random 16-byte read-modify-write that DRAM serves with a whole line, every first touch reads zeros from DRAM,
// and the threshold sweep goes back to every touched pair at random to read it and to zero it again: about four
// DRAM accesses per touched node (profiles/r2_compact_state.md).  Here
//   * the touched nodes of a walk are numbered in first-touch order (the order of the `touched` list, which the
//     reference's support/membership order already follows) and their pairs live at that number in a compact
//     array `cst`: a first touch is a plain coalesced WRITE of consecutive pairs (nothing is read), the threshold
//     sweep is a sequential read, and a re-touch lands in lines whose other pairs belong to the same walk;
//   * the only per-node array is a 4-byte index map: entry = (epoch of the walk that wrote it, compact index).
//     An entry of another epoch means "not touched by this walk", so neither the map nor the pairs are ever reset;
//     a slot's map is cleared once every 2^(32 - index bits) walks, when its epoch counter wraps;
//   * the FIFO ring carries {node, compact index}: a pop goes straight to its pair.
// The random accesses of a neighbour touch are therefore 4 bytes wide (16-32 nodes per DRAM line instead of 4-8;
// with the walk labels of transition.cu, K2c, neighbours of one hub share those lines).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "push.cuh"

namespace arcte {

struct CWalk {
    unsigned head, tail;   // FIFO positions (monotone; ring index = pos & mask)
    int nt;                // touched count = next free compact index
};

__device__ __forceinline__ uint32_t ld_map(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ void st_map(uint32_t *p, uint32_t v) { __stcg(p, v); }

// s of node v in this walk (0.0 when the walk has not touched it)
__device__ __forceinline__ double lookup_s(const uint32_t *__restrict__ map, const double2 *__restrict__ cst, int v,
                                           uint32_t tag, uint32_t imask)
{
    const uint32_t m = ld_map(map + v);
    return (m & ~imask) == tag ? ld_state(&cst[m & imask]).x : 0.0;
}

// One push of the node at compact index ui (row [begin, begin+len), pair su just read), then -- when `scan` -- the
// enqueue scan over the same neighbours (similarity.py:194-196 / :214-216).  Returns PUSH_RING / PUSH_CAP when the
// FIFO ring / the compact arrays would overflow (the caller abandons the walk: the next epoch invalidates everything
// it wrote, and the seed is re-run with more room).
enum { PUSH_OK = 0, PUSH_RING = 1, PUSH_CAP = 2 };
template <int RULE>
__device__ __forceinline__ int push_node_c(const PushParams &P, uint32_t *__restrict__ map, double2 *__restrict__ cst,
                                            int32_t *__restrict__ touched, int2 *__restrict__ queue, CWalk &wk,
                                            unsigned long long *ws, int ui, double2 su, unsigned begin, unsigned len,
                                            const Threshold &eps, bool scan, uint32_t tag, uint32_t imask, int lane,
                                            unsigned lt)
{
    double c;
    if (RULE == ARCTE_RULE_ABSORBING) {
        c = __dmul_rn(P.one_minus_rho, su.y);                                        // push.py:57
        if (lane == 0) st_state(&cst[ui], make_double2(su.x, 0.0));                  // push.py:60
    } else if (RULE == ARCTE_RULE_PAGERANK) {
        const double a = __dmul_rn(P.rho, su.y);                                     // push.py:9
        c = __dmul_rn(P.one_minus_rho, su.y);                                        // push.py:10
        if (lane == 0) st_state(&cst[ui], make_double2(__dadd_rn(su.x, a), 0.0));    // push.py:13-14
    } else {
        const double a = __dmul_rn(P.rho, su.y);                                     // push.py:29
        c = __dmul_rn(P.lazy_b, su.y);                                               // push.py:30
        const double keep = __dmul_rn(P.lazy_c, su.y);                               // push.py:31
        if (lane == 0) st_state(&cst[ui], make_double2(__dadd_rn(su.x, a), keep));   // push.py:34-35
    }
    if (lane == 0) {
        ws[WS_PUSHES] += 1;
        ws[WS_EDGES] += len;
    }
    __syncwarp();
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const int32_t *__restrict__ idx = P.indices + begin;
    const double *__restrict__ wgt = P.w + begin;
    // unit adjacency weights: every entry of the row is 1/len (transition.py:61-63), one product per push
    const double p_unit = P.unit_rows ? __dmul_rn(c, __ddiv_rn(1.0, (double)len)) : 0.0;
    for (unsigned base = 0; base < len; base += 32) {
        const unsigned j = base + lane;
        int v = -1;
        double p = 0.0, dv = 1.0;
        uint32_t m = 0;
        if (j < len) {
            v = ld_index(idx + j);
            p = P.unit_rows ? p_unit : __dmul_rn(c, ld_weight(wgt + j));
        }
        if (v >= 0) {
            m = ld_map(map + v);
            dv = ld_info_din(&P.info[v]);
        }
        const bool valid = v >= 0 && (m & ~imask) == tag;
        int vi = (int)(m & imask);
        double2 o = make_double2(0.0, 0.0);
        if (valid) o = ld_state(&cst[vi]);
        double2 nw = o;
        if (v >= 0) {
            if (RULE == ARCTE_RULE_ABSORBING) nw.x = __dadd_rn(o.x, p);   // push.py:63
            nw.y = __dadd_rn(o.y, p);                                     // push.py:64 / :17 / :38
        }
        if (valid) st_state(&cst[vi], nw);
        // first touch: the next compact indices, in CSR order = lane order (the order of the reference's support)
        const bool is_new = v >= 0 && !valid && (nw.x != 0.0 || nw.y != 0.0);
        const unsigned m_new = __ballot_sync(kFull, is_new);
        if (wk.nt + __popc(m_new) > (int)P.ccap) return PUSH_CAP;
        if (is_new) {
            vi = wk.nt + __popc(m_new & lt);
            st_map(map + v, tag | (uint32_t)vi);
            st_state(&cst[vi], nw);
            touched[vi] = v;
        }
        wk.nt += __popc(m_new);
        if (scan) {
            const bool enq = (valid || is_new) && quot_ge(nw.y, dv, eps);   // similarity.py:194 / :214
            const unsigned m_enq = __ballot_sync(kFull, enq);
            const unsigned cnt = __popc(m_enq);
            if (cnt) {
                if (wk.tail - wk.head + cnt > (unsigned)P.queue_cap) return PUSH_RING;
                if (enq) queue[(wk.tail + __popc(m_enq & lt)) & qmask] = make_int2(v, vi);
                wk.tail += cnt;
                if (lane == 0) ws[WS_ENQ] += cnt;
            }
        }
    }
    if (lane == 0 && wk.tail - wk.head > ws[WS_MAXQ]) ws[WS_MAXQ] = wk.tail - wk.head;
    __syncwarp();
    return PUSH_OK;
}

// K4 of one finished walk, by the whole warp: threshold, membership, the per-warp totals.  Nothing is reset.
template <int RULE>
__device__ __forceinline__ void threshold_and_emit_c(const PushParams &P, const uint32_t *__restrict__ map,
                                                     const double2 *__restrict__ cst, int32_t *__restrict__ touched,
                                                     int nt, const unsigned long long *ws, unsigned long long *wtot,
                                                     int pos, int seed, uint32_t tag, uint32_t imask, int lane, unsigned lt)
{
    // ---------------- K4: threshold + membership (arcte.py:352-376) ----------------
    const NodeInfo si = P.info[seed];
    const int base_size = (int)si.len + 1;   // np.append(adjacent_nodes[n], n), arcte.py:358
    bool emit = true;
    if (RULE != ARCTE_RULE_ABSORBING) {
        // arcte.py:129-133 / :241-245: intersect1d(base, support).size >= base.size
        int inside = 0;
        for (unsigned j = lane; j < si.len; j += 32) {
            const int v = P.indices[si.begin + j];
            inside += (v != seed && lookup_s(map, cst, v, tag, imask) != 0.0);
        }
        inside = warp_sum_i(inside) + (ld_state(&cst[0]).x != 0.0 ? 1 : 0);
        emit = inside >= base_size;
    }
    double tau_v = 0.0;
    if (emit) {
        double q = __ddiv_rn(ld_state(&cst[0]).x, si.d_in);
        for (unsigned j = lane; j < si.len; j += 32) {
            const int v = P.indices[si.begin + j];
            q = fmin(q, __ddiv_rn(lookup_s(map, cst, v, tag, imask), ld_info_din(&P.info[v])));   // arcte.py:355-356
        }
        tau_v = warp_min(q);   // arcte.py:359-360
    }
    const Threshold tau = make_threshold(tau_v);
    // One sequential sweep over the compact pairs: count the support, keep (compacted in place, in first-touch
    // order) the nodes with s/d_in >= tau -- arcte.py:363-367, searchsorted 'left'.
    int m = 0, support = 0;
    for (int i0 = 0; i0 < nt; i0 += 32) {
        const int i = i0 + lane;
        int x = -1;
        double sx = 0.0, dx = 1.0;
        if (i < nt) {
            x = touched[i];
            sx = ld_state(&cst[i]).x;
            dx = ld_info_din(&P.info[x]);
        }
        const bool in_sup = x >= 0 && sx != 0.0;
        const bool pass = emit && in_sup && quot_ge(sx, dx, tau);
        // arcte.pyx:164-191: centrality += s / d_in over the support of every seed, in 2^-38 fixed point so that
        // the sum does not depend on the order the seeds finish in
        if (P.centrality && in_sup)
            atomicAdd(&P.centrality[P.from_walk ? P.from_walk[x] : x],
                      __double2ull_rn(__dmul_rn(__ddiv_rn(sx, dx), P.cent_scale)));
        support += __popc(__ballot_sync(kFull, in_sup));
        if (emit) {
            const unsigned mp = __ballot_sync(kFull, pass);
            __syncwarp();   // every lane has read touched[i0 .. i0+31] before slots <= i0+31 are overwritten
            if (pass) touched[m + __popc(mp & lt)] = x;
            m += __popc(mp);
        }
    }
    __syncwarp();
    emit = emit && (m > base_size);   // arcte.py:370
    bool write = false;
    if (emit) {
        int64_t off;
        if (P.retry_pass && P.seg_count[pos] > 0) {
            off = P.seg_offset[pos];   // offset was assigned in the pass that overflowed
        } else {
            unsigned long long o = 0;
            if (lane == 0) o = atomicAdd(&P.counters[PC_MEMBER_CURSOR], (unsigned long long)m);
            off = (int64_t)__shfl_sync(kFull, o, 0);
        }
        write = off + m <= P.member_cap;
        if (write)
            for (int i = lane; i < m; i += 32)
                P.members[off + i] = P.from_walk ? P.from_walk[touched[i]] : touched[i];   // arcte.py:372-376
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
    if (lane == 0 && (!emit || write)) {   // a seed whose members did not fit is re-run and counted then
        wtot[WS_PUSHES] += ws[WS_PUSHES];
        wtot[WS_EDGES] += ws[WS_EDGES];
        wtot[WS_ENQ] += ws[WS_ENQ];
        if (ws[WS_MAXQ] > wtot[WS_MAXQ]) wtot[WS_MAXQ] = ws[WS_MAXQ];
        wtot[WS_SUPPORT] += support;
        wtot[WS_TOUCHED] += nt;
        wtot[WS_SEEDDEG] += si.len;
        if (emit) {
            wtot[WS_MEMBERS] += m;
            wtot[WS_EMITTED] += 1;
        }
    }
    __syncwarp();
}

template <int RULE, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
k_push_compact(const PushParams P)
{
    __shared__ unsigned long long wstat[8][2][WS_COUNT];
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= P.n_slots) return;
    uint32_t *__restrict__ map = P.cmap + slot * P.map_stride;
    double2 *__restrict__ cst = P.sr + slot * P.ccap;
    int32_t *__restrict__ touched = P.touched + slot * P.ccap;
    int2 *__restrict__ queue = reinterpret_cast<int2 *>(P.queue) + slot * P.queue_cap;
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const uint32_t imask = (1u << P.idx_bits) - 1u;
    const uint32_t epoch_end = 1u << (32 - P.idx_bits);
    uint32_t epoch = P.cepoch[slot];   // epoch of the last walk of this slot (0: the map is all-zero and unused)
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

        // a new epoch: everything earlier walks of this slot left in the map and in cst is void
        if (++epoch >= epoch_end) {
            uint4 *m4 = reinterpret_cast<uint4 *>(map);   // map_stride is a multiple of 4 entries
            for (int64_t i = lane; i < P.map_stride / 4; i += 32) m4[i] = make_uint4(0u, 0u, 0u, 0u);
            epoch = 1;
            __syncwarp();
        }
        const uint32_t tag = epoch << P.idx_bits;

        CWalk wk;
        wk.head = wk.tail = 0;
        wk.nt = 1;
        if (lane < WS_COUNT) ws[lane] = 0ull;
        double2 su = make_double2(RULE == ARCTE_RULE_ABSORBING ? 1.0 : 0.0, 1.0);   // similarity.py:176-177 / :26, :84
        if (lane == 0) {
            st_map(map + seed, tag);   // compact index 0
            st_state(&cst[0], su);
            touched[0] = seed;
        }
        __syncwarp();

        // The walk: iteration 0 is the seed's unconditional push (similarity.py:183-196); every later iteration pops
        // the FIFO head and pushes it if its residual still passes (similarity.py:199-216).
        int ui = 0;
        NodeInfo iu = ld_info(&P.info[seed]);
        bool first = true;
        int ok = PUSH_OK;
        for (;;) {
            if (first || quot_ge(su.y, iu.d_in, eps)) {   // similarity.py:204
                ok = push_node_c<RULE>(P, map, cst, touched, queue, wk, ws, ui, su, iu.begin, iu.len, eps, true, tag, imask,
                                       lane, lt);
                if (ok != PUSH_OK) break;
            }
            first = false;
            if (RULE == ARCTE_RULE_LAZY) {   // similarity.py:106-114 / :134-142: repeated pushes, no scan
                su = ld_state(&cst[ui]);
                while (ok == PUSH_OK && quot_ge(su.y, iu.d_in, eps)) {
                    ok = push_node_c<RULE>(P, map, cst, touched, queue, wk, ws, ui, su, iu.begin, iu.len, eps, false, tag,
                                           imask, lane, lt);
                    su = ld_state(&cst[ui]);
                }
                if (ok != PUSH_OK) break;
            }
            if (wk.head == wk.tail) break;
            const int2 e = queue[wk.head & qmask];
            wk.head += 1;
            ui = e.y;
            iu = ld_info(&P.info[e.x]);
            su = ld_state(&cst[ui]);
        }

        if (P.debug_keep) {   // operator seam: the host scatters cst through `touched` (compact_scatter)
            if (lane == 0) {
                P.counters[PC_PUSHES] = ws[WS_PUSHES];
                P.counters[PC_TOUCHED] = (unsigned long long)wk.nt;
                P.counters[PC_OVERFLOW_SEEDS] = ok == PUSH_OK ? 0ull : 1ull;
                P.counters[PC_TOVERFLOW] = ok == PUSH_CAP ? 1ull : 0ull;
                P.cepoch[slot] = epoch;
            }
            return;
        }
        if (ok != PUSH_OK) {   // FIFO ring or compact arrays too small: hand the seed to the retry pass (nothing to undo)
            if (lane == 0) {
                const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                P.retry_list[r] = pos;
                atomicAdd(&P.counters[ok == PUSH_RING ? PC_QOVERFLOW : PC_TOVERFLOW], 1ull);
                P.seg_count[pos] = -1;
            }
            __syncwarp();
            continue;
        }

        threshold_and_emit_c<RULE>(P, map, cst, touched, wk.nt, ws, wtot, pos, seed, tag, imask, lane, lt);
    }

    if (lane == 0 && !P.debug_keep) {
        P.cepoch[slot] = epoch;
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

// Operator seam: dense s and r of the one walk slot 0 holds (outputs pre-zeroed).
__global__ void k_compact_scatter(int64_t nt, const int32_t *__restrict__ touched, const double2 *__restrict__ cst,
                                  const int32_t *__restrict__ from_walk, double *__restrict__ s_out,
                                  double *__restrict__ r_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int x = touched[i];
    const int o = from_walk ? from_walk[x] : x;
    const double2 v = cst[i];
    s_out[o] = v.x;
    r_out[o] = v.y;
}

int compact_launch(arcte_cuda_ctx *c, int rule, const PushParams &P)
{
    const unsigned grid = (unsigned)((P.n_slots * 32 + 255) / 256);
    // more walk states than 6 CTAs per SM can hold: the 32-register instantiation keeps all of them resident
    const bool hi = P.n_slots > (int64_t)c->sm_count * 8 * ARCTE_COMPACT_MIN_BLOCKS;
    constexpr int LO = ARCTE_COMPACT_MIN_BLOCKS, HI = ARCTE_COMPACT_MIN_BLOCKS_HI;
    switch (rule) {
    case ARCTE_RULE_ABSORBING:
        if (hi) k_push_compact<ARCTE_RULE_ABSORBING, HI><<<grid, 256, 0, c->stream>>>(P);
        else k_push_compact<ARCTE_RULE_ABSORBING, LO><<<grid, 256, 0, c->stream>>>(P);
        break;
    case ARCTE_RULE_PAGERANK:
        if (hi) k_push_compact<ARCTE_RULE_PAGERANK, HI><<<grid, 256, 0, c->stream>>>(P);
        else k_push_compact<ARCTE_RULE_PAGERANK, LO><<<grid, 256, 0, c->stream>>>(P);
        break;
    case ARCTE_RULE_LAZY:
        if (hi) k_push_compact<ARCTE_RULE_LAZY, HI><<<grid, 256, 0, c->stream>>>(P);
        else k_push_compact<ARCTE_RULE_LAZY, LO><<<grid, 256, 0, c->stream>>>(P);
        break;
    default: set_error("unknown push rule"); return ARCTE_E_ARG;
    }
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

int compact_scatter(arcte_cuda_ctx *c, const PushParams &P, int64_t nt, double *s_dev, double *r_dev)
{
    ARCTE_CUDA_TRY(cudaMemsetAsync(s_dev, 0, sizeof(double) * (size_t)c->n, c->stream));
    ARCTE_CUDA_TRY(cudaMemsetAsync(r_dev, 0, sizeof(double) * (size_t)c->n, c->stream));
    if (nt > 0) {
        k_compact_scatter<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(nt, P.touched, P.sr, P.from_walk, s_dev, r_dev);
        ++c->stats.launches;
    }
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

}  // namespace arcte
