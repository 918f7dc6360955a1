// push_batched.cu -- K3 + K4, second generation: the reference's FIFO order, replayed exactly, with
// several queue entries per warp iteration and a compact per-walk state.
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   driver    fast_approximate_cumulative_pagerank_difference   eps_randomwalk/similarity.py:149-222
//   rule      cumulative_pagerank_difference_limit_push         eps_randomwalk/push.py:41-64
//   worker    arcte_worker (threshold + membership)             embedding/arcte/arcte.py:328-376
//
// Why a second engine.  The first one (push.cu) walks one queue entry per warp iteration against a
// dense {s, r} array per walk.  On the 1.1 M-node bench shape a push has 7.5 neighbours on average, so
// three quarters of the lanes idle, every push is a chain of four dependent memory round trips, and the
// dense arrays of thousands of walks in flight are cold DRAM sectors on every first touch plus a random
// read + reset per touched node in the threshold sweep (profiles/r1_push_youtube_full.md: 4.35x the
// algorithmic bytes through DRAM).  This engine changes three things and keeps the arithmetic and its
// order bit-for-bit:
//
//   1. BATCH.  Up to 32 consecutive queue entries whose rows hold at most kE = 64 stored entries
//      together are taken in one warp iteration: one round trip fetches all node records, one all
//      neighbour ids and transition weights, one all walk-state entries.  The distinct nodes of the
//      batch are staged in a per-warp SHARED-MEMORY cache (open addressing, node id -> {s, r, d_in});
//      the pushes are then applied to that cache strictly in queue order -- pop check, r[u] = 0,
//      neighbours in CSR order, enqueue test, ordered append (warp ballot + prefix popcount) -- so a
//      node that occurs in several rows of the batch, or is itself a later queue entry, sees exactly
//      the values the sequential reference sees.  The cache is written back once per batch.
//   2. COMPACT STATE.  Walk state lives in an open-addressing hash table per walk (32-byte entries
//      {s, r, d_in, node}: one sector, one 256-bit load or store per touch) that starts at 1024
//      entries and grows fourfold on demand by moving into the other half of the slot's region.  The
//      next walk of the slot reuses the same sectors, so light walks stay L2-resident; the threshold
//      sweep is a linear scan of the table instead of a random gather per touched node, and it leaves
//      the table empty (no separate reset pass).  Graphs whose dense state is small (n <= 2^18) keep
//      the dense layout (ENGINE_BATCHED_DENSE), where it is the more compact of the two.
//   3. HUB ROWS.  A row longer than kE is pushed alone: its column indices and weights are staged
//      through shared memory with cp.async, 256 stored entries per stage, double-buffered, while
//      the previous stage's state entries are gathered.
//
// A walk whose table would outgrow its region, or whose FIFO ring overflows, is undone and handed to
// the retry pass (push.cu's engine / a larger ring), exactly like a ring overflow in push.cu.
#include <stdlib.h>

#include "common.cuh"
#include "push.cuh"

namespace arcte {

constexpr int kU = 2;                // edge slots per lane
constexpr int kE = 32 * kU;          // stored entries per batch
constexpr int kCacheSlots = 256;     // shared-memory node cache per warp (<= 96 nodes per batch live in it)
constexpr int kWarpsPerCta = 4;
constexpr int kStage = 256;          // stored entries per cp.async stage of a hub row
constexpr int kInitLg = 10;          // a walk's table starts at 1024 entries
constexpr int kGrowLg = 2;           // and grows fourfold

struct WarpShared {
    double2 cval[kCacheSlots];       // {s, r} of the cached node
    double cdin[kCacheSlots];        // its in-degree
    int32_t ckey[kCacheSlots];       // node id or -1
    int32_t epre[33];                // exclusive prefix of the batch entries' row lengths
    int32_t eu[32];                  // batch entries: node
    uint32_t ebeg[32];               //                row begin
    int32_t eq[32];                  //                cache slot
    double erw[32];                  //                row weight (uniform rows)
    unsigned long long stat[2][WS_COUNT];  // [0] totals of finished seeds, [1] the seed being walked
};
// hub-row staging reuses the value arrays of the node cache
static_assert(2 * kStage * sizeof(int32_t) <= sizeof(double) * kCacheSlots, "stage ids fit cdin");
static_assert(2 * kStage * sizeof(double) <= sizeof(double2) * kCacheSlots, "stage weights fit cval");

// ---- 256-bit table accesses (one sector per entry), L2-coherent --------------------------------------
__device__ __forceinline__ TableEntry ld_entry(const TableEntry *p)
{
    unsigned long long a, b, c, d;
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
    TableEntry e;
    e.s = __longlong_as_double((long long)a);
    e.r = __longlong_as_double((long long)b);
    e.d_in = __longlong_as_double((long long)c);
    e.key = (int32_t)(uint32_t)d;
    e.aux = (int32_t)(d >> 32);
    return e;
}
__device__ __forceinline__ void st_entry(TableEntry *p, double s, double r, double d_in, int32_t key)
{
    asm volatile("st.global.cg.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(s)),
                 "l"((unsigned long long)__double_as_longlong(r)), "l"((unsigned long long)__double_as_longlong(d_in)),
                 "l"((unsigned long long)(uint32_t)key)
                 : "memory");
}
__device__ __forceinline__ unsigned table_hash(int v, int lg) { return ((unsigned)v * 0x9E3779B1u) >> (32 - lg); }

// Find v or claim a free entry for it.  Lanes of one warp insert DISTINCT nodes concurrently; the
// compare-and-swap on the key settles two of them reaching the same free entry.
__device__ __forceinline__ unsigned table_find_or_insert(TableEntry *T, int lg, int v, TableEntry first, TableEntry &out, bool &is_new)
{
    const unsigned mask = (1u << lg) - 1u;
    unsigned h = table_hash(v, lg);
    TableEntry e = first;  // the caller already loaded T[h] (keeps the gathers of a batch in flight together)
    for (;;) {
        if (e.key == v) { out = e; is_new = false; return h; }
        if (e.key == kEmptyKey) {
            const int old = atomicCAS(&T[h].key, kEmptyKey, v);
            if (old == kEmptyKey) { is_new = true; out = e; return h; }
            if (old == v) { out = ld_entry(T + h); is_new = false; return h; }  // not reachable: owners are distinct
        }
        h = (h + 1) & mask;
        e = ld_entry(T + h);
    }
}
// Read-only lookup.
__device__ __forceinline__ bool table_find(const TableEntry *T, int lg, int v, TableEntry &out, unsigned &pos)
{
    const unsigned mask = (1u << lg) - 1u;
    unsigned h = table_hash(v, lg);
    for (;;) {
        const TableEntry e = ld_entry(T + h);
        if (e.key == v) { out = e; pos = h; return true; }
        if (e.key == kEmptyKey) return false;
        h = (h + 1) & mask;
    }
}

// ---- per-warp shared-memory node cache ----------------------------------------------------------------
__device__ __forceinline__ int cache_insert(int32_t *ckey, int v, bool &owner)
{
    unsigned h = ((unsigned)v * 0x9E3779B1u) >> 24;
    for (;;) {
        const int old = atomicCAS(&ckey[h], -1, v);
        if (old == -1) { owner = true; return (int)h; }
        if (old == v) { owner = false; return (int)h; }
        h = (h + 1) & (kCacheSlots - 1);
    }
}

// Walk state of one slot.
struct Slot {
    // dense
    double2 *sr;
    int32_t *touched;   // dense: touched list; hash: member staging list
    // hash
    TableEntry *half[2];
    int32_t clean[2];   // entries of each half known EMPTY from 0
    TableEntry *T;      // active table
    int cur, lg;
    int64_t cap_max;
    // both
    int32_t *queue;
    int nt;             // touched nodes of the current walk
};

__device__ __forceinline__ int table_limit(int lg) { return (5 << lg) >> 3; }  // grow beyond 62.5 % load

// Makes [0, 2^lg) of `half` all-EMPTY if it is not known to be.
__device__ __forceinline__ void table_make_clean(Slot &S, int which, int lg, int lane)
{
    const int C = 1 << lg;
    if (S.clean[which] < C) {
        for (int i = S.clean[which] + lane; i < C; i += 32) st_entry(S.half[which] + i, 0.0, 0.0, 0.0, kEmptyKey);
        S.clean[which] = C;
        __syncwarp();
    }
}

// Moves the walk's table into the other half at (up to) four times the capacity.  false: the region is
// exhausted (the caller aborts the walk).
__device__ bool table_grow(Slot &S, int need, int lane)
{
    int nlg = S.lg + kGrowLg;
    while (table_limit(nlg) < need && nlg < 27) ++nlg;
    while (nlg > S.lg && ((int64_t)1 << nlg) > S.cap_max) --nlg;
    if (nlg <= S.lg || table_limit(nlg) < need) return false;
    const int other = S.cur ^ 1;
    table_make_clean(S, other, nlg, lane);
    TableEntry *Tn = S.half[other];
    const int C = 1 << S.lg;
    for (int i0 = 0; i0 < C; i0 += 32) {
        const TableEntry e = ld_entry(S.T + i0 + lane);
        if (e.key != kEmptyKey) {
            TableEntry dummy;
            bool is_new;
            const unsigned h0 = table_hash(e.key, nlg);
            const unsigned h = table_find_or_insert(Tn, nlg, e.key, ld_entry(Tn + h0), dummy, is_new);
            st_entry(Tn + h, e.s, e.r, e.d_in, e.key);
            S.T[i0 + lane].key = kEmptyKey;
        }
    }
    __syncwarp();
    S.cur = other;
    S.T = Tn;
    S.lg = nlg;
    return true;
}

// Leaves the active table all-EMPTY (undo of an aborted walk).
__device__ void table_clear(Slot &S, int lane)
{
    const int C = 1 << S.lg;
    for (int i = lane; i < C; i += 32) S.T[i].key = kEmptyKey;
    __syncwarp();
}

enum WalkResult { WALK_OK = 0, WALK_RING_OVERFLOW = 1, WALK_TABLE_OVERFLOW = 2 };

// A row longer than kE, pushed alone (similarity.py:204-216 for one queue entry).  Column indices and
// weights are staged through shared memory with cp.async, kStage stored entries per stage, the next
// stage in flight while the current one is applied.
template <bool HASH>
__device__ int push_hub_row(const PushParams &P, WarpShared &W, Slot &S, int u, const NodeInfo iu, double row_w,
                            double eps, bool first, unsigned &head, unsigned &tail, int lane, unsigned lt)
{
    unsigned long long *ws = W.stat[1];
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const unsigned len = iu.len;
    // state of u
    double su_s, su_r, din_u = iu.d_in;
    unsigned tu = 0;
    if (HASH) {
        if (S.nt + (int)min(len, (unsigned)P.n) + 1 > table_limit(S.lg)) {
            if (!table_grow(S, S.nt + (int)min(len, (unsigned)P.n) + 1, lane)) return WALK_TABLE_OVERFLOW;
        }
        TableEntry e;
        e.s = e.r = 0.0;
        if (lane == 0) table_find(S.T, S.lg, u, e, tu);   // u was enqueued, so it is in the table
        su_s = __shfl_sync(kFull, e.s, 0);
        su_r = __shfl_sync(kFull, e.r, 0);
        tu = __shfl_sync(kFull, tu, 0);
    } else {
        const double2 v = ld_state(&S.sr[u]);
        su_s = v.x;
        su_r = v.y;
    }
    head += 1;
    if (!(first || __ddiv_rn(su_r, din_u) >= eps)) return WALK_OK;  // similarity.py:204
    const double c = __dmul_rn(P.one_minus_rho, su_r);               // push.py:57
    if (lane == 0) {                                                 // push.py:60
        if (HASH) st_entry(S.T + tu, su_s, 0.0, din_u, u);
        else st_state(&S.sr[u], make_double2(su_s, 0.0));
        ws[WS_PUSHES] += 1;
        ws[WS_EDGES] += len;
    }
    __syncwarp();

    // staging buffers over the node cache, which is empty between batches (its keys are not touched)
    int32_t *sidx = reinterpret_cast<int32_t *>(W.cdin);    // [2][kStage] ints  = the 2 KB of cdin
    double *swgt = reinterpret_cast<double *>(W.cval);      // [2][kStage] doubles = the 4 KB of cval
    const int32_t *gidx = P.indices + iu.begin;
    const double *gw = P.w + iu.begin;
    const bool uni = P.uniform_rows != 0;
    auto issue = [&](unsigned base, int buf) {
        for (int k = lane; k < kStage; k += 32) {
            const unsigned j = base + k;
            if (j < len) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&sidx[buf * kStage + k]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gidx + j) : "memory");
                if (!uni) {
                    const unsigned sb = (unsigned)__cvta_generic_to_shared(&swgt[buf * kStage + k]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb), "l"(gw + j) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int result = WALK_OK;
    issue(0, 0);
    int buf = 0;
    for (unsigned base = 0; base < len && result == WALK_OK; base += kStage, buf ^= 1) {
        if (base + kStage < len) {
            issue(base + kStage, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const unsigned stage_n = min((unsigned)kStage, len - base);
        for (unsigned b2 = 0; b2 < stage_n; b2 += 32 * kU) {
            int v[kU];
            double p[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                const unsigned j = b2 + k * 32 + lane;
                v[k] = -1;
                if (j < stage_n) {
                    v[k] = sidx[buf * kStage + j];
                    p[k] = __dmul_rn(c, uni ? row_w : swgt[buf * kStage + j]);
                }
            }
            double os[kU], orr[kU], dv[kU];
            unsigned tk[kU];
            unsigned f_new = 0, f_enq = 0;
            if (HASH) {
                TableEntry e0[kU];
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (v[k] >= 0) e0[k] = ld_entry(S.T + table_hash(v[k], S.lg));
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (v[k] >= 0) {
                        TableEntry e;
                        bool is_new;
                        tk[k] = table_find_or_insert(S.T, S.lg, v[k], e0[k], e, is_new);
                        if (is_new) {
                            os[k] = orr[k] = 0.0;
                            dv[k] = ld_info_din(&P.info[v[k]]);
                            f_new |= 1u << k;
                        } else {
                            os[k] = e.s;
                            orr[k] = e.r;
                            dv[k] = e.d_in;
                        }
                    }
            } else {
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (v[k] >= 0) {
                        const double2 o = ld_state(&S.sr[v[k]]);
                        os[k] = o.x;
                        orr[k] = o.y;
                        dv[k] = ld_info_din(&P.info[v[k]]);
                    }
            }
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (v[k] >= 0) {
                    const double ns = __dadd_rn(os[k], p[k]);   // push.py:63
                    const double nr = __dadd_rn(orr[k], p[k]);  // push.py:64
                    if (HASH) st_entry(S.T + tk[k], ns, nr, dv[k], v[k]);
                    else {
                        st_state(&S.sr[v[k]], make_double2(ns, nr));
                        if ((os[k] == 0.0 && orr[k] == 0.0) && (ns != 0.0 || nr != 0.0)) f_new |= 1u << k;
                    }
                    if (__ddiv_rn(nr, dv[k]) >= eps) f_enq |= 1u << k;  // similarity.py:214
                }
            // ordered appends: stored-entry order = chunk order, then lane order
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                if (b2 + k * 32 >= stage_n) break;  // warp-uniform
                const bool is_new = (f_new >> k) & 1u;
                const unsigned m_new = __ballot_sync(kFull, is_new);
                if (!HASH && is_new) S.touched[S.nt + __popc(m_new & lt)] = v[k];
                S.nt += __popc(m_new);
            }
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                if (b2 + k * 32 >= stage_n) break;  // warp-uniform
                const bool enq = (f_enq >> k) & 1u;
                const unsigned m_enq = __ballot_sync(kFull, enq);
                const unsigned cnt = __popc(m_enq);
                if (cnt) {
                    if (tail - head + cnt > (unsigned)P.queue_cap) { result = WALK_RING_OVERFLOW; break; }
                    if (enq) S.queue[(tail + __popc(m_enq & lt)) & qmask] = v[k];
                    tail += cnt;
                    if (lane == 0) ws[WS_ENQ] += cnt;
                }
            }
            if (result != WALK_OK) break;
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (lane == 0 && tail - head > ws[WS_MAXQ]) ws[WS_MAXQ] = tail - head;
    __syncwarp();
    return result;
}

template <bool HASH>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 4)
k_walk_batched(const PushParams P)
{
    __shared__ WarpShared wsm[kWarpsPerCta];
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * kWarpsPerCta + wib;
    if (slot >= P.n_slots) return;
    WarpShared &W = wsm[wib];
    unsigned long long *wtot = W.stat[0];
    unsigned long long *ws = W.stat[1];
    for (int i = lane; i < kCacheSlots; i += 32) W.ckey[i] = -1;
    if (lane < WS_COUNT) wtot[lane] = 0ull;
    __syncwarp();
    if (lane == 0) {
        unsigned long long t_begin;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        wtot[WS_T_BEGIN] = t_begin;
    }

    Slot S;
    S.queue = P.queue + slot * P.queue_cap;
    S.touched = P.touched + slot * P.touched_stride;
    S.sr = nullptr;
    S.half[0] = S.half[1] = S.T = nullptr;
    S.clean[0] = S.clean[1] = 0;
    S.cur = 0;
    S.lg = kInitLg;
    S.cap_max = P.tbl_cap_max;
    S.nt = 0;
    if (HASH) {
        S.half[0] = P.tbl + slot * 2 * P.tbl_cap_max;
        S.half[1] = S.half[0] + P.tbl_cap_max;
        S.clean[0] = P.tbl_clean[slot * 2 + 0];
        S.clean[1] = P.tbl_clean[slot * 2 + 1];
    } else {
        S.sr = P.sr + slot * P.n;
    }
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const bool uni = P.uniform_rows != 0;

    for (;;) {
        unsigned long long wk = 0;
        if (lane == 0) wk = atomicAdd(&P.counters[PC_WORK_CURSOR], 1ull);
        wk = __shfl_sync(kFull, wk, 0);
        if ((int64_t)wk >= P.n_work) break;
        const int pos = P.work_ids ? P.work_ids[wk] : (int)wk;
        const int seed = P.work_seed[pos];
        const double eps = P.work_eps[pos];
        const NodeInfo si = P.info[seed];
        if (lane < WS_COUNT) ws[lane] = 0ull;

        // ---- initial state: s[seed] = r[seed] = 1 (similarity.py:176-177), queue = [seed] ----
        if (HASH) {
            int lg0 = kInitLg;
            while (table_limit(lg0) < (int)min(si.len, 4096u) + 2 + kE && ((int64_t)2 << lg0) <= S.cap_max) ++lg0;
            S.lg = lg0;
            S.cur = 0;
            S.T = S.half[0];
            table_make_clean(S, 0, lg0, lane);
            if (lane == 0) st_entry(S.T + table_hash(seed, lg0), 1.0, 1.0, si.d_in, seed);
        } else if (lane == 0) {
            st_state(&S.sr[seed], make_double2(1.0, 1.0));
            S.touched[0] = seed;
        }
        S.nt = 1;
        if (lane == 0) S.queue[0] = seed;
        __syncwarp();

        unsigned head = 0, tail = 1;
        bool first = true;   // "Do one push for free", similarity.py:183-196
        int result = WALK_OK;

        while (head != tail && result == WALK_OK) {
            // ---- the batch: consecutive queue entries whose rows fit kE edge slots ----
            const unsigned avail = tail - head;
            const int B0 = avail < 32u ? (int)avail : 32;
            int u = -1;
            NodeInfo iu;
            iu.d_in = 1.0;
            iu.begin = 0;
            iu.len = 0;
            double rw = 0.0;
            if (lane < B0) {
                u = S.queue[(head + lane) & qmask];
                iu = ld_info(&P.info[u]);
                if (uni) rw = P.row_w[u];
            }
            unsigned incl = iu.len > (unsigned)kE ? (unsigned)kE + 1u : iu.len;   // saturate: sums stay small
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const bool fit = lane < B0 && incl <= (unsigned)kE;
            const int Bp = __popc(__ballot_sync(kFull, fit));   // rows are taken from the front: fit is a prefix
            if (Bp == 0) {
                const int u0 = __shfl_sync(kFull, u, 0);
                NodeInfo i0;
                i0.d_in = __shfl_sync(kFull, iu.d_in, 0);
                i0.begin = __shfl_sync(kFull, iu.begin, 0);
                i0.len = __shfl_sync(kFull, iu.len, 0);
                const double rw0 = __shfl_sync(kFull, rw, 0);
                result = push_hub_row<HASH>(P, W, S, u0, i0, rw0, eps, first, head, tail, lane, lt);
                first = false;
                continue;
            }
            const int total = (int)__shfl_sync(kFull, incl, Bp - 1);
            if (HASH && S.nt + total + 1 > table_limit(S.lg)) {
                if (!table_grow(S, S.nt + total + 1, lane)) { result = WALK_TABLE_OVERFLOW; break; }
            }
            if (lane < Bp) {
                W.eu[lane] = u;
                W.ebeg[lane] = iu.begin;
                W.epre[lane + 1] = (int)incl;
                W.erw[lane] = rw;
            }
            if (lane == 0) W.epre[0] = 0;
            __syncwarp();

            // ---- gather: neighbour ids + weights of all edge slots, then every distinct node's state ----
            int vk[kU], jk[kU], qk[kU];
            double wgt[kU];
            unsigned tk[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                const int e = k * 32 + lane;
                vk[k] = -1;
                jk[k] = -1;
                qk[k] = 0;
                tk[k] = 0;
                wgt[k] = 0.0;
                if (e < total) {
                    int lo = 0, hi = Bp - 1;   // the entry j with epre[j] <= e < epre[j+1]
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (W.epre[mid + 1] > e) hi = mid;
                        else lo = mid + 1;
                    }
                    const unsigned at = W.ebeg[lo] + (unsigned)(e - W.epre[lo]);
                    jk[k] = lo;
                    vk[k] = ld_index(P.indices + at);
                    wgt[k] = uni ? W.erw[lo] : ld_weight(P.w + at);
                }
            }
            unsigned own = 0, f_new = 0;   // bit k: edge slot k, bit kU: the lane's batch entry
            int qe = 0;
            unsigned te = 0;
            if (lane < Bp) {
                bool o;
                qe = cache_insert(W.ckey, u, o);
                W.eq[lane] = qe;
                if (o) own |= 1u << kU;
            }
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (vk[k] >= 0) {
                    bool o;
                    qk[k] = cache_insert(W.ckey, vk[k], o);
                    if (o) own |= 1u << k;
                }
            if (HASH) {
                TableEntry e0[kU + 1];
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if ((own >> k) & 1u) e0[k] = ld_entry(S.T + table_hash(vk[k], S.lg));
                if ((own >> kU) & 1u) e0[kU] = ld_entry(S.T + table_hash(u, S.lg));
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if ((own >> k) & 1u) {
                        TableEntry e;
                        bool is_new;
                        tk[k] = table_find_or_insert(S.T, S.lg, vk[k], e0[k], e, is_new);
                        if (is_new) {
                            W.cval[qk[k]] = make_double2(0.0, 0.0);
                            W.cdin[qk[k]] = ld_info_din(&P.info[vk[k]]);
                            f_new |= 1u << k;
                        } else {
                            W.cval[qk[k]] = make_double2(e.s, e.r);
                            W.cdin[qk[k]] = e.d_in;
                        }
                    }
                if ((own >> kU) & 1u) {
                    TableEntry e;
                    bool is_new;
                    te = table_find_or_insert(S.T, S.lg, u, e0[kU], e, is_new);   // always found: u was enqueued
                    W.cval[qe] = is_new ? make_double2(0.0, 0.0) : make_double2(e.s, e.r);
                    W.cdin[qe] = iu.d_in;
                    if (is_new) f_new |= 1u << kU;
                }
            } else {
                double2 o[kU + 1];
                double dv[kU];
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if ((own >> k) & 1u) {
                        o[k] = ld_state(&S.sr[vk[k]]);
                        dv[k] = ld_info_din(&P.info[vk[k]]);
                    }
                if ((own >> kU) & 1u) o[kU] = ld_state(&S.sr[u]);
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if ((own >> k) & 1u) {
                        W.cval[qk[k]] = o[k];
                        W.cdin[qk[k]] = dv[k];
                        if (o[k].x == 0.0 && o[k].y == 0.0) f_new |= 1u << k;   // untouched so far
                    }
                if ((own >> kU) & 1u) {
                    W.cval[qe] = o[kU];
                    W.cdin[qe] = iu.d_in;
                }
            }
            __syncwarp();

            // ---- apply the pushes to the cache strictly in queue order (similarity.py:199-216) ----
            for (int j = 0; j < Bp; ++j) {
                const int qu = W.eq[j];
                const double2 su = W.cval[qu];
                const double din_u = W.cdin[qu];
                const bool pass = first || __ddiv_rn(su.y, din_u) >= eps;   // similarity.py:204
                first = false;
                __syncwarp();
                if (!pass) continue;   // warp-uniform
                const double c = __dmul_rn(P.one_minus_rho, su.y);          // push.py:57
                if (lane == 0) W.cval[qu] = make_double2(su.x, 0.0);         // push.py:60
                __syncwarp();
                unsigned f_enq = 0;
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (jk[k] == j) {   // rows hold distinct columns: no two lanes share a cache slot here
                        const double2 o = W.cval[qk[k]];
                        const double p = __dmul_rn(c, wgt[k]);
                        const double2 nw = make_double2(__dadd_rn(o.x, p), __dadd_rn(o.y, p));   // push.py:63-64
                        W.cval[qk[k]] = nw;
                        if (__ddiv_rn(nw.y, W.cdin[qk[k]]) >= eps) f_enq |= 1u << k;              // similarity.py:214
                    }
                const unsigned popped = head + (unsigned)j + 1u;
#pragma unroll
                for (int k = 0; k < kU; ++k) {
                    const unsigned m_enq = __ballot_sync(kFull, (f_enq >> k) & 1u);
                    const unsigned cnt = __popc(m_enq);
                    if (cnt) {
                        // entries of this batch are already in shared memory: their ring cells are free
                        if (tail - (head + (unsigned)Bp) + cnt > (unsigned)P.queue_cap) { result = WALK_RING_OVERFLOW; break; }
                        if ((f_enq >> k) & 1u) S.queue[(tail + __popc(m_enq & lt)) & qmask] = vk[k];
                        tail += cnt;
                        if (lane == 0) ws[WS_ENQ] += cnt;
                    }
                }
                if (result != WALK_OK) break;
                if (lane == 0) {
                    ws[WS_PUSHES] += 1;
                    ws[WS_EDGES] += (unsigned)(W.epre[j + 1] - W.epre[j]);
                    if (tail - popped > ws[WS_MAXQ]) ws[WS_MAXQ] = tail - popped;
                }
                __syncwarp();
            }
            if (result != WALK_OK) {
                for (int i = lane; i < kCacheSlots; i += 32) W.ckey[i] = -1;
                __syncwarp();
                break;
            }

            // ---- write the cache back, one store per distinct node ----
            if (HASH) {
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if ((own >> k) & 1u) {
                        const double2 nv = W.cval[qk[k]];
                        st_entry(S.T + tk[k], nv.x, nv.y, W.cdin[qk[k]], vk[k]);
                        W.ckey[qk[k]] = -1;
                    }
                if ((own >> kU) & 1u) {
                    const double2 nv = W.cval[qe];
                    st_entry(S.T + te, nv.x, nv.y, iu.d_in, u);
                    W.ckey[qe] = -1;
                }
#pragma unroll
                for (int k = 0; k <= kU; ++k) S.nt += __popc(__ballot_sync(kFull, (f_new >> k) & 1u));
            } else {
#pragma unroll
                for (int k = 0; k < kU; ++k) {
                    bool is_new = false;
                    if ((own >> k) & 1u) {
                        const double2 nv = W.cval[qk[k]];
                        st_state(&S.sr[vk[k]], nv);
                        W.ckey[qk[k]] = -1;
                        is_new = ((f_new >> k) & 1u) && (nv.x != 0.0 || nv.y != 0.0);
                    }
                    const unsigned m_new = __ballot_sync(kFull, is_new);
                    if (is_new) S.touched[S.nt + __popc(m_new & lt)] = vk[k];
                    S.nt += __popc(m_new);
                }
                if ((own >> kU) & 1u) {
                    st_state(&S.sr[u], W.cval[qe]);
                    W.ckey[qe] = -1;
                }
            }
            __syncwarp();
            head += (unsigned)Bp;
        }

        if (P.debug_keep) {   // operator seam: dense s and r of this one walk
            if (HASH) {
                const int C = 1 << S.lg;
                for (int i = lane; i < C; i += 32) {
                    const TableEntry e = ld_entry(S.T + i);
                    if (e.key != kEmptyKey) {
                        if (result == WALK_OK) {
                            P.dbg_s[e.key] = e.s;
                            P.dbg_r[e.key] = e.r;
                        }
                        S.T[i].key = kEmptyKey;
                    }
                }
            }
            if (lane == 0) {
                P.counters[PC_PUSHES] = ws[WS_PUSHES];
                P.counters[PC_TOUCHED] = (unsigned long long)S.nt;
                P.counters[PC_OVERFLOW_SEEDS] = result == WALK_OK ? 0ull : 1ull;
                if (result == WALK_TABLE_OVERFLOW) P.counters[PC_TOVERFLOW] = 1ull;
            }
            break;
        }

        if (result != WALK_OK) {
            // undo and hand the seed to the retry pass
            if (HASH) table_clear(S, lane);
            else {
                __syncwarp();
                for (int i = lane; i < S.nt; i += 32) st_state(&S.sr[S.touched[i]], make_double2(0.0, 0.0));
            }
            if (lane == 0) {
                const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                P.retry_list[r] = pos;
                atomicAdd(&P.counters[result == WALK_RING_OVERFLOW ? PC_QOVERFLOW : PC_TOVERFLOW], 1ull);
                P.seg_count[pos] = -1;
            }
            __syncwarp();
            continue;
        }

        // ---------------- K4: threshold + membership (arcte.py:352-376) ----------------
        const int base_size = (int)si.len + 1;   // np.append(adjacent_nodes[n], n), arcte.py:358
        double q;
        int m = 0, support = 0;
        if (HASH) {
            // tau = min over N(seed) + seed of s/d_in (arcte.py:355-360); every one of them is in the table
            {
                TableEntry e;
                unsigned at;
                e.s = 0.0;
                table_find(S.T, S.lg, seed, e, at);
                q = __ddiv_rn(e.s, si.d_in);
            }
            for (unsigned j0 = 0; j0 < si.len; j0 += 32) {
                const unsigned j = j0 + lane;
                if (j < si.len) {
                    const int v = P.indices[si.begin + j];
                    TableEntry e;
                    unsigned at;
                    if (table_find(S.T, S.lg, v, e, at)) q = fmin(q, __ddiv_rn(e.s, e.d_in));
                    else q = fmin(q, __ddiv_rn(0.0, ld_info_din(&P.info[v])));
                }
            }
            const double tau = warp_min(q);
            // one linear scan: support count, members (ties included, arcte.py:363-367), table reset
            const int C = 1 << S.lg;
            for (int i0 = 0; i0 < C; i0 += 64) {
                TableEntry e[2];
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) e[k2] = ld_entry(S.T + i0 + k2 * 32 + lane);   // C >= 1024: in range
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const bool occ = e[k2].key != kEmptyKey;
                    const bool in_sup = occ && e[k2].s != 0.0;
                    const bool pass = in_sup && (__ddiv_rn(e[k2].s, e[k2].d_in) >= tau);
                    if (occ) S.T[i0 + k2 * 32 + lane].key = kEmptyKey;
                    support += __popc(__ballot_sync(kFull, in_sup));
                    const unsigned mp = __ballot_sync(kFull, pass);
                    if (pass) S.touched[m + __popc(mp & lt)] = e[k2].key;
                    m += __popc(mp);
                }
            }
        } else {
            q = __ddiv_rn(ld_state(&S.sr[seed]).x, si.d_in);
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
                        o[k2] = ld_state(&S.sr[v[k2]]);
                        d[k2] = ld_info_din(&P.info[v[k2]]);
                    }
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2)
                    if (v[k2] >= 0) q = fmin(q, __ddiv_rn(o[k2].x, d[k2]));   // arcte.py:355-356
            }
            const double tau = warp_min(q);   // arcte.py:359-360
            for (int i0 = 0; i0 < S.nt; i0 += 64) {
                int x[2];
                double sx[2], dx[2];
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const int i = i0 + k2 * 32 + lane;
                    x[k2] = i < S.nt ? S.touched[i] : -1;
                }
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2)
                    if (x[k2] >= 0) {
                        sx[k2] = ld_state(&S.sr[x[k2]]).x;
                        dx[k2] = ld_info_din(&P.info[x[k2]]);
                    }
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    if (i0 + k2 * 32 >= S.nt) break;   // warp-uniform
                    bool in_sup = false, pass = false;
                    if (x[k2] >= 0) {
                        st_state(&S.sr[x[k2]], make_double2(0.0, 0.0));
                        in_sup = sx[k2] != 0.0;
                        pass = in_sup && (__ddiv_rn(sx[k2], dx[k2]) >= tau);
                    }
                    support += __popc(__ballot_sync(kFull, in_sup));
                    const unsigned mp = __ballot_sync(kFull, pass);
                    if (pass) S.touched[m + __popc(mp & lt)] = x[k2];
                    m += __popc(mp);
                }
            }
        }
        __syncwarp();
        const bool emit = m > base_size;   // arcte.py:370
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
                for (int i = lane; i < m; i += 32) P.members[off + i] = S.touched[i];   // arcte.py:372-376
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
            wtot[WS_TOUCHED] += S.nt;
            wtot[WS_SEEDDEG] += si.len;
            if (emit) {
                wtot[WS_MEMBERS] += m;
                wtot[WS_EMITTED] += 1;
            }
        }
        __syncwarp();
    }

    if (HASH && lane == 0) {
        P.tbl_clean[slot * 2 + 0] = S.clean[0];
        P.tbl_clean[slot * 2 + 1] = S.clean[1];
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

// ---- host side ---------------------------------------------------------------------------------------
static int64_t pow2_ge(int64_t v)
{
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static int default_warps(const arcte_cuda_ctx *c)
{
    int wps = c->warps_per_sm > 0 ? c->warps_per_sm : 16;
    if (wps > 24) wps = 24;   // 7.8 KB of shared memory per warp
    wps = ((wps + kWarpsPerCta - 1) / kWarpsPerCta) * kWarpsPerCta;
    return wps;
}

int batched_plan(arcte_cuda_ctx *c, int engine, int64_t n_work, int64_t *n_slots, int64_t *queue_cap)
{
    int64_t want = (int64_t)c->sm_count * default_warps(c);
    const int64_t work_r = ((n_work + kWarpsPerCta - 1) / kWarpsPerCta) * kWarpsPerCta;
    if (want > work_r) want = work_r;
    if (want < kWarpsPerCta) want = kWarpsPerCta;
    int64_t qcap = c->queue_cap_cfg > 0 ? c->queue_cap_cfg : (c->n < 65536 ? c->n : 65536);
    if (c->queue_cap_cfg <= 0 && qcap < 8192) qcap = 8192;
    if (qcap < 64) qcap = 64;
    qcap = pow2_ge(qcap);
    size_t free_b = 0, total_b = 0;
    ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const int pct = c->mem_percent > 0 ? c->mem_percent : 60;
    if (engine == ARCTE_ENGINE_BATCHED_HASH) {
        BatchedPool &bp = c->bpool;
        free_b += bp.tbl.bytes + bp.stage.bytes + bp.queue.bytes;
        // table region: two halves of cap entries; cap large enough for every node at 62.5 % load, bounded by memory
        int64_t cap = c->tbl_cap_cfg > 0 ? pow2_ge(c->tbl_cap_cfg) : pow2_ge(2 * c->n);
        if (cap < (1 << kInitLg)) cap = 1 << kInitLg;
        if (cap > ((int64_t)1 << 26)) cap = (int64_t)1 << 26;
        const double budget = (double)free_b * pct / 100.0;
        while (cap > (1 << kInitLg) && (double)want * (68.0 * (double)cap + 4.0 * (double)qcap) > budget) cap >>= 1;
        if ((double)want * (68.0 * (double)cap + 4.0 * (double)qcap) > budget) {
            want = (int64_t)(budget / (68.0 * (double)cap + 4.0 * (double)qcap));
            want = (want / kWarpsPerCta) * kWarpsPerCta;
            if (want < kWarpsPerCta) { set_error("not enough device memory for the walk tables"); return ARCTE_E_NOMEM; }
        }
        bp.plan_cap = cap;
    } else {
        free_b += c->slots.sr.bytes + c->slots.touched.bytes + c->slots.queue.bytes;
        const double budget = (double)free_b * pct / 100.0;
        const double per_slot = 20.0 * (double)c->n + 4.0 * (double)qcap;
        int64_t fit = (int64_t)(budget / per_slot);
        fit = (fit / kWarpsPerCta) * kWarpsPerCta;
        if (fit < kWarpsPerCta) { set_error("not enough device memory for the walk states of this graph"); return ARCTE_E_NOMEM; }
        if (want > fit) want = fit;
    }
    *n_slots = want;
    *queue_cap = qcap;
    return ARCTE_OK;
}

int batched_ensure(arcte_cuda_ctx *c, int engine, int64_t n_slots, int64_t qcap)
{
    if (!c->row_w_valid) { set_error("batched engine: graph not prepared"); return ARCTE_E_ARG; }
    if (engine == ARCTE_ENGINE_BATCHED_HASH) {
        BatchedPool &bp = c->bpool;
        const int64_t cap = bp.plan_cap;
        if (!(bp.n_slots >= n_slots && bp.cap == cap)) {
            dev_free(bp.tbl);
            dev_free(bp.stage);
            dev_free(bp.clean);
            bp.n_slots = 0;
            ARCTE_TRY(dev_reserve(bp.tbl, sizeof(TableEntry) * 2 * (size_t)n_slots * (size_t)cap));
            ARCTE_TRY(dev_reserve(bp.stage, sizeof(int32_t) * (size_t)n_slots * (size_t)cap));
            ARCTE_TRY(dev_reserve(bp.clean, sizeof(int32_t) * 2 * (size_t)n_slots));
            // nothing of the tables is initialised here: a warp makes the part it is about to use all-EMPTY
            // the first time it needs it (tbl_clean), so an extraction never pays for the whole region
            ARCTE_CUDA_TRY(cudaMemsetAsync(bp.clean.p, 0, sizeof(int32_t) * 2 * (size_t)n_slots, c->stream));
            bp.n_slots = n_slots;
            bp.cap = cap;
        }
        if (bp.queue_cap != qcap || bp.queue_slots < n_slots) {
            dev_free(bp.queue);
            ARCTE_TRY(dev_reserve(bp.queue, sizeof(int32_t) * (size_t)bp.n_slots * (size_t)qcap));
            bp.queue_cap = qcap;
            bp.queue_slots = bp.n_slots;
        }
        return ARCTE_OK;
    }
    SlotPool &sp = c->slots;
    if (!(sp.n == c->n && sp.n_slots >= n_slots)) {
        dev_free(sp.sr);
        dev_free(sp.touched);
        dev_free(sp.queue);
        sp.n_slots = sp.queue_slots = 0;
        sp.queue_cap = 0;
        ARCTE_TRY(dev_reserve(sp.sr, sizeof(double2) * (size_t)n_slots * (size_t)c->n));
        ARCTE_TRY(dev_reserve(sp.touched, sizeof(int32_t) * (size_t)n_slots * (size_t)c->n));
        ARCTE_CUDA_TRY(cudaMemsetAsync(sp.sr.p, 0, sizeof(double2) * (size_t)n_slots * (size_t)c->n, c->stream));
        sp.n = c->n;
        sp.n_slots = n_slots;
    }
    if (sp.queue_cap != qcap || sp.queue_slots < n_slots) {
        dev_free(sp.queue);
        ARCTE_TRY(dev_reserve(sp.queue, sizeof(int32_t) * (size_t)sp.n_slots * (size_t)qcap));
        sp.queue_cap = qcap;
        sp.queue_slots = sp.n_slots;
    }
    return ARCTE_OK;
}

void batched_fill_params(arcte_cuda_ctx *c, int engine, PushParams &P)
{
    P.uniform_rows = c->uniform_rows ? 1 : 0;
    P.row_w = c->row_w.as<double>();
    if (engine == ARCTE_ENGINE_BATCHED_HASH) {
        const BatchedPool &bp = c->bpool;
        P.tbl = bp.tbl.as<TableEntry>();
        P.tbl_cap_max = bp.cap;
        P.tbl_clean = bp.clean.as<int32_t>();
        P.touched = bp.stage.as<int32_t>();
        P.touched_stride = bp.cap;
        P.queue = bp.queue.as<int32_t>();
        P.queue_cap = bp.queue_cap;
        P.sr = nullptr;
    } else {
        P.sr = c->slots.sr.as<double2>();
        P.touched = c->slots.touched.as<int32_t>();
        P.touched_stride = c->n;
        P.queue = c->slots.queue.as<int32_t>();
        P.queue_cap = c->slots.queue_cap;
        P.tbl = nullptr;
        P.tbl_cap_max = 0;
        P.tbl_clean = nullptr;
    }
}

int batched_launch(arcte_cuda_ctx *c, int engine, const PushParams &P)
{
    const unsigned grid = (unsigned)((P.n_slots + kWarpsPerCta - 1) / kWarpsPerCta);
    // shared memory wanted per SM: resident warps x sizeof(WarpShared); leave the rest to L1
    const int want_kb = (int)((default_warps(c) * sizeof(WarpShared) + 1023) / 1024) + 8;
    int carve = (want_kb * 100 + 227) / 228;
    if (carve > 100) carve = 100;
    if (engine == ARCTE_ENGINE_BATCHED_HASH) {
        ARCTE_CUDA_TRY(cudaFuncSetAttribute(k_walk_batched<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        k_walk_batched<true><<<grid, 32 * kWarpsPerCta, 0, c->stream>>>(P);
    } else {
        ARCTE_CUDA_TRY(cudaFuncSetAttribute(k_walk_batched<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        k_walk_batched<false><<<grid, 32 * kWarpsPerCta, 0, c->stream>>>(P);
    }
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

}  // namespace arcte
