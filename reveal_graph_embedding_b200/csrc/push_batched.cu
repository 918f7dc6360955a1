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
constexpr int kCacheSlots = 128;     // shared-memory node cache per warp (<= 96 distinct nodes per batch)
constexpr int kWarpsPerCta = 4;
constexpr int kStage = 64;           // stored entries per cp.async stage of a hub row
constexpr int kInitLg = 10;          // a walk's table starts at 1024 entries
constexpr int kGrowLg = 1;           // and doubles on demand
static_assert(kU == 2, "edge-slot masks are 64 bits wide");

// Per-warp shared memory.  The node cache of a batch maps a node to its references: which edge slots add
// to it (cref) and which batch entry it is (centry); `contrib` holds the addend of every edge slot whose
// node has more than one reference.  A hub row is staged over the same bytes (the cache is empty then).
struct BatchShared {
    unsigned long long cref[kCacheSlots];
    uint32_t centry[kCacheSlots];
    double contrib[kE];
};
struct StageShared {
    int32_t idx[2][kStage];
    double w[2][kStage];
    double din[2][kStage];
};
struct WarpShared {
    union {
        BatchShared b;
        StageShared s;
    } u;
    int32_t ckey[kCacheSlots];       // node id or -1
    unsigned long long enqmask;      // edge slots that enqueue, from the multi-reference nodes
    int32_t epre[33];                // exclusive prefix of the batch entries' row lengths
    uint32_t ebeg[32];               // batch entries: row begin
    double erw[32];                  //                row weight (uniform rows)
    unsigned long long stat[WS_COUNT];   // totals of the finished seeds
};
// the hub-row stage overwrites cref, centry (re-zeroed afterwards) and contrib (needs no reset)

__device__ __forceinline__ unsigned long long below64(int x) { return x >= 64 ? ~0ull : (x <= 0 ? 0ull : ((1ull << x) - 1ull)); }
__device__ __forceinline__ unsigned below32(int x) { return x >= 32 ? 0xffffffffu : (x <= 0 ? 0u : ((1u << x) - 1u)); }

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
    unsigned h = ((unsigned)v * 0x9E3779B1u) >> 25;
    for (;;) {
        const int old = atomicCAS(&ckey[h], -1, v);
        if (old == -1) { owner = true; return (int)h; }
        if (old == v) { owner = false; return (int)h; }
        h = (h + 1) & (kCacheSlots - 1);
    }
}

// Walk state of one slot.
struct Slot {
    // direct-mapped: T[v] is node v's entry, valid when its key is `epoch`
    int32_t epoch;
    int32_t *touched;   // direct: touched list (then the members); hash: member staging list
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

// Phase profile (-DARCTE_BATCH_PROFILE, ARCTE_CUDA_PROFILE=1 prints it): clock64 sums of lane 0 per phase.
#ifdef ARCTE_BATCH_PROFILE
#define PROF_DECL long long prof_t = clock64(); long long prof_acc[20] = {0}
#define PROF(i) do { const long long t_ = clock64(); prof_acc[i] += t_ - prof_t; prof_t = t_; } while (0)
#define PROF_ADD(i, v) do { prof_acc[i] += (v); } while (0)
#define PROF_FLUSH() do { if (lane == 0) for (int i_ = 0; i_ < 20; ++i_) if (prof_acc[i_]) atomicAdd(&P.counters[PC_PROF0 + i_], (unsigned long long)prof_acc[i_]); } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_ADD(i, v)
#define PROF_FLUSH()
#endif

enum WalkResult { WALK_OK = 0, WALK_RING_OVERFLOW = 1, WALK_TABLE_OVERFLOW = 2 };

// Counters of the walk in progress (warp-uniform registers; added to the warp's totals when the seed is done).
struct WalkStat {
    unsigned long long pushes, edges, enq;
    unsigned maxq;
};

// Hash engine: the state of node v for a push that is about to add to it.  The probe of the home entry and
// the claim of a free entry (compare-and-swap on the key: lanes of the warp insert DISTINCT nodes at the
// same time) are issued together, so a touch is one memory round trip whether the node is new or not; only
// a collision costs another one.  (The in-degree a new entry needs arrives with the edge, PushParams::edge_din.)
// Direct-mapped engine: entry v IS node v's, valid when tagged with the walk's epoch.
__device__ __forceinline__ void direct_touch(const TableEntry *T, int epoch, int v, double &s, double &r, bool &is_new)
{
    const TableEntry e = ld_entry(T + v);
    is_new = e.key != epoch;
    s = is_new ? 0.0 : e.s;
    r = is_new ? 0.0 : e.r;
}
__device__ __forceinline__ unsigned table_touch(TableEntry *T, int lg, int v, double &s, double &r, bool &is_new)
{
    const unsigned mask = (1u << lg) - 1u;
    unsigned h = table_hash(v, lg);
    TableEntry e = ld_entry(T + h);
    int old = atomicCAS(&T[h].key, kEmptyKey, v);
    for (;;) {
        if (old == kEmptyKey) { s = r = 0.0; is_new = true; return h; }
        if (old == v) {
            if (e.key != v) e = ld_entry(T + h);   // cannot happen (the entry is complete since an earlier batch)
            s = e.s; r = e.r; is_new = false; return h;
        }
        h = (h + 1) & mask;
        e = ld_entry(T + h);
        old = atomicCAS(&T[h].key, kEmptyKey, v);
    }
}

// A row longer than kE, pushed alone (similarity.py:204-216 for one queue entry).  Column indices and
// weights are staged through shared memory with cp.async, kStage stored entries per stage, the next
// stage in flight while the current one is applied.  The caller has popped the entry.
template <bool HASH>
__device__ int push_hub_row(const PushParams &P, WarpShared &W, Slot &S, WalkStat &ws, int u, const NodeInfo iu, double row_w,
                            const Threshold &eps, bool first, unsigned head, unsigned &tail, int lane, unsigned lt)
{
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const unsigned len = iu.len;
    // state of u
    double su_s, su_r;
    const double din_u = iu.d_in;
    unsigned tu = 0;
    if (HASH) {
        TableEntry e;
        e.s = e.r = 0.0;
        if (lane == 0) table_find(S.T, S.lg, u, e, tu);   // u was enqueued, so it is in the table
        su_s = __shfl_sync(kFull, e.s, 0);
        su_r = __shfl_sync(kFull, e.r, 0);
        tu = __shfl_sync(kFull, tu, 0);
    } else {
        const TableEntry e = ld_entry(S.T + u);   // queue entries carry the walk's epoch
        su_s = e.s;
        su_r = e.r;
    }
    if (!(first || quot_ge(su_r, din_u, eps))) return WALK_OK;  // similarity.py:204
    if (HASH) {
        const int need = S.nt + (int)min(len, (unsigned)P.n) + 1;
        if (need > table_limit(S.lg)) {
            if (!table_grow(S, need, lane)) return WALK_TABLE_OVERFLOW;
            TableEntry e;
            if (lane == 0) table_find(S.T, S.lg, u, e, tu);   // the table moved
            tu = __shfl_sync(kFull, tu, 0);
        }
    }
    const double c = __dmul_rn(P.one_minus_rho, su_r);               // push.py:57
    if (lane == 0) {                                                 // push.py:60
        if (HASH) st_entry(S.T + tu, su_s, 0.0, din_u, u);
        else st_entry(S.T + u, su_s, 0.0, din_u, S.epoch);
    }
    ws.pushes += 1;
    ws.edges += len;
    __syncwarp();

    // staging buffers over the reference masks of the node cache, which is empty between batches
    int32_t *sidx = &W.u.s.idx[0][0];
    double *swgt = &W.u.s.w[0][0];
    double *sdin = &W.u.s.din[0][0];
    const int32_t *gidx = P.indices + iu.begin;
    const double *gw = P.w + iu.begin;
    const double *gdin = P.edge_din + iu.begin;
    const bool uni = P.uniform_rows != 0;
    auto issue = [&](unsigned base, int buf) {
        for (int k = lane; k < kStage; k += 32) {
            const unsigned j = base + k;
            if (j < len) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&sidx[buf * kStage + k]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gidx + j) : "memory");
                const unsigned sc = (unsigned)__cvta_generic_to_shared(&sdin[buf * kStage + k]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sc), "l"(gdin + j) : "memory");
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
            double p[kU], dv[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                const unsigned j = b2 + k * 32 + lane;
                v[k] = -1;
                dv[k] = 1.0;
                if (j < stage_n) {
                    v[k] = sidx[buf * kStage + j];
                    dv[k] = sdin[buf * kStage + j];
                    p[k] = __dmul_rn(c, uni ? row_w : swgt[buf * kStage + j]);
                }
            }
            double os[kU], orr[kU];
            unsigned tk[kU];
            unsigned f_new = 0, f_enq = 0;
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (v[k] >= 0) {
                    bool is_new;
                    if (HASH) tk[k] = table_touch(S.T, S.lg, v[k], os[k], orr[k], is_new);
                    else direct_touch(S.T, S.epoch, v[k], os[k], orr[k], is_new);
                    if (is_new) f_new |= 1u << k;
                }
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (v[k] >= 0) {
                    const double ns = __dadd_rn(os[k], p[k]);   // push.py:63
                    const double nr = __dadd_rn(orr[k], p[k]);  // push.py:64
                    if (HASH) st_entry(S.T + tk[k], ns, nr, dv[k], v[k]);
                    else st_entry(S.T + v[k], ns, nr, dv[k], S.epoch);
                    if (quot_ge(nr, dv[k], eps)) f_enq |= 1u << k;  // similarity.py:214
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
                    ws.enq += cnt;
                }
            }
            if (result != WALK_OK) break;
        }
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    for (int i = lane; i < kCacheSlots; i += 32) {   // the reference masks again, all clear
        W.u.b.cref[i] = 0ull;
        W.u.b.centry[i] = 0u;
    }
    __syncwarp();
    if (tail - head > ws.maxq) ws.maxq = tail - head;
    return result;
}

#ifndef ARCTE_BATCH_MIN_CTAS
#define ARCTE_BATCH_MIN_CTAS 4
#endif

template <bool HASH>
__global__ void __launch_bounds__(32 * kWarpsPerCta, ARCTE_BATCH_MIN_CTAS)
k_walk_batched(const PushParams P)
{
    __shared__ WarpShared wsm[kWarpsPerCta];
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    const int wib = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * kWarpsPerCta + wib;
    if (slot >= P.n_slots) return;
    WarpShared &W = wsm[wib];
    unsigned long long *wtot = W.stat;
    for (int i = lane; i < kCacheSlots; i += 32) {
        W.ckey[i] = -1;
        W.u.b.cref[i] = 0ull;
        W.u.b.centry[i] = 0u;
    }
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
    S.epoch = 0;
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
        S.T = P.tbl + slot * P.tbl_cap_max;     // one entry per node
        S.epoch = P.tbl_clean[slot * 2 + 0];    // last epoch this slot used (0 = none: the pool starts zeroed)
    }
    const unsigned qmask = (unsigned)P.queue_cap - 1u;
    const bool uni = P.uniform_rows != 0;
    PROF_DECL;

    for (;;) {
        unsigned long long wk = 0;
        if (lane == 0) wk = atomicAdd(&P.counters[PC_WORK_CURSOR], 1ull);
        wk = __shfl_sync(kFull, wk, 0);
        if ((int64_t)wk >= P.n_work) break;
        const int pos = P.work_ids ? P.work_ids[wk] : (int)wk;
        const int seed = P.work_seed[pos];
        const Threshold eps = make_threshold(P.work_eps[pos]);
        const NodeInfo si = P.info[seed];
        WalkStat ws;
        ws.pushes = ws.edges = ws.enq = 0ull;
        ws.maxq = 0u;

        // ---- initial state: s[seed] = r[seed] = 1 (similarity.py:176-177), queue = [seed] ----
        if (HASH) {
            int lg0 = kInitLg;
            while (table_limit(lg0) < (int)min(si.len, 4096u) + 2 + kE && ((int64_t)2 << lg0) <= S.cap_max) ++lg0;
            S.lg = lg0;
            S.cur = 0;
            S.T = S.half[0];
            table_make_clean(S, 0, lg0, lane);
            if (lane == 0) st_entry(S.T + table_hash(seed, lg0), 1.0, 1.0, si.d_in, seed);
        } else {
            S.epoch += 1;   // every entry of the previous walk (finished or aborted) is stale from here on
            if (lane == 0) {
                st_entry(S.T + seed, 1.0, 1.0, si.d_in, S.epoch);
                S.touched[0] = seed;
            }
        }
        S.nt = 1;
        if (lane == 0) S.queue[0] = seed;
        __syncwarp();

        unsigned head = 0, tail = 1;
        bool first = true;   // "Do one push for free", similarity.py:183-196
        int result = WALK_OK;
        // Queue entries already in registers: lane i holds the entry at queue position head + i (node, node
        // record, row weight) for i < n_main.  Entries a batch does not consume stay (shifted down), and the
        // ones behind them are fetched while the batch is in flight: every node record is read once.
        int eu = -1;
        double ed = 1.0, erw = 0.0;
        unsigned eb = 0, el = 0;
        int n_main = 0;
        PROF(8);

        while (head != tail && result == WALK_OK) {
            const unsigned avail = tail - head;
            const int want = avail < 32u ? (int)avail : 32;
            const bool first_in = first;
            // ---- (A) queue entries not yet in registers: node ids now, node records a little later ----
            const bool incoming = lane >= n_main && lane < want;
            const int n_main0 = n_main;
            int nu = -1;
            double nd = 1.0, nrw = 0.0;
            unsigned nb = 0, nl = 0;
            if (incoming) nu = S.queue[(head + lane) & qmask];
            if (n_main0 == 0) {   // nothing pre-fetched (start of a walk, or the queue ran dry): wait for them
                if (incoming) {
                    const NodeInfo t = ld_info(&P.info[nu]);
                    eu = nu; ed = t.d_in; eb = t.begin; el = t.len;
                    if (uni) erw = P.row_w[nu];
                }
                n_main = want;
            }
            // ---- (B) the batch: leading entries whose rows fit kE edge slots; a row longer than kE takes no
            //      slot here: only its pop check is done in the batch, the push itself (if any) on its own ----
            const bool have = lane < n_main;
            const bool hub = have && el > (unsigned)kE;
            unsigned incl = have ? (hub ? 0u : el) : (unsigned)kE + 1u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const int Bp = __popc(__ballot_sync(kFull, have && incl <= (unsigned)kE));   // >= 1: fit is a prefix
            const int total = (int)__shfl_sync(kFull, incl, Bp - 1);
            const int my_lo = (int)(incl - (hub ? 0u : el));   // first edge slot of this lane's entry (lane < Bp)
            PROF(0);
            if (HASH && S.nt + total + 1 > table_limit(S.lg)) {
                if (!table_grow(S, S.nt + total + 1, lane)) { result = WALK_TABLE_OVERFLOW; break; }
                PROF(1);
                PROF_ADD(13, 1);
            }
            PROF_ADD(9, 1);
            PROF_ADD(10, Bp);
            PROF_ADD(11, total);
            if (lane < Bp) {
                W.ebeg[lane] = eb;
                W.epre[lane + 1] = (int)incl;
                W.erw[lane] = erw;
            }
            if (lane == 0) {
                W.epre[0] = 0;
                W.enqmask = 0ull;
            }
            __syncwarp();

            // ---- (C) neighbour ids + weights of all edge slots ----
            int vk[kU], jk[kU], qk[kU];
            double wgt[kU], dv[kU];
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                const int e = k * 32 + lane;
                vk[k] = -1;
                jk[k] = 0;
                qk[k] = 0;
                wgt[k] = 0.0;
                dv[k] = 1.0;
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
                    dv[k] = P.edge_din[at];
                    wgt[k] = uni ? W.erw[lo] : ld_weight(P.w + at);
                }
            }
            // node records of the incoming queue entries (their ids have arrived by now); used next iteration
            if (incoming && n_main0 != 0) {
                const NodeInfo t = ld_info(&P.info[nu]);
                nd = t.d_in; nb = t.begin; nl = t.len;
                if (uni) nrw = P.row_w[nu];
            }
            // ---- (D) who references which node: cref = edge slots, centry = batch entries, per distinct node ----
            unsigned own = 0;   // bit k: this lane's edge slot k created the cache slot, bit kU: its batch entry did
            int qe = 0;
            if (lane < Bp) {
                bool o;
                qe = cache_insert(W.ckey, eu, o);
                if (o) own |= 1u << kU;
                atomicOr(&W.u.b.centry[qe], 1u << lane);
            }
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if (vk[k] >= 0) {
                    bool o;
                    qk[k] = cache_insert(W.ckey, vk[k], o);
                    if (o) own |= 1u << k;
                    atomicOr(&W.u.b.cref[qk[k]], 1ull << (k * 32 + lane));
                }
            __syncwarp();
            // An entry whose node is referenced EARLIER in the batch (by the row of an earlier entry, or as an
            // earlier entry: a duplicate in the queue) needs the value those pushes leave: the batch ends before
            // it.  Every entry that stays can then be checked and its push coefficient computed at once.
            // A node that is simply in the queue twice with no such reference in between needs no such care: its
            // second pop sees what the first one left (r = 0 after a push, or the same failing value) and does
            // nothing; the first of the entries owns the node.
            unsigned long long mref_e = 0ull;
            bool dep = false, dupl = false;
            if (lane < Bp) {
                mref_e = W.u.b.cref[qe];
                dep = (mref_e & below64(my_lo)) != 0ull;
                dupl = (W.u.b.centry[qe] & ((1u << lane) - 1u)) != 0u;
            }
            const unsigned depm = __ballot_sync(kFull, dep);
            const int j_lim = depm ? __ffs(depm) - 1 : Bp;   // >= 1
            const int e_lim = W.epre[j_lim];
            const unsigned long long in64 = below64(e_lim);
            // per edge slot: the node's references inside the batch, whether an entry owns it, whether this slot leads
            unsigned long long mref_k[kU];
            unsigned f_in = 0, f_lead = 0, f_single = 0;
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                mref_k[k] = 0ull;
                const int e = k * 32 + lane;
                if (vk[k] >= 0 && e < e_lim) {
                    f_in |= 1u << k;
                    mref_k[k] = W.u.b.cref[qk[k]] & in64;
                    const bool entry_owned = (W.u.b.centry[qk[k]] & below32(j_lim)) != 0u;   // its first entry owns it
                    if (!entry_owned && (mref_k[k] & (0ull - mref_k[k])) == (1ull << e)) {
                        f_lead |= 1u << k;
                        if ((mref_k[k] & (mref_k[k] - 1ull)) == 0ull) f_single |= 1u << k;
                    }
                }
            }
            PROF(2);
            PROF_ADD(15, j_lim < Bp ? 1 : 0);

            // ---- (E) state of every distinct node, fetched by its leading reference ----
            double os[kU], orr[kU];
            unsigned tk[kU + 1];
            unsigned f_new = 0;   // the node was not part of the walk yet
            double es = 0.0, er = 0.0;
            const bool e_owner = lane < j_lim && !dupl;   // this lane's entry is the first (or only) one of its node
#pragma unroll
            for (int k = 0; k <= kU; ++k) tk[k] = 0;
#pragma unroll
            for (int k = 0; k < kU; ++k) os[k] = orr[k] = 0.0;
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if ((f_lead >> k) & 1u) {
                    bool is_new;
                    if (HASH) tk[k] = table_touch(S.T, S.lg, vk[k], os[k], orr[k], is_new);
                    else direct_touch(S.T, S.epoch, vk[k], os[k], orr[k], is_new);
                    if (is_new) f_new |= 1u << k;
                }
            if (e_owner) {   // queue entries are part of the walk: found / tagged for sure
                TableEntry e;
                e.s = e.r = 0.0;
                if (HASH) table_find(S.T, S.lg, eu, e, tk[kU]);
                else e = ld_entry(S.T + eu);
                es = e.s;
                er = e.r;
            }
            PROF(3);

            // ---- (F) pop checks of all entries (similarity.py:204); a long row that passes ends the batch ----
            const bool pass = e_owner && ((first_in && lane == 0) || quot_ge(er, ed, eps));
            const unsigned stopm = __ballot_sync(kFull, pass && hub);
            const int j_stop = stopm ? __ffs(stopm) - 1 : j_lim;
            const bool act = pass && lane < j_stop;
            const unsigned actm = __ballot_sync(kFull, act);
            const double c = act ? __dmul_rn(P.one_minus_rho, er) : 0.0;   // push.py:57
            if (act) er = 0.0;                                               // push.py:60

            // ---- (G) contributions; a node with one reference is finished in registers ----
            unsigned f_act = 0, f_enq = 0;
#pragma unroll
            for (int k = 0; k < kU; ++k) {
                const double ce = __shfl_sync(kFull, c, jk[k]);
                if (((f_in >> k) & 1u) && ((actm >> jk[k]) & 1u)) {
                    const double p = __dmul_rn(ce, wgt[k]);
                    f_act |= 1u << k;
                    if ((f_single >> k) & 1u) {
                        os[k] = __dadd_rn(os[k], p);     // push.py:63
                        orr[k] = __dadd_rn(orr[k], p);   // push.py:64
                        if (quot_ge(orr[k], dv[k], eps)) f_enq |= 1u << k;   // similarity.py:214
                    } else {
                        W.u.b.contrib[k * 32 + lane] = p;
                    }
                }
            }
            unsigned long long act64 = 0ull;
#pragma unroll
            for (int k = 0; k < kU; ++k) act64 |= (unsigned long long)__ballot_sync(kFull, (f_act >> k) & 1u) << (32 * k);
            const bool any_multi = __any_sync(kFull, (f_act & ~f_single) != 0u);
            unsigned f_mod = f_act & f_single;   // leading slots whose value changed
            bool e_mod = act;
            if (any_multi) {
                // ---- (H) a node with several references: its leading reference (or the entry that IS the node)
                //      adds the contributions in edge-slot order = queue order, then CSR order ----
                __syncwarp();
                unsigned long long flags = 0ull;
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (((f_lead & ~f_single) >> k) & 1u) {
                        unsigned long long todo = mref_k[k] & act64;
                        while (todo) {
                            const int e2 = __ffsll((long long)todo) - 1;
                            todo &= todo - 1ull;
                            const double p = W.u.b.contrib[e2];
                            os[k] = __dadd_rn(os[k], p);
                            orr[k] = __dadd_rn(orr[k], p);
                            f_mod |= 1u << k;
                            if (quot_ge(orr[k], dv[k], eps)) flags |= 1ull << e2;
                        }
                    }
                if (e_owner) {   // references to an entry's node all come after its pop
                    unsigned long long todo = mref_e & in64 & act64;
                    while (todo) {
                        const int e2 = __ffsll((long long)todo) - 1;
                        todo &= todo - 1ull;
                        const double p = W.u.b.contrib[e2];
                        es = __dadd_rn(es, p);
                        er = __dadd_rn(er, p);
                        e_mod = true;
                        if (quot_ge(er, ed, eps)) flags |= 1ull << e2;
                    }
                }
                if (flags) atomicOr(&W.enqmask, flags);
                __syncwarp();
                const unsigned long long em = W.enqmask;
#pragma unroll
                for (int k = 0; k < kU; ++k)
                    if (((f_act & ~f_single) >> k) & 1u)
                        if ((em >> (k * 32 + lane)) & 1ull) f_enq |= 1u << k;
            }
            PROF(4);

            // ---- (I) ordered append: edge slots are numbered entry by entry, each row in CSR order ----
            {
                unsigned m_enq[kU];
                unsigned cnt = 0;
#pragma unroll
                for (int k = 0; k < kU; ++k) {
                    m_enq[k] = __ballot_sync(kFull, (f_enq >> k) & 1u);
                    cnt += __popc(m_enq[k]);
                }
                if (tail - (head + (unsigned)j_stop) + cnt > (unsigned)P.queue_cap) {
                    result = WALK_RING_OVERFLOW;
                } else {
                    unsigned before = 0;
#pragma unroll
                    for (int k = 0; k < kU; ++k) {
                        if ((f_enq >> k) & 1u) S.queue[(tail + before + __popc(m_enq[k] & lt)) & qmask] = vk[k];
                        before += __popc(m_enq[k]);
                    }
                    // counters exactly as the one-at-a-time walk would leave them
                    ws.pushes += __popc(actm);
                    ws.enq += cnt;
                    unsigned elen = act ? el : 0u;
                    unsigned qlen = 0;
                    if (act) {   // queue length right after this entry's push
                        const int upto = (int)incl;   // edge slots below this belong to entries <= lane
                        unsigned cum = 0;
#pragma unroll
                        for (int k = 0; k < kU; ++k) cum += __popc(m_enq[k] & below32(upto - 32 * k));
                        qlen = tail + cum - (head + (unsigned)lane + 1u);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        elen += __shfl_xor_sync(kFull, elen, o);
                        qlen = max(qlen, __shfl_xor_sync(kFull, qlen, o));
                    }
                    ws.edges += elen;
                    if (qlen > ws.maxq) ws.maxq = qlen;
                    tail += cnt;
                }
            }
            // ---- (J) write back what changed (a hash entry that was claimed but not added to is completed) ----
            if (result == WALK_OK) {
                if (HASH) {
#pragma unroll
                    for (int k = 0; k < kU; ++k)
                        if (((f_mod | f_new) >> k) & 1u) st_entry(S.T + tk[k], os[k], orr[k], dv[k], vk[k]);
                    if (e_mod) st_entry(S.T + tk[kU], es, er, ed, eu);
#pragma unroll
                    for (int k = 0; k < kU; ++k) S.nt += __popc(__ballot_sync(kFull, (f_new >> k) & 1u));
                } else {
#pragma unroll
                    for (int k = 0; k < kU; ++k) {
                        const bool wr = (f_mod >> k) & 1u;
                        if (wr) st_entry(S.T + vk[k], os[k], orr[k], dv[k], S.epoch);
                        const bool is_new = wr && ((f_new >> k) & 1u);   // joins the walk: listed for the sweep
                        const unsigned m_new = __ballot_sync(kFull, is_new);
                        if (is_new) S.touched[S.nt + __popc(m_new & lt)] = vk[k];
                        S.nt += __popc(m_new);
                    }
                    if (e_mod) st_entry(S.T + eu, es, er, ed, S.epoch);
                }
            }
            // ---- (K) hand the node cache back empty ----
            if ((own >> kU) & 1u) {
                W.ckey[qe] = -1;
                W.u.b.cref[qe] = 0ull;
                W.u.b.centry[qe] = 0u;
            }
#pragma unroll
            for (int k = 0; k < kU; ++k)
                if ((own >> k) & 1u) {
                    W.ckey[qk[k]] = -1;
                    W.u.b.cref[qk[k]] = 0ull;
                    W.u.b.centry[qk[k]] = 0u;
                }
            __syncwarp();
            PROF(5);
            if (result != WALK_OK) break;
            const bool hub_is_first = first_in && j_stop == 0;   // the seed itself has a long row
            first = false;

            int consumed = j_stop;
            head += (unsigned)j_stop;
            if (stopm) {
                // the entry at j_stop is a long row whose pop check passed: push it now, alone
                const int u0 = __shfl_sync(kFull, eu, j_stop);
                NodeInfo i0;
                i0.d_in = __shfl_sync(kFull, ed, j_stop);
                i0.begin = __shfl_sync(kFull, eb, j_stop);
                i0.len = __shfl_sync(kFull, el, j_stop);
                const double rw0 = __shfl_sync(kFull, erw, j_stop);
                head += 1u;
                consumed += 1;
                result = push_hub_row<HASH>(P, W, S, ws, u0, i0, rw0, eps, hub_is_first, head, tail, lane, lt);
                PROF(6);
                PROF_ADD(14, 1);
                if (result != WALK_OK) break;
            }
            // ---- shift the entries that were not consumed to the front, append the incoming ones ----
            {
                const int n_keep = n_main - consumed;            // >= 0
                const int n_inc = n_main0 == 0 ? 0 : want - n_main0;
                const int d = consumed & 31;
                const int t_u = __shfl_down_sync(kFull, eu, d), t_nu = __shfl_down_sync(kFull, nu, d);
                const double t_d = __shfl_down_sync(kFull, ed, d), t_nd = __shfl_down_sync(kFull, nd, d);
                const double t_w = __shfl_down_sync(kFull, erw, d), t_nw = __shfl_down_sync(kFull, nrw, d);
                const unsigned t_b = __shfl_down_sync(kFull, eb, d), t_nb = __shfl_down_sync(kFull, nb, d);
                const unsigned t_l = __shfl_down_sync(kFull, el, d), t_nl = __shfl_down_sync(kFull, nl, d);
                if (lane < n_keep) { eu = t_u; ed = t_d; erw = t_w; eb = t_b; el = t_l; }
                else if (lane < n_keep + n_inc) { eu = t_nu; ed = t_nd; erw = t_nw; eb = t_nb; el = t_nl; }
                n_main = n_keep + n_inc;
            }
        }

        if (P.debug_keep) {   // operator seam: dense s and r of this one walk
            if (HASH) {
                const int C = 1 << S.lg;
                for (int i = lane; i < C; i += 32) {
                    const TableEntry e = ld_entry(S.T + i);
                    if (e.key != kEmptyKey) {
                        if (result == WALK_OK) {
                            const int o = P.from_walk ? P.from_walk[e.key] : e.key;
                            P.dbg_s[o] = e.s;
                            P.dbg_r[o] = e.r;
                        }
                        S.T[i].key = kEmptyKey;
                    }
                }
            } else if (result == WALK_OK) {
                __syncwarp();
                for (int i = lane; i < S.nt; i += 32) {
                    const int x = S.touched[i];
                    const TableEntry e = ld_entry(S.T + x);
                    const int o = P.from_walk ? P.from_walk[x] : x;
                    P.dbg_s[o] = e.s;
                    P.dbg_r[o] = e.r;
                }
            }
            if (lane == 0) {
                P.counters[PC_PUSHES] = ws.pushes;
                P.counters[PC_TOUCHED] = (unsigned long long)S.nt;
                P.counters[PC_OVERFLOW_SEEDS] = result == WALK_OK ? 0ull : 1ull;
                if (result == WALK_TABLE_OVERFLOW) P.counters[PC_TOVERFLOW] = 1ull;
            }
            break;
        }

        if (result != WALK_OK) {
            // undo and hand the seed to the retry pass
            for (int i = lane; i < kCacheSlots; i += 32) {
                W.ckey[i] = -1;
                W.u.b.cref[i] = 0ull;
                W.u.b.centry[i] = 0u;
            }
            if (HASH) table_clear(S, lane);   // (direct-mapped: the next walk's epoch makes these entries stale)
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
            const Threshold tau = make_threshold(warp_min(q));
            // one linear scan: support count, members (ties included, arcte.py:363-367), table reset
            const int C = 1 << S.lg;
            for (int i0 = 0; i0 < C; i0 += 128) {
                TableEntry e[4];
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) e[k2] = ld_entry(S.T + i0 + k2 * 32 + lane);   // C >= 1024: in range
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const bool occ = e[k2].key != kEmptyKey;
                    const bool in_sup = occ && e[k2].s != 0.0;
                    const bool pass = in_sup && quot_ge(e[k2].s, e[k2].d_in, tau);
                    if (occ) S.T[i0 + k2 * 32 + lane].key = kEmptyKey;
                    support += __popc(__ballot_sync(kFull, in_sup));
                    const unsigned mp = __ballot_sync(kFull, pass);
                    if (pass) S.touched[m + __popc(mp & lt)] = e[k2].key;
                    m += __popc(mp);
                }
            }
        } else {
            // every node of N(seed) + seed was tagged by the seed's own push
            q = __ddiv_rn(ld_entry(S.T + seed).s, si.d_in);
            for (unsigned j0 = 0; j0 < si.len; j0 += 64) {
                int v[2];
                TableEntry o[2];
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const unsigned j = j0 + k2 * 32 + lane;
                    v[k2] = j < si.len ? P.indices[si.begin + j] : -1;
                }
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2)
                    if (v[k2] >= 0) o[k2] = ld_entry(S.T + v[k2]);
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2)
                    if (v[k2] >= 0) {
                        if (o[k2].key == S.epoch) q = fmin(q, __ddiv_rn(o[k2].s, o[k2].d_in));   // arcte.py:355-356
                        else q = fmin(q, __ddiv_rn(0.0, ld_info_din(&P.info[v[k2]])));
                    }
            }
            const Threshold tau = make_threshold(warp_min(q));   // arcte.py:359-360
            // sweep over the touched list: one 32-byte gather per node (in-degree included), nothing is reset
            for (int i0 = 0; i0 < S.nt; i0 += 128) {
                int x[4];
                TableEntry o[4];
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const int i = i0 + k2 * 32 + lane;
                    x[k2] = i < S.nt ? S.touched[i] : -1;
                }
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2)
                    if (x[k2] >= 0) o[k2] = ld_entry(S.T + x[k2]);
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    if (i0 + k2 * 32 >= S.nt) break;   // warp-uniform
                    bool in_sup = false, pass = false;
                    if (x[k2] >= 0) {
                        in_sup = o[k2].s != 0.0;
                        pass = in_sup && quot_ge(o[k2].s, o[k2].d_in, tau);
                    }
                    support += __popc(__ballot_sync(kFull, in_sup));
                    const unsigned mp = __ballot_sync(kFull, pass);
                    if (pass) S.touched[m + __popc(mp & lt)] = x[k2];
                    m += __popc(mp);
                }
            }
        }
        __syncwarp();
        PROF(7);
        PROF_ADD(12, HASH ? (1 << S.lg) : S.nt);
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
                for (int i = lane; i < m; i += 32)
                    P.members[off + i] = P.from_walk ? P.from_walk[S.touched[i]] : S.touched[i];   // arcte.py:372-376
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
            wtot[WS_PUSHES] += ws.pushes;
            wtot[WS_EDGES] += ws.edges;
            wtot[WS_ENQ] += ws.enq;
            if (ws.maxq > wtot[WS_MAXQ]) wtot[WS_MAXQ] = ws.maxq;
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

    PROF_FLUSH();
    if (lane == 0) {
        P.tbl_clean[slot * 2 + 0] = HASH ? S.clean[0] : S.epoch;
        P.tbl_clean[slot * 2 + 1] = HASH ? S.clean[1] : 0;
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
    if (wps > 32) wps = 32;
    wps = ((wps + kWarpsPerCta - 1) / kWarpsPerCta) * kWarpsPerCta;
    return wps;
}

// Bytes of one slot: hash = two table halves of cap entries + a staging list of cap ints; direct-mapped =
// one entry and one list cell per node.
static double slot_bytes(int engine, int64_t cap, int64_t qcap)
{
    return (engine == ARCTE_ENGINE_BATCHED_HASH ? 68.0 : 36.0) * (double)cap + 4.0 * (double)qcap;
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
    BatchedPool &bp = c->bpool;
    int64_t cap;
    if (engine == ARCTE_ENGINE_BATCHED_HASH) {
        // table region: two halves of cap entries; cap large enough for every node at 62.5 % load, bounded by memory
        cap = c->tbl_cap_cfg > 0 ? pow2_ge(c->tbl_cap_cfg) : pow2_ge(2 * c->n);
        if (cap < (1 << kInitLg)) cap = 1 << kInitLg;
        if (cap > ((int64_t)1 << 26)) cap = (int64_t)1 << 26;
    } else {
        cap = c->n;
    }
    if (bp.mode == engine && bp.n_slots >= want && (engine == ARCTE_ENGINE_BATCHED_HASH ? bp.cap <= cap && bp.cap >= (1 << kInitLg) && c->tbl_cap_cfg == bp.cap_cfg
                                                                                         : bp.cap == cap)) {
        bp.plan_cap = bp.cap;   // keep the pool as it is
        *n_slots = want;
        *queue_cap = qcap;
        return ARCTE_OK;
    }
    size_t free_b = 0, total_b = 0;
    ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += bp.tbl.bytes + bp.stage.bytes + bp.queue.bytes;
    const int pct = c->mem_percent > 0 ? c->mem_percent : 60;
    const double budget = (double)free_b * pct / 100.0;
    if (engine == ARCTE_ENGINE_BATCHED_HASH)
        while (cap > (1 << kInitLg) && (double)want * slot_bytes(engine, cap, qcap) > budget) cap >>= 1;
    if ((double)want * slot_bytes(engine, cap, qcap) > budget) {
        want = (int64_t)(budget / slot_bytes(engine, cap, qcap));
        want = (want / kWarpsPerCta) * kWarpsPerCta;
        if (want < kWarpsPerCta) { set_error("not enough device memory for the walk states of this graph"); return ARCTE_E_NOMEM; }
    }
    bp.plan_cap = cap;
    *n_slots = want;
    *queue_cap = qcap;
    return ARCTE_OK;
}

int batched_ensure(arcte_cuda_ctx *c, int engine, int64_t n_slots, int64_t qcap)
{
    if (!c->row_w_valid) { set_error("batched engine: graph not prepared"); return ARCTE_E_ARG; }
    BatchedPool &bp = c->bpool;
    const int64_t cap = bp.plan_cap;
    const bool hash = engine == ARCTE_ENGINE_BATCHED_HASH;
    if (!(bp.mode == engine && bp.n_slots >= n_slots && bp.cap == cap)) {
        dev_free(bp.tbl);
        dev_free(bp.stage);
        dev_free(bp.clean);
        dev_free(bp.queue);
        bp.n_slots = bp.queue_slots = 0;
        bp.queue_cap = 0;
        bp.mode = -1;
        const size_t entries = (size_t)(hash ? 2 : 1) * (size_t)n_slots * (size_t)cap;
        ARCTE_TRY(dev_reserve(bp.tbl, sizeof(TableEntry) * entries));
        ARCTE_TRY(dev_reserve(bp.stage, sizeof(int32_t) * (size_t)n_slots * (size_t)cap));
        ARCTE_TRY(dev_reserve(bp.clean, sizeof(int32_t) * 2 * (size_t)n_slots));
        ARCTE_CUDA_TRY(cudaMemsetAsync(bp.clean.p, 0, sizeof(int32_t) * 2 * (size_t)n_slots, c->stream));
        // hash: nothing of the tables is initialised here; a warp makes the part it is about to use all-EMPTY the
        // first time it needs it (tbl_clean), so an extraction never pays for the whole region.
        // direct-mapped: epoch 0 everywhere, once; walks tag their entries with epochs 1, 2, ...
        if (!hash) ARCTE_CUDA_TRY(cudaMemsetAsync(bp.tbl.p, 0, sizeof(TableEntry) * entries, c->stream));
        bp.n_slots = n_slots;
        bp.cap = cap;
        bp.cap_cfg = c->tbl_cap_cfg;
        bp.mode = engine;
    }
    if (bp.queue_cap != qcap || bp.queue_slots < n_slots) {
        dev_free(bp.queue);
        ARCTE_TRY(dev_reserve(bp.queue, sizeof(int32_t) * (size_t)bp.n_slots * (size_t)qcap));
        bp.queue_cap = qcap;
        bp.queue_slots = bp.n_slots;
    }
    return ARCTE_OK;
}

void batched_fill_params(arcte_cuda_ctx *c, int engine, PushParams &P)
{
    (void)engine;
    const BatchedPool &bp = c->bpool;
    P.uniform_rows = c->uniform_rows ? 1 : 0;
    P.row_w = c->row_w.as<double>();
    P.edge_din = c->edge_din.as<double>();
    P.tbl = bp.tbl.as<TableEntry>();
    P.tbl_cap_max = bp.cap;
    P.tbl_clean = bp.clean.as<int32_t>();
    P.touched = bp.stage.as<int32_t>();
    P.touched_stride = bp.cap;
    P.queue = bp.queue.as<int32_t>();
    P.queue_cap = bp.queue_cap;
    P.sr = nullptr;
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
