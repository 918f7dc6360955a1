// push_frontier.cu -- the SECOND walk schedule of the push engine (opt-in): synchronous
// frontier rounds on fixed-point state, one CTA per seed.
//
// Why it exists.  The exact FIFO replay of push.cu is bit-identical to the reference but its
// pushes are sequential inside a seed, so thousands of walks must be in flight to hide memory
// latency, their dense states (gigabytes) live in DRAM and the kernel runs at the random-access
// rate of HBM.  The epsilon-push iteration does not depend on the push
// order for its guarantees: any order keeps  s + sum_v r[v] (G_v - e_v) = G_seed  and ends with
// r[v]/d_in[v] < eps everywhere, hence 0 <= (G_seed - s)[x]/d[x] < eps (1-rho)/rho  (SURVEY.md
// section 8a, error-bound note) -- the tolerance BASELINE.json's north_star states.  This file
// uses that freedom for parallelism INSIDE a seed: all nodes over the threshold are pushed
// together, a whole CTA works on one seed, only a few hundred walks are in flight and the
// sectors they touch stay in the 126 MB L2.
//
// Reference semantics kept (paths relative to /root/reference/reveal_graph_embedding/):
//   initial state and the unconditional first push    eps_randomwalk/similarity.py:176-192
//   push rule  c = (1-rho) r[u]; r[u] = 0; s[N] += c w; r[N] += c w   eps_randomwalk/push.py:56-64
//   a node is pushed when r[v]/d_in[v] >= eps         eps_randomwalk/similarity.py:194-216
//   threshold / membership / emission                 embedding/arcte/arcte.py:352-376
// Deliberately different: the ORDER of pushes (rounds instead of a FIFO) and the number format
// of s and r.
//
// Determinism.  Rounds are Jacobi steps: phase 1 takes the residual of every frontier node,
// phase 2 distributes all of them; s and r are unsigned 64-bit fixed-point numbers (unit
// 2^-F), so the atomic additions of a round commute exactly.  The state after every round,
// the frontier SETS and therefore the communities do not depend on thread timing, CTA size,
// walk-slot count or seed sharding; the test suite restates the schedule on the CPU and compares
// bit for bit.  Against the reference the support
// is identical up to documented in-band ties (tests/test_frontier_schedule.py).
#include "common.cuh"
#include "push.cuh"

// Timing experiments only (results are WRONG when a bit is set): 1 = no list appends, 2 = no
// in-degree gather / threshold test, 4 = no state access, 8 = no binary search (edge e reads CSR
// position e), 16 = no CSR reads.
#ifndef ARCTE_FRONTIER_ABLATE
#define ARCTE_FRONTIER_ABLATE 0
#endif
#ifndef ARCTE_FRONTIER_PRELOAD
#define ARCTE_FRONTIER_PRELOAD 0
#endif

#ifdef ARCTE_FRONTIER_PROFILE
#define PROF_DECL long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_t = clock64();
#define PROF(k) do { const long long _n = clock64(); prof[k] += _n - prof_t; prof_t = _n; } while (0)
#define PROF_COUNT(k, x) prof[k] += (x)
#define PROF_PARAMS , long long *prof, long long &prof_t
#define PROF_ARGS , prof, prof_t
#else
#define PROF_DECL
#define PROF(k) do { } while (0)
#define PROF_COUNT(k, x) do { } while (0)
#define PROF_PARAMS
#define PROF_ARGS
#endif

namespace arcte {

namespace {

__device__ __forceinline__ unsigned long long ld_fixed(const unsigned long long *p) { return __ldcg(p); }
__device__ __forceinline__ double fixed_to_double(unsigned long long x, double inv_scale)
{
    return __dmul_rn(__ull2double_rn(x), inv_scale);
}
// similarity.py:194 / :204 / :214: r[v] / in_degree[v] >= epsilon
__device__ __forceinline__ bool over_threshold(unsigned long long r, double d_in, double eps, double inv_scale)
{
    return __ddiv_rn(fixed_to_double(r, inv_scale), d_in) >= eps;
}

template <int T> struct WalkShared {
    int cur_n, next_n, nt, m, dummy;
    long long work;
    long long off;
    unsigned long long edges;
    unsigned wsum[32];
    unsigned eoff[T];     // per frontier entry of the current chunk: first edge on the chunk's edge line
    unsigned ebeg[T];     //   first CSR position of its row
    double ec[T];          //   mass it distributes this round
    double red[32];
    unsigned long long tot[10];  // pushes, edges, enqueues, max frontier, support, touched, seed degree, members, emitted, rounds
};

// Two neighbour touches per lane, all lanes of the warp together (inactive lanes pass pf = 0):
// push.py:63-64 with atomics, first-touch bookkeeping and threshold-crossing detection.  Exactly
// one addition moves r[v] across the threshold in a round (additions are positive and r[v] is
// not reset during phase 2), so v enters the next frontier exactly once.  The exact threshold
// test (two divisions) runs only when the new residual is within 1e-9 of the threshold or
// above; below that the quotient cannot reach eps (the quotient is monotone and its rounding
// error is 2^-53), so skipping the test cannot change the outcome.
template <typename SH>
__device__ __forceinline__ void touch2(const PushParams &P, unsigned long long *__restrict__ sr,
                                       int32_t *__restrict__ touched, int32_t *__restrict__ next, SH &sh,
                                       const int (&v)[2], const unsigned long long (&pf)[2], double eps, int lane,
                                       unsigned lt, bool second PROF_PARAMS)
{
    unsigned long long old_s[2], old_r[2], add[2];
    double d[2];
#if ARCTE_FRONTIER_PRELOAD
    // Experiment switch (default off, measured neutral: profiles/r1_frontier_schedule.md): fetch
    // the pair with an ordinary load first and make the atomics depend on it, so that they hit
    // L2.  (s < 2^63 always: the shifted value is zero.)
    ulonglong2 pre[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (pf[k]) pre[k] = __ldcg(reinterpret_cast<const ulonglong2 *>(&sr[2 * (int64_t)v[k]]));
#pragma unroll
    for (int k = 0; k < 2; ++k) add[k] = pf[k] ? pf[k] + (pre[k].x >> 63) : 0ull;
#else
#pragma unroll
    for (int k = 0; k < 2; ++k) add[k] = pf[k];
#endif
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (pf[k]) {
#if ARCTE_FRONTIER_ABLATE & 4
            old_s[k] = 1ull; old_r[k] = add[k];
#else
            old_s[k] = atomicAdd(&sr[2 * (int64_t)v[k]], add[k]);
            old_r[k] = atomicAdd(&sr[2 * (int64_t)v[k] + 1], add[k]);
#endif
#if ARCTE_FRONTIER_ABLATE & 2
            d[k] = 1.0e30;
#else
            d[k] = P.info[v[k]].d_in;
#endif
        }
    }
#ifdef ARCTE_FRONTIER_PROFILE
    if (((old_s[0] ^ old_r[0] ^ old_s[1] ^ old_r[1]) == 0x123456789abcdefull) || d[0] + d[1] == -1.0) prof[9] += 1;  // consume
    PROF(9);  // atomics + in-degree returned
#endif
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !second) break;  // warp-uniform
        bool is_new = false, cross = false;
        if (pf[k]) {
            is_new = old_s[k] == 0ull;
            const double rn = fixed_to_double(old_r[k] + pf[k], P.inv_scale);
            if (rn >= __dmul_rn(__dmul_rn(eps, d[k]), 0.999999999))
                cross = over_threshold(old_r[k] + pf[k], d[k], eps, P.inv_scale) &&
                        !over_threshold(old_r[k], d[k], eps, P.inv_scale);
        }
#if ARCTE_FRONTIER_ABLATE & 1
        if (is_new && v[k] == -7) touched[0] = v[k];
        if (cross && v[k] == -7) next[0] = v[k];
        continue;
#endif
        const unsigned m_new = __ballot_sync(kFull, is_new);
        if (m_new) {
            int base = 0;
            if (lane == __ffs(m_new) - 1) base = atomicAdd(&sh.nt, __popc(m_new));
#if ARCTE_FRONTIER_ABLATE & 64   // sensitivity test: every append pays a second shared atomic
            if (lane == __ffs(m_new) - 1) base += atomicAdd(&sh.dummy, 0);
#endif
            base = __shfl_sync(kFull, base, __ffs(m_new) - 1);
#if ARCTE_FRONTIER_ABLATE & 128
            if (is_new && base < 0) touched[0] = v[k];
#else
            if (is_new) touched[base + __popc(m_new & lt)] = v[k];
#endif
        }
        const unsigned m_x = __ballot_sync(kFull, cross);
        if (m_x) {
            int base = 0;
            if (lane == __ffs(m_x) - 1) base = atomicAdd(&sh.next_n, __popc(m_x));
            base = __shfl_sync(kFull, base, __ffs(m_x) - 1);
            if (cross) next[base + __popc(m_x & lt)] = v[k];
        }
    }
}

template <int T>
__global__ void __launch_bounds__(T) k_push_frontier(const PushParams P)
{
    constexpr int NW = T / 32;
    __shared__ WalkShared<T> sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int64_t slot = blockIdx.x;
    if (slot >= P.n_slots) return;
    unsigned long long *__restrict__ sr = reinterpret_cast<unsigned long long *>(P.sr + slot * P.n);
    int32_t *__restrict__ touched = P.touched + slot * P.n;
    int32_t *fa = P.frontier + slot * 2 * P.n;
    int32_t *fb = fa + P.n;
    double *__restrict__ fval = P.fval + slot * P.n;
    if (tid < 10) sh.tot[tid] = 0ull;
    __syncthreads();
    PROF_DECL

    for (;;) {
        if (tid == 0) sh.work = (long long)atomicAdd(&P.counters[P.cursor], 1ull);
        __syncthreads();
        const long long k = sh.work;
        if (k >= P.n_work) break;
        const int pos = P.work_ids ? P.work_ids[k] : (int)(P.work_lo + k);
        const int seed = P.work_seed[pos];
        const double eps = P.work_eps[pos];
        const unsigned long long one = (unsigned long long)__double2ll_rn(P.scale);

        // similarity.py:176-177: s[seed] = r[seed] = 1; the seed is the first frontier
        if (tid == 0) {
            __stcg(&sr[2 * (int64_t)seed], one);
            __stcg(&sr[2 * (int64_t)seed + 1], one);
            touched[0] = seed;
            fa[0] = seed;
            sh.nt = 1;
            sh.cur_n = 1;
            sh.next_n = 0;
            sh.edges = 0ull;
            sh.dummy = 0;
        }
        __syncthreads();
        int32_t *cur = fa, *next = fb;
        unsigned long long pushes = 0, enq = 0, maxf = 0, rounds = 0;
        PROF(0);  // fetch + init

        for (;;) {
            const int cur_n = sh.cur_n;
            if (cur_n == 0) break;
            // ---- phase 1: take the residual of every frontier node (push.py:57, :60) ----
            unsigned long long my_edges = 0;
            for (int i = tid; i < cur_n; i += T) {
                const int u = cur[i];
                const unsigned long long r = ld_fixed(&sr[2 * (int64_t)u + 1]);
                __stcg(&sr[2 * (int64_t)u + 1], 0ull);
                fval[i] = __dmul_rn(P.one_minus_rho, fixed_to_double(r, P.inv_scale));
                my_edges += P.info[u].len;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_edges += __shfl_xor_sync(kFull, my_edges, o);
            if (lane == 0 && my_edges) atomicAdd(&sh.edges, my_edges);
            __syncthreads();
            PROF(1);  // phase 1

            // ---- phase 2: distribute (push.py:63-64), edge-balanced.  The frontier is taken T entries
            // at a time: each thread loads one entry's row and mass, a CTA-wide exclusive scan of the
            // row lengths lays the chunk's edges out on a line, and the threads walk that line T (x2)
            // edges at a time, finding the owning entry by binary search in shared memory.  Every lane
            // has an edge whatever the degree mix of the frontier. ----
            for (int c0 = 0; c0 < cur_n; c0 += T) {
                const int i = c0 + tid;
                unsigned len = 0, beg = 0;
                double c = 0.0;
                if (i < cur_n) {
                    const NodeInfo iu = P.info[cur[i]];
                    len = iu.len;
                    beg = iu.begin;
                    c = fval[i];
                }
                // exclusive scan of len over the CTA
                unsigned incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sh.wsum[warp] = incl;
                __syncthreads();
                unsigned wbase = 0, total = 0;
                for (int wv = 0; wv < NW; ++wv) {
                    const unsigned t = sh.wsum[wv];
                    if (wv < warp) wbase += t;
                    total += t;
                }
                sh.eoff[tid] = wbase + incl - len;
                sh.ebeg[tid] = beg;
                sh.ec[tid] = c;
                __syncthreads();
                PROF(2);  // chunk load + scan
                for (unsigned e0 = 0; e0 < total; e0 += 2 * T) {
                    int v[2];
                    unsigned long long pf[2];
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        const unsigned e = e0 + k2 * T + tid;
                        v[k2] = 0;
                        pf[k2] = 0ull;
                        if (e < total) {
                            int lo = 0, hi = T - 1;  // largest entry with eoff <= e
#if ARCTE_FRONTIER_ABLATE & 8
                            lo = (int)(e % (unsigned)T);
                            const unsigned j = sh.ebeg[0] + e;
#else
                            while (lo < hi) {
                                const int mid = (lo + hi + 1) >> 1;
                                if (sh.eoff[mid] <= e) lo = mid; else hi = mid - 1;
                            }
                            const unsigned j = sh.ebeg[lo] + (e - sh.eoff[lo]);
#endif
#if ARCTE_FRONTIER_ABLATE & 16
                            v[k2] = (int)((j * 2654435761u) % (unsigned)P.n);
                            pf[k2] = (unsigned long long)__double2ll_rn(__dmul_rn(__dmul_rn(sh.ec[lo], 1e-3), P.scale));
#else
                            v[k2] = P.indices[j];
                            pf[k2] = (unsigned long long)__double2ll_rn(__dmul_rn(__dmul_rn(sh.ec[lo], P.w[j]), P.scale));
#endif
                        }
                    }
#ifdef ARCTE_FRONTIER_PROFILE
                    if ((pf[0] ^ pf[1]) == 0x123456789abcdefull) prof[7] += 1;  // consume
                    PROF(7);  // search + row loads (emit time is folded into 3 in this build)
#endif
                    touch2(P, sr, touched, next, sh, v, pf, eps, lane, lt, e0 + T < total PROF_ARGS);
                    PROF_COUNT(8, 1);
                }
                PROF(3);  // edge loop (this thread)
                __syncthreads();
                PROF(4);  // waiting for the other warps
            }
            pushes += cur_n;
            rounds += 1;
            if ((unsigned long long)cur_n > maxf) maxf = cur_n;
            enq += sh.next_n;
            __syncthreads();
            if (tid == 0) {
                sh.cur_n = sh.next_n;
                sh.next_n = 0;
                }
            int32_t *t = cur; cur = next; next = t;
            __syncthreads();
            PROF(5);  // round bookkeeping
        }

        const int nt = sh.nt;
        if (P.debug_keep) {
            if (tid == 0) {
                P.counters[PC_PUSHES] = pushes;
                P.counters[PC_TOUCHED] = (unsigned long long)nt;
                P.counters[PC_OVERFLOW_SEEDS] = 0ull;
            }
            return;
        }

        // ---------------- threshold + membership (arcte.py:352-376) ----------------
        const NodeInfo si = P.info[seed];
        const int base_size = (int)si.len + 1;  // np.append(adjacent_nodes[n], n), arcte.py:358
        double q = __ddiv_rn(fixed_to_double(ld_fixed(&sr[2 * (int64_t)seed]), P.inv_scale), si.d_in);
        for (unsigned j = tid; j < si.len; j += T) {
            const int v = P.indices[si.begin + j];
            q = fmin(q, __ddiv_rn(fixed_to_double(ld_fixed(&sr[2 * (int64_t)v]), P.inv_scale), P.info[v].d_in));  // arcte.py:355-356
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q = fmin(q, __shfl_xor_sync(kFull, q, o));
        if (lane == 0) sh.red[warp] = q;
        if (tid == 0) sh.m = 0;
        __syncthreads();
        double tau = sh.red[0];
        for (int wv = 1; wv < NW; ++wv) tau = fmin(tau, sh.red[wv]);  // arcte.py:359-360
        // one sweep over the touched list: members (s/d_in >= tau, ties included: arcte.py:363-367)
        // are staged in the free frontier buffer, the state is zeroed (arcte.py:337-338)
        int32_t *stage = cur;
        for (int i0 = 0; i0 < nt; i0 += T) {
            const int i = i0 + tid;
            bool pass = false;
            int x = 0;
            if (i < nt) {
                x = touched[i];
                const unsigned long long sx = ld_fixed(&sr[2 * (int64_t)x]);
                __stcg(reinterpret_cast<ulonglong2 *>(&sr[2 * (int64_t)x]), make_ulonglong2(0ull, 0ull));
                pass = __ddiv_rn(fixed_to_double(sx, P.inv_scale), P.info[x].d_in) >= tau;
            }
            const unsigned mp = __ballot_sync(kFull, pass);
            if (mp) {
                int base = 0;
                if (lane == __ffs(mp) - 1) base = atomicAdd(&sh.m, __popc(mp));
                base = __shfl_sync(kFull, base, __ffs(mp) - 1);
                if (pass) stage[base + __popc(mp & lt)] = x;
            }
        }
        __syncthreads();
        PROF(6);  // tau + sweep
        const int m = sh.m;
        const bool emit = m > base_size;  // arcte.py:370
        if (tid == 0) {
            long long off = 0;
            if (emit) {
                if (P.retry_pass && P.seg_count[pos] > 0) off = P.seg_offset[pos];  // claimed by the pass that overflowed
                else off = (long long)atomicAdd(&P.counters[PC_MEMBER_CURSOR], (unsigned long long)m);
            }
            sh.off = off;
        }
        __syncthreads();
        const long long off = sh.off;
        const bool write = emit && off + m <= P.member_cap;
        if (write)
            for (int i = tid; i < m; i += T) P.members[off + i] = stage[i];  // arcte.py:372-376
        if (tid == 0) {
            if (emit) {
                P.seg_count[pos] = m;
                P.seg_offset[pos] = off;
                if (!write) {
                    const unsigned long long r = atomicAdd(&P.counters[PC_OVERFLOW_SEEDS], 1ull);
                    P.retry_list[r] = pos;
                }
            } else {
                P.seg_count[pos] = 0;
                P.seg_offset[pos] = 0;
            }
            if (!emit || write) {  // a seed whose members did not fit is re-run and counted then
                sh.tot[0] += pushes;
                sh.tot[1] += sh.edges;
                sh.tot[2] += enq;
                if (maxf > sh.tot[3]) sh.tot[3] = maxf;
                sh.tot[4] += (unsigned long long)nt;
                sh.tot[5] += (unsigned long long)nt;
                sh.tot[6] += si.len;
                if (emit) {
                    sh.tot[7] += (unsigned long long)m;
                    sh.tot[8] += 1;
                }
                sh.tot[9] += rounds;
            }
        }
        __syncthreads();
        PROF(3);  // emit (counted with the appends)
    }
#ifdef ARCTE_FRONTIER_PROFILE
    if (tid == 0)
        for (int k = 0; k < 10; ++k) atomicAdd(&P.counters[PC_PROF0 + k], (unsigned long long)prof[k]);
#endif

    if (tid == 0 && !P.debug_keep) {
        atomicAdd(&P.counters[PC_PUSHES], sh.tot[0]);
        atomicAdd(&P.counters[PC_EDGES], sh.tot[1]);
        atomicAdd(&P.counters[PC_ENQUEUES], sh.tot[2]);
        atomicMax(&P.counters[PC_MAXQ], sh.tot[3]);
        atomicAdd(&P.counters[PC_SUPPORT], sh.tot[4]);
        atomicAdd(&P.counters[PC_TOUCHED], sh.tot[5]);
        atomicAdd(&P.counters[PC_SEEDDEG], sh.tot[6]);
        atomicAdd(&P.counters[PC_MEMBERS], sh.tot[7]);
        atomicAdd(&P.counters[PC_EMITTED], sh.tot[8]);
        atomicAdd(&P.counters[PC_ROUNDS], sh.tot[9]);
    }
}

int launch_frontier_kernel(arcte_cuda_ctx *c, const PushParams &P, int threads, int64_t grid)
{
    if (grid < 1) return ARCTE_OK;
    switch (threads) {
    case 32: k_push_frontier<32><<<(unsigned)grid, 32, 0, c->stream>>>(P); break;
    case 64: k_push_frontier<64><<<(unsigned)grid, 64, 0, c->stream>>>(P); break;
    case 128: k_push_frontier<128><<<(unsigned)grid, 128, 0, c->stream>>>(P); break;
    case 256: k_push_frontier<256><<<(unsigned)grid, 256, 0, c->stream>>>(P); break;
    case 512: k_push_frontier<512><<<(unsigned)grid, 512, 0, c->stream>>>(P); break;
    case 1024: k_push_frontier<1024><<<(unsigned)grid, 1024, 0, c->stream>>>(P); break;
    default: set_error("frontier schedule: threads per walk must be 32, 64, 128, 256, 512 or 1024"); return ARCTE_E_ARG;
    }
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    return ARCTE_OK;
}

// Launch geometry: the head of the count-descending seed list (the long walks) can be given
// fewer, larger CTAs than the rest.  Measured on the YouTube shape the geometry hardly matters
// (3.9-4.9 s from 296 walks of 512 threads to 2580 walks of 32 threads); on small and medium
// graphs it does (profiles/r1_frontier_schedule.md), hence the two defaults below.
struct Geometry {
    int heavy_threads, heavy_ctas, light_threads, light_ctas, heavy_permille;
};
Geometry geometry(const arcte_cuda_ctx *c, int64_t n_work)
{
    Geometry g;
    g.heavy_permille = c->fr_heavy_permille >= 0 ? c->fr_heavy_permille : 0;
    g.heavy_threads = c->fr_heavy_threads > 0 ? c->fr_heavy_threads : 512;
    g.heavy_ctas = c->fr_heavy_ctas > 0 ? c->fr_heavy_ctas : 2;
    // Few seeds (up to a few rounds of the walks in flight): the launch is as long as its longest
    // walk, so a walk gets 512 threads (BA(5000,5): 1.4 ms against 3.5 ms at 64 threads and 18 ms for
    // the FIFO).  Many seeds: throughput counts, 64 threads x 16 walks per SM measured best.
    const bool few = n_work <= 50000;
    g.light_threads = c->fr_light_threads > 0 ? c->fr_light_threads : (few ? 512 : 64);
    g.light_ctas = c->fr_light_ctas > 0 ? c->fr_light_ctas : (few ? 2 : 16);
    return g;
}

}  // namespace

// F fractional bits such that s <= 1/rho (the geometric series of similarity.py:176-216) fits
// 63 bits with one bit to spare.
double frontier_scale(double rho)
{
    int f = 62;
    double top = 1.0 / rho + 1.0;
    while (top > 1.0 && f > 20) { top *= 0.5; --f; }
    return ldexp(1.0, f);
}

int frontier_plan_slots(arcte_cuda_ctx *c, int64_t n_work, int64_t *n_slots)
{
    const Geometry g = geometry(c, n_work);
    int64_t want = (int64_t)c->sm_count * (g.heavy_ctas > g.light_ctas ? g.heavy_ctas : g.light_ctas);
    if (want > n_work) want = n_work;
    if (want < 1) want = 1;
    if (c->slots.n == c->n && c->slots.n_slots >= want && c->slots.frontier_slots >= want && !c->slots.compact) {
        *n_slots = want;
        return ARCTE_OK;
    }
    size_t free_b = 0, total_b = 0;
    ARCTE_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->slots.sr.bytes + c->slots.touched.bytes + c->slots.queue.bytes + c->slots.frontier.bytes +
              c->slots.fval.bytes + c->slots.cmap.bytes;
    const int pct = c->mem_percent > 0 ? c->mem_percent : 60;
    const double per_slot = 36.0 * (double)c->n;  // state 16 + touched 4 + two frontiers 8 + taken mass 8
    const int64_t fit = (int64_t)((double)free_b * pct / 100.0 / per_slot);
    if (fit < 1) {
        set_error("not enough device memory for one frontier walk state of this graph");
        return ARCTE_E_NOMEM;
    }
    if (want > fit) want = fit;
    *n_slots = want;
    return ARCTE_OK;
}

int frontier_ensure_slots(arcte_cuda_ctx *c, int64_t want)
{
    SlotPool &sp = c->slots;
    if (!(sp.n == c->n && sp.n_slots >= want) || sp.compact) {   // (the compact engine leaves its pairs in sr)
        dev_free(sp.sr);
        dev_free(sp.touched);
        dev_free(sp.queue);
        dev_free(sp.cmap);
        dev_free(sp.cepoch);
        dev_free(sp.frontier);
        dev_free(sp.fval);
        sp.compact = false;
        sp.n_slots = sp.queue_slots = sp.frontier_slots = 0;
        sp.queue_cap = 0;
        ARCTE_TRY(dev_reserve(sp.sr, sizeof(double2) * (size_t)want * (size_t)c->n));
        ARCTE_TRY(dev_reserve(sp.touched, sizeof(int32_t) * (size_t)want * (size_t)c->n));
        ARCTE_CUDA_TRY(cudaMemsetAsync(sp.sr.p, 0, sizeof(double2) * (size_t)want * (size_t)c->n, c->stream));
        sp.n = c->n;
        sp.n_slots = want;
    }
    if (sp.frontier_slots < want) {
        dev_free(sp.frontier);
        dev_free(sp.fval);
        ARCTE_TRY(dev_reserve(sp.frontier, sizeof(int32_t) * 2 * (size_t)want * (size_t)c->n));
        ARCTE_TRY(dev_reserve(sp.fval, sizeof(double) * (size_t)want * (size_t)c->n));
        sp.frontier_slots = want;
    }
    return ARCTE_OK;
}

// One pass over P.n_work positions: the head of the list with the heavy geometry, the rest with
// the light one (a retry pass or a single seed: one launch).
int frontier_launch(arcte_cuda_ctx *c, PushParams P, int64_t n_work, bool retry_pass)
{
    const Geometry g = geometry(c, n_work);
    P.frontier = c->slots.frontier.as<int32_t>();
    P.fval = c->slots.fval.as<double>();
    const int64_t max_slots = P.n_slots;
    int64_t n_heavy = retry_pass || P.debug_keep ? 0 : (n_work * g.heavy_permille) / 1000;
    if (n_heavy > 0) {
        PushParams H = P;
        H.n_work = n_heavy;
        H.work_lo = 0;
        H.cursor = PC_WORK_CURSOR;
        int64_t grid = (int64_t)c->sm_count * g.heavy_ctas;
        if (grid > max_slots) grid = max_slots;
        if (grid > n_heavy) grid = n_heavy;
        H.n_slots = grid;
        ARCTE_TRY(launch_frontier_kernel(c, H, g.heavy_threads, grid));
    }
    if (n_work - n_heavy > 0) {
        PushParams L = P;
        L.n_work = n_work - n_heavy;
        L.work_lo = n_heavy;
        L.cursor = n_heavy > 0 ? PC_WORK_CURSOR2 : PC_WORK_CURSOR;
        if (P.work_ids) L.work_lo = 0;
        int64_t grid = (int64_t)c->sm_count * g.light_ctas;
        if (grid > max_slots) grid = max_slots;
        if (grid > L.n_work) grid = L.n_work;
        L.n_slots = grid;
        ARCTE_TRY(launch_frontier_kernel(c, L, P.debug_keep ? g.heavy_threads : g.light_threads, grid));
    }
    return ARCTE_OK;
}

}  // namespace arcte
