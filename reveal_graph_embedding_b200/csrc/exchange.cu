// exchange.cu -- (e) the one exchange step of the multi-GPU path, inside the library.
//
// Reference being replaced: the parent process sums the workers' n x n matrices
// (embedding/arcte/arcte.py:650-673).  Here rank r of G has walked the seeds at positions r, r + G, ... of
// the seed list (arcte.py:19-23) and owns the rows [n r / G, n (r+1) / G) of the result.  What rank d needs
// from rank r is, for every seed of r, the members that fall into d's row block.  So each rank
//   1. splits every community by destination row block, segment order and member order kept
//      (k_part_count, one scan, k_part_scatter: one warp per segment, ballot compaction per destination);
//   2. exchanges the split communities with ONE grouped NCCL send/recv (an all-to-all: 1/G of a rank's
//      members goes to each peer, nothing is sent twice, the 1.0 values are never sent);
//   3. assembles its own row block from the G parts (assemble.cu, unchanged).
// The only host round trip is the G x G table of member counts every rank needs to size its receives.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a torch process that is the library torch
// already loaded, otherwise the system one); the communicator is created by arcte_cuda_comm_init (one
// process per GPU: the 128-byte id comes from arcte_cuda_comm_unique_id on rank 0 and is distributed by the
// caller) or arcte_cuda_comm_init_all (one process driving several GPUs).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <vector>

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
int assemble_parts(arcte_cuda_ctx *c, int n_parts, const SegPart *parts, int64_t row_lo, int64_t row_hi);

// ---- NCCL entry points, resolved once ------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *(*GetLastError)(ncclComm_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static int load_nccl()
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return ARCTE_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(std::string("multi-GPU exchange: cannot load libnccl.so.2 (") + dlerror() + ")");
        return ARCTE_E_ARG;
    }
    NcclApi a;
    a.handle = h;
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) ok = false;
        return p;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    if (!ok) {
        set_error("multi-GPU exchange: libnccl.so.2 lacks a required entry point");
        return ARCTE_E_ARG;
    }
    g_nccl = a;
    return ARCTE_OK;
}

#define ARCTE_NCCL_TRY(expr)                                                                            \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            ::arcte::set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r) + " (" + __FILE__ + \
                               ":" + std::to_string(__LINE__) + ")");                                   \
            return ARCTE_E_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

// ---- splitting the communities by destination row block -------------------------------------------
__device__ __forceinline__ int dest_of(int64_t row, int64_t n, int G)
{
    int d = (int)((row * G) / n);
    // block d holds rows [n d / G, n (d+1) / G): the quotient above can be one too small
    while (d + 1 < G && (n * (int64_t)(d + 1)) / G <= row) ++d;
    while (d > 0 && (n * (int64_t)d) / G > row) --d;
    return d;
}

// cnt[d * S + k] = members of segment k whose row lies in block d.  One warp per segment.
__global__ void __launch_bounds__(256)
k_part_count(int64_t S, const int32_t *__restrict__ seg_count, const int64_t *__restrict__ seg_offset,
             const int32_t *__restrict__ members, int64_t n, int G, int32_t *__restrict__ cnt)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= S) return;
    const int lane = threadIdx.x & 31;
    const int32_t m = seg_count[k];
    int mine = 0;   // lane d counts destination d
    if (m > 0) {
        const int64_t src = seg_offset[k];
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            const int d = i < m ? dest_of(members[src + i], n, G) : -1;
            for (int t = 0; t < G; ++t) {
                const int c = __popc(__ballot_sync(kFull, d == t));
                if (lane == t) mine += c;
            }
        }
    }
    if (lane < G) cnt[(int64_t)lane * S + k] = mine;
}

// Writes the members of segment k that go to block d at off[d * S + k] of `out`, order kept.
__global__ void __launch_bounds__(256)
k_part_scatter(int64_t S, const int32_t *__restrict__ seg_count, const int64_t *__restrict__ seg_offset,
               const int32_t *__restrict__ members, int64_t n, int G, const int64_t *__restrict__ off,
               int32_t *__restrict__ out)
{
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= S) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int32_t m = seg_count[k];
    if (m <= 0) return;
    const int64_t src = seg_offset[k];
    int64_t cur = lane < G ? off[(int64_t)lane * S + k] : 0;   // lane d holds the write cursor of destination d
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        int32_t x = 0;
        int d = -1;
        if (i < m) {
            x = members[src + i];
            d = dest_of(x, n, G);
        }
        for (int t = 0; t < G; ++t) {
            const unsigned mask = __ballot_sync(kFull, d == t);
            const int64_t base = __shfl_sync(kFull, cur, t);
            if (d == t) out[base + __popc(mask & lt)] = x;
            if (lane == t) cur += __popc(mask);
        }
    }
}

__global__ void k_shard_seeds(int64_t n_work, int shard_rank, int shard_count, const int32_t *__restrict__ seeds,
                              int32_t *__restrict__ work_seed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_work) work_seed[i] = seeds[shard_rank + i * shard_count];   // arcte.py:19-23
}

static inline unsigned grid_for(int64_t items, int block) { return (unsigned)((items + block - 1) / block); }
static inline int64_t shard_size(int64_t n_seeds, int rank, int world)
{
    return n_seeds > rank ? (n_seeds - rank + world - 1) / world : 0;
}

int exchange_assemble(arcte_cuda_ctx *c)
{
    if (!c->comm) { set_error("exchange: no communicator (arcte_cuda_comm_init / _comm_init_all)"); return ARCTE_E_ARG; }
    if (!c->have_segments) { set_error("exchange: call extract first"); return ARCTE_E_ARG; }
    const int G = c->comm_world, r = c->comm_rank;
    if (c->shard_count != G || c->shard_rank != r) {
        set_error("exchange: the last extract must have walked shard comm_rank of comm_world");
        return ARCTE_E_ARG;
    }
    ncclComm_t comm = (ncclComm_t)c->comm;
    cudaStream_t st = c->stream;
    int64_t *launches = &c->stats.launches;
    const int64_t n = c->n, S = c->n_segments;
    const int64_t S1 = S > 0 ? S : 1;
    ARCTE_CUDA_TRY(cudaEventRecord(c->xev0, st));

    // ---- 1. split by destination ----
    DevBuf &cnt = c->xbuf[0], &off = c->xbuf[1], &sendbuf = c->xbuf[2], &tot = c->xbuf[3], &alltot = c->xbuf[4];
    ARCTE_TRY(dev_reserve(cnt, sizeof(int32_t) * (size_t)G * (size_t)S1));
    ARCTE_TRY(dev_reserve(off, sizeof(int64_t) * ((size_t)G * (size_t)S1 + 1)));
    ARCTE_TRY(dev_reserve(tot, sizeof(int64_t) * (size_t)(G + 1)));
    ARCTE_TRY(dev_reserve(alltot, sizeof(int64_t) * (size_t)G * (size_t)(G + 1)));
    if (S > 0) {
        k_part_count<<<grid_for(S * 32, 256), 256, 0, st>>>(S, c->seg_count.as<int32_t>(), c->seg_offset.as<int64_t>(),
                                                            c->members.as<int32_t>(), n, G, cnt.as<int32_t>());
        ++*launches;
        ARCTE_TRY(exclusive_scan_i32(cnt.as<int32_t>(), off.as<int64_t>(), (int64_t)G * S, c->scratch[6], st, launches));
    } else {
        ARCTE_CUDA_TRY(cudaMemsetAsync(off.p, 0, sizeof(int64_t) * ((size_t)G * (size_t)S1 + 1), st));
    }
    // start offset of every destination (and the total) -> all ranks
    for (int d = 0; d <= G; ++d)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(tot.as<int64_t>() + d, off.as<int64_t>() + (int64_t)d * S, sizeof(int64_t),
                                       cudaMemcpyDeviceToDevice, st));
    ARCTE_NCCL_TRY(g_nccl.AllGather(tot.p, alltot.p, (size_t)(G + 1), ncclInt64, comm, st));
    std::vector<int64_t> h_all((size_t)G * (size_t)(G + 1));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(h_all.data(), alltot.p, sizeof(int64_t) * h_all.size(), cudaMemcpyDeviceToHost, st));
    const int64_t M = c->n_members > 0 ? c->n_members : 1;
    ARCTE_TRY(dev_reserve(sendbuf, sizeof(int32_t) * (size_t)M));
    if (S > 0) {
        k_part_scatter<<<grid_for(S * 32, 256), 256, 0, st>>>(S, c->seg_count.as<int32_t>(), c->seg_offset.as<int64_t>(),
                                                              c->members.as<int32_t>(), n, G, off.as<int64_t>(),
                                                              sendbuf.as<int32_t>());
        ++*launches;
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(st));   // the count table is on the host now
    auto sent = [&](int from, int to) { return h_all[(size_t)from * (G + 1) + to + 1] - h_all[(size_t)from * (G + 1) + to]; };

    // ---- 2. all-to-all ----
    std::vector<SegPart> parts((size_t)G);
    for (int p = 0; p < G; ++p) {
        if (p == r) continue;
        const int64_t Sp = shard_size(c->n_seeds, p, G), Mp = sent(p, r);
        ARCTE_TRY(dev_reserve(c->peer_stage[(size_t)p * 4 + 0], sizeof(int32_t) * (size_t)(Sp > 0 ? Sp : 1)));   // seeds
        ARCTE_TRY(dev_reserve(c->peer_stage[(size_t)p * 4 + 1], sizeof(int32_t) * (size_t)(Sp > 0 ? Sp : 1)));   // counts
        ARCTE_TRY(dev_reserve(c->peer_stage[(size_t)p * 4 + 2], sizeof(int64_t) * (size_t)(Sp + 1)));            // offsets
        ARCTE_TRY(dev_reserve(c->peer_stage[(size_t)p * 4 + 3], sizeof(int32_t) * (size_t)(Mp > 0 ? Mp : 1)));   // members
    }
    ARCTE_NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < G; ++p) {
        if (p == r) continue;
        const int64_t Sp = shard_size(c->n_seeds, p, G), Mp = sent(p, r), Mo = sent(r, p);
        if (S > 0) ARCTE_NCCL_TRY(g_nccl.Send(cnt.as<int32_t>() + (int64_t)p * S, (size_t)S, ncclInt32, p, comm, st));
        if (Mo > 0) ARCTE_NCCL_TRY(g_nccl.Send(sendbuf.as<int32_t>() + h_all[(size_t)r * (G + 1) + p], (size_t)Mo, ncclInt32, p, comm, st));
        if (Sp > 0) ARCTE_NCCL_TRY(g_nccl.Recv(c->peer_stage[(size_t)p * 4 + 1].p, (size_t)Sp, ncclInt32, p, comm, st));
        if (Mp > 0) ARCTE_NCCL_TRY(g_nccl.Recv(c->peer_stage[(size_t)p * 4 + 3].p, (size_t)Mp, ncclInt32, p, comm, st));
    }
    ARCTE_NCCL_TRY(g_nccl.GroupEnd());

    // ---- 3. the G parts of this rank's row block ----
    for (int p = 0; p < G; ++p) {
        if (p == r) {
            // own members of own block: segment k starts at off[r * S + k] of the send buffer
            parts[(size_t)p] = SegPart{S, c->work_seed.as<int32_t>(), cnt.as<int32_t>() + (int64_t)r * S,
                                       off.as<int64_t>() + (int64_t)r * S, sendbuf.as<int32_t>()};
            continue;
        }
        const int64_t Sp = shard_size(c->n_seeds, p, G);
        int32_t *seedp = c->peer_stage[(size_t)p * 4 + 0].as<int32_t>();
        int32_t *cntp = c->peer_stage[(size_t)p * 4 + 1].as<int32_t>();
        int64_t *offp = c->peer_stage[(size_t)p * 4 + 2].as<int64_t>();
        if (Sp > 0) {
            k_shard_seeds<<<grid_for(Sp, 256), 256, 0, st>>>(Sp, p, G, c->seeds.as<int32_t>(), seedp);
            ++*launches;
            ARCTE_TRY(exclusive_scan_i32(cntp, offp, Sp, c->scratch[6], st, launches));
        }
        parts[(size_t)p] = SegPart{Sp, seedp, cntp, offp, c->peer_stage[(size_t)p * 4 + 3].as<int32_t>()};
    }
    ARCTE_CUDA_TRY(cudaEventRecord(c->xev1, st));
    const int64_t lo = (n * r) / G, hi = (n * (int64_t)(r + 1)) / G;
    ARCTE_TRY(assemble_parts(c, G, parts.data(), lo, hi));   // synchronises the stream
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->xev0, c->xev1));
    c->stats.ms_exchange = ms;
    return ARCTE_OK;
}

void comm_free(arcte_cuda_ctx *c)
{
    if (c->comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    c->comm_world = c->comm_rank = 0;
}

}  // namespace arcte

using namespace arcte;

extern "C" {

int arcte_cuda_comm_unique_id(void *id_out)
{
    if (!id_out) { set_error("comm_unique_id: null pointer"); return ARCTE_E_ARG; }
    ARCTE_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI passes the id as 128 bytes");
    ncclUniqueId id;
    ARCTE_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return ARCTE_OK;
}

int arcte_cuda_comm_init(arcte_cuda_ctx *c, int world, int rank, const void *id_in)
{
    if (!c || !id_in || world < 1 || world > 16 || rank < 0 || rank >= world) {
        set_error("comm_init: bad arguments (world must be 1..16)");
        return ARCTE_E_ARG;
    }
    ARCTE_CUDA_TRY(cudaSetDevice(c->device));
    ARCTE_TRY(load_nccl());
    comm_free(c);
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof(id));
    ncclComm_t comm = nullptr;
    ARCTE_NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
    c->comm = comm;
    c->comm_world = world;
    c->comm_rank = rank;
    return ARCTE_OK;
}

int arcte_cuda_comm_init_all(arcte_cuda_ctx *const *ctxs, int n)
{
    if (!ctxs || n < 1 || n > 16) { set_error("comm_init_all: bad arguments (1..16 contexts)"); return ARCTE_E_ARG; }
    ARCTE_TRY(load_nccl());
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) { set_error("comm_init_all: null context"); return ARCTE_E_ARG; }
        devs[(size_t)i] = ctxs[i]->device;
        comm_free(ctxs[i]);
    }
    std::vector<ncclComm_t> comms((size_t)n);
    ARCTE_NCCL_TRY(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->comm = comms[(size_t)i];
        ctxs[i]->comm_world = n;
        ctxs[i]->comm_rank = i;
    }
    return ARCTE_OK;
}

int arcte_cuda_comm_info(arcte_cuda_ctx *c, int *world, int *rank, int *nccl_version)
{
    if (!c) { set_error("null context"); return ARCTE_E_ARG; }
    if (world) *world = c->comm ? c->comm_world : 0;
    if (rank) *rank = c->comm ? c->comm_rank : 0;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    return ARCTE_OK;
}

int arcte_cuda_exchange_assemble(arcte_cuda_ctx *c, int64_t *nnz_out)
{
    if (!c) { set_error("null context"); return ARCTE_E_ARG; }
    ARCTE_CUDA_TRY(cudaSetDevice(c->device));
    ARCTE_TRY(exchange_assemble(c));
    if (nnz_out) *nnz_out = c->out_nnz;
    return ARCTE_OK;
}

}  // extern "C"
