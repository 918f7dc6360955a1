// api.cu -- the extern "C" boundary declared in include/arcte_cuda.h.
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace arcte {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

int dev_reserve(DevBuf &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return ARCTE_OK;
    if (b.borrowed) {
        set_error("internal: a view into the graph arena cannot grow");
        return ARCTE_E_ARG;
    }
    if (b.p) {
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
    static const bool dbg = getenv("ARCTE_CUDA_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (dbg && bytes >= (size_t(64) << 20))
        fprintf(stderr, "[arcte] cudaMalloc %.2f GB: %.1f ms\n", 1e-9 * (double)bytes,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
        b.p = nullptr;
        return e == cudaErrorMemoryAllocation ? ARCTE_E_NOMEM : ARCTE_E_CUDA;
    }
    b.bytes = bytes;
    return ARCTE_OK;
}

void dev_free(DevBuf &b)
{
    if (b.p && !b.borrowed) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    b.borrowed = false;
}

// implemented in the other translation units
int build_transition(arcte_cuda_ctx *c);
int select_seeds(arcte_cuda_ctx *c);
int count_columns(arcte_cuda_ctx *c);
int compute_eps_effective(arcte_cuda_ctx *c, double epsilon, const int32_t *dev_seeds, int64_t n_seeds,
                          double *dev_eps_out);
int extract_shard(arcte_cuda_ctx *c, int rule, double rho, double epsilon, int shard_rank, int shard_count,
                  const double *host_eps_override);
int push_single(arcte_cuda_ctx *c, int rule, int64_t seed, double rho, double eps_eff, double *host_s,
                double *host_r, int64_t *n_push);
struct SegPart {
    int64_t n_segments;
    const int32_t *seg_seed;
    const int32_t *seg_count;
    const int64_t *seg_offset;
    const int32_t *members;
};
int assemble_parts(arcte_cuda_ctx *c, int n_parts, const SegPart *parts, int64_t row_lo, int64_t row_hi);

}  // namespace arcte

namespace arcte {
__global__ void k_iota(int64_t n, int32_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)i;
}
__global__ void k_cent_finish(int64_t n, const unsigned long long *acc, double inv_scale, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __dmul_rn(__ull2double_rn(acc[i]), inv_scale);
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

const char *arcte_cuda_last_error(void) { return g_last_error.c_str(); }

int arcte_cuda_device_count(int *count)
{
    if (!count) { set_error("device_count: null pointer"); return ARCTE_E_ARG; }
    *count = 0;
    ARCTE_CUDA_TRY(cudaGetDeviceCount(count));
    return ARCTE_OK;
}

int arcte_cuda_create(arcte_cuda_ctx **out, int device_id)
{
    if (!out) { set_error("create: null out pointer"); return ARCTE_E_ARG; }
    *out = nullptr;
    int count = 0;
    ARCTE_CUDA_TRY(cudaGetDeviceCount(&count));
    if (device_id < 0 || device_id >= count) {
        set_error("create: device " + std::to_string(device_id) + " not present (" + std::to_string(count) +
                  " CUDA devices)");
        return ARCTE_E_ARG;
    }
    ARCTE_CUDA_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    ARCTE_CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) {
        set_error(std::string("create: this library is built for sm_100a only; device is ") + prop.name +
                  " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return ARCTE_E_ARG;
    }
    arcte_cuda_ctx *c = new arcte_cuda_ctx();
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    ARCTE_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->ev0));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->ev1));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->tm0));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->tm1));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->pk0));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->pk1));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->xev0));
    ARCTE_CUDA_TRY(cudaEventCreate(&c->xev1));
    {   // Experiment switch, off by default: ARCTE_CUDA_L2_PERSIST_MB=<n> sets n MB of L2 aside for persisting
        // accesses to the graph arena (upload_structure).  Measured on the bench shape it is a loss: the walks'
        // own state needs the capacity more (profiles/r2_engines.md, item 7).
        const char *env = getenv("ARCTE_CUDA_L2_PERSIST_MB");
        size_t want = env ? (size_t)atol(env) << 20 : 0;
        if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            c->l2_persist_bytes = want;
            c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        } else {
            (void)cudaGetLastError();
        }
        if (getenv("ARCTE_CUDA_DEBUG"))
            fprintf(stderr, "[arcte] L2 %d MB, persisting max %d MB (set %zu MB), window max %d MB\n", prop.l2CacheSize >> 20,
                    prop.persistingL2CacheMaxSize >> 20, c->l2_persist_bytes >> 20, prop.accessPolicyMaxWindowSize >> 20);
    }
    {   // Experiment switch: ARCTE_CUDA_L2_FETCH=32|64|128 sets the L2 fetch granularity (bytes DRAM delivers per missing sector)
        const char *env = getenv("ARCTE_CUDA_L2_FETCH");
        size_t cur = 0;
        if (env && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atol(env)) != cudaSuccess) (void)cudaGetLastError();
        if (getenv("ARCTE_CUDA_DEBUG") && cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity) == cudaSuccess)
            fprintf(stderr, "[arcte] L2 fetch granularity %zu bytes\n", cur);
    }
    *out = c;
    return ARCTE_OK;
}

void arcte_cuda_destroy(arcte_cuda_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->graph_arena, &c->indptr, &c->indices, &c->adj, &c->w, &c->d_out, &c->d_in, &c->colcnt, &c->node_info, &c->edge_wd, &c->edge_din, &c->seeds,
                      &c->work_seed, &c->work_eps, &c->seg_count, &c->seg_offset, &c->members, &c->retry_list,
                      &c->slots.sr, &c->slots.touched, &c->slots.queue, &c->slots.cmap, &c->slots.cepoch, &c->slots.frontier, &c->slots.fval, &c->counters, &c->out_indptr,
                      &c->out_indices, &c->out_data};
    for (DevBuf *b : bufs) dev_free(*b);
    for (DevBuf &b : c->scratch) dev_free(b);
    DevBuf *fbufs[] = {&c->fs_indptr, &c->fs_indices, &c->fs_data, &c->fg_indptr, &c->fg_indices, &c->fg_data, &c->fg_rows,
                       &c->fo_indptr[0], &c->fo_indptr[1], &c->fo_indices[0], &c->fo_indices[1], &c->fo_data[0], &c->fo_data[1]};
    for (DevBuf *b : fbufs) dev_free(*b);
    for (DevBuf &b : c->peer_stage) dev_free(b);
    DevBuf *pbufs[] = {&c->bpool.tbl, &c->bpool.stage, &c->bpool.clean, &c->bpool.queue, &c->row_w, &c->to_walk, &c->from_walk,
                       &c->walk_info, &c->walk_row_w, &c->walk_indices, &c->work_seed_w, &c->work_order};
    for (DevBuf *b : pbufs) dev_free(*b);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->tm0);
    cudaEventDestroy(c->tm1);
    cudaEventDestroy(c->pk0);
    cudaEventDestroy(c->pk1);
    cudaEventDestroy(c->xev0);
    cudaEventDestroy(c->xev1);
    comm_free(c);
    for (DevBuf &b : c->xbuf) dev_free(b);
    dev_free(c->l2_flush);
    free_ring(c);
    cudaStreamDestroy(c->stream);
    delete c;
}

int arcte_cuda_configure(arcte_cuda_ctx *c, int warps_per_sm, int64_t queue_capacity, int mem_percent,
                         int64_t member_capacity)
{
    CHECK_CTX(c);
    if (warps_per_sm < 0 || warps_per_sm > 64 || queue_capacity < 0 || mem_percent < 0 || mem_percent > 95 ||
        member_capacity < 0) {
        set_error("configure: argument out of range");
        return ARCTE_E_ARG;
    }
    c->warps_per_sm = warps_per_sm;
    c->queue_cap_cfg = queue_capacity;
    c->mem_percent = mem_percent;
    c->member_cap_cfg = member_capacity;
    dev_free(c->members);  // re-created with the new capacity at the next extract
    c->member_cap = 0;
    return ARCTE_OK;
}

int arcte_cuda_set_schedule(arcte_cuda_ctx *c, int schedule, int heavy_permille, int heavy_threads,
                            int heavy_ctas_per_sm, int light_threads, int light_ctas_per_sm)
{
    CHECK_CTX(c);
    auto threads_ok = [](int t) { return t <= 0 || t == 32 || t == 64 || t == 128 || t == 256 || t == 512 || t == 1024; };
    if ((schedule != ARCTE_SCHEDULE_FIFO && schedule != ARCTE_SCHEDULE_FRONTIER) || heavy_permille > 1000 ||
        !threads_ok(heavy_threads) || !threads_ok(light_threads) || heavy_ctas_per_sm > 32 || light_ctas_per_sm > 32) {
        set_error("set_schedule: argument out of range");
        return ARCTE_E_ARG;
    }
    c->schedule = schedule;
    c->fr_heavy_permille = heavy_permille >= 0 ? heavy_permille : -1;
    c->fr_heavy_threads = heavy_threads > 0 ? heavy_threads : 0;
    c->fr_heavy_ctas = heavy_ctas_per_sm > 0 ? heavy_ctas_per_sm : 0;
    c->fr_light_threads = light_threads > 0 ? light_threads : 0;
    c->fr_light_ctas = light_ctas_per_sm > 0 ? light_ctas_per_sm : 0;
    return ARCTE_OK;
}

int arcte_cuda_set_engine(arcte_cuda_ctx *c, int engine, int64_t table_capacity)
{
    CHECK_CTX(c);
    if (engine < ARCTE_ENGINE_AUTO || engine > ARCTE_ENGINE_FIFO_COMPACT || table_capacity < 0 ||
        table_capacity > (int64_t(1) << 26)) {
        set_error("set_engine: argument out of range");
        return ARCTE_E_ARG;
    }
    c->engine = engine;
    c->tbl_cap_cfg = table_capacity;
    return ARCTE_OK;
}

int arcte_cuda_set_graph(arcte_cuda_ctx *c, int64_t n, int64_t nnz, const int64_t *host_indptr,
                         const int32_t *host_indices, const double *host_data);

static int upload_structure(arcte_cuda_ctx *c, int64_t n, int64_t nnz, const int64_t *host_indptr,
                            const int32_t *host_indices)
{
    if (n <= 0 || nnz < 0 || !host_indptr || (nnz > 0 && !host_indices)) {
        set_error("graph upload: bad arguments");
        return ARCTE_E_ARG;
    }
    if (n >= (int64_t(1) << 30)) { set_error("graph upload: n must be < 2^30"); return ARCTE_E_ARG; }
    if (nnz >= (int64_t(1) << 32)) { set_error("graph upload: nnz must be < 2^32"); return ARCTE_E_ARG; }
    if (host_indptr[0] != 0 || host_indptr[n] != nnz) {
        set_error("graph upload: indptr[0] must be 0 and indptr[n] must equal nnz");
        return ARCTE_E_ARG;
    }
    c->have_graph = c->have_transition = c->have_segments = c->have_features = false;
    c->row_w_valid = false;
    c->walk_labels_valid = false;
    if (n != c->n) {
        // slot geometry depends on n: drop the pool (re-created lazily)
        dev_free(c->slots.sr);
        dev_free(c->slots.touched);
        dev_free(c->slots.queue);
        dev_free(c->slots.cmap);
        dev_free(c->slots.cepoch);
        dev_free(c->slots.frontier);
        dev_free(c->slots.fval);
        c->slots = SlotPool();
    }
    c->n = n;
    c->nnz = nnz;
    c->stats = arcte_cuda_stats();
    ARCTE_TRY(dev_reserve(c->indptr, sizeof(int64_t) * (size_t)(n + 1)));
    {   // graph arena: node records, row weights, column indices, transition weights, in this order
        auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
        const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
        const size_t b_info = up(sizeof(NodeInfo) * (size_t)n), b_rw = up(sizeof(double) * (size_t)n);
        const size_t b_idx = up(sizeof(int32_t) * nz), b_w = up(sizeof(double) * nz);
        DevBuf *views[] = {&c->node_info, &c->row_w, &c->indices, &c->w};
        for (DevBuf *v : views) dev_free(*v);
        ARCTE_TRY(dev_reserve(c->graph_arena, b_info + b_rw + b_idx + b_w));
        char *base = c->graph_arena.as<char>();
        const size_t sizes[] = {b_info, b_rw, b_idx, b_w};
        for (int i = 0; i < 4; ++i) {
            views[i]->p = base;
            views[i]->bytes = sizes[i];
            views[i]->borrowed = true;
            base += sizes[i];
        }
        // L2 persistence: the walks re-read node records (random gathers), row weights and column indices
        // millions of times while their own state streams through L2 once
        if (c->l2_persist_bytes > 0) {
            size_t win = b_info + b_rw + b_idx;
            if (win > c->l2_window_max) win = c->l2_window_max;
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = c->graph_arena.p;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = win <= c->l2_persist_bytes ? 1.0f : (float)((double)c->l2_persist_bytes / (double)win);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            ARCTE_CUDA_TRY(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
    }
    ARCTE_TRY(copy_from_host(c, c->indptr.p, host_indptr, sizeof(int64_t) * (size_t)(n + 1)));
    if (nnz > 0) ARCTE_TRY(copy_from_host(c, c->indices.p, host_indices, sizeof(int32_t) * (size_t)nnz));
    return ARCTE_OK;
}

int arcte_cuda_set_transition(arcte_cuda_ctx *c, int64_t n, int64_t nnz, const int64_t *host_indptr,
                              const int32_t *host_indices, const double *host_w, const double *host_d_out,
                              const double *host_d_in)
{
    CHECK_CTX(c);
    if (!host_d_out || !host_d_in || (nnz > 0 && !host_w)) { set_error("set_transition: null array"); return ARCTE_E_ARG; }
    ARCTE_TRY(upload_structure(c, n, nnz, host_indptr, host_indices));
    ARCTE_TRY(dev_reserve(c->w, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    ARCTE_TRY(dev_reserve(c->d_out, sizeof(double) * (size_t)n));
    ARCTE_TRY(dev_reserve(c->d_in, sizeof(double) * (size_t)n));
    if (nnz > 0)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->w.p, host_w, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->d_out.p, host_d_out, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->d_in.p, host_d_in, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    c->have_graph = true;       // structure resident (adjacency weights are not: K1 cannot be re-run)
    dev_free(c->adj);
    ARCTE_TRY(count_columns(c));
    c->have_transition = true;
    return select_seeds(c);
}

int arcte_cuda_set_seeds(arcte_cuda_ctx *c, int64_t n_seeds, const int64_t *host_seeds)
{
    CHECK_CTX(c);
    if (!c->have_transition) { set_error("set_seeds: no graph"); return ARCTE_E_ARG; }
    if (n_seeds < 0 || n_seeds > c->n || (n_seeds > 0 && !host_seeds)) { set_error("set_seeds: bad arguments"); return ARCTE_E_ARG; }
    std::vector<int32_t> tmp((size_t)n_seeds);
    for (int64_t i = 0; i < n_seeds; ++i) {
        if (host_seeds[i] < 0 || host_seeds[i] >= c->n) { set_error("set_seeds: seed out of range"); return ARCTE_E_ARG; }
        tmp[(size_t)i] = (int32_t)host_seeds[i];
    }
    if (n_seeds > 0) {
        ARCTE_CUDA_TRY(cudaMemcpyAsync(c->seeds.p, tmp.data(), sizeof(int32_t) * (size_t)n_seeds, cudaMemcpyHostToDevice, c->stream));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    c->n_seeds = n_seeds;
    c->stats.n_seeds_total = n_seeds;
    c->have_segments = c->have_features = false;
    return ARCTE_OK;
}

int arcte_cuda_set_graph(arcte_cuda_ctx *c, int64_t n, int64_t nnz, const int64_t *host_indptr,
                         const int32_t *host_indices, const double *host_data)
{
    CHECK_CTX(c);
    if (nnz > 0 && !host_data) { set_error("set_graph: null data"); return ARCTE_E_ARG; }
    ARCTE_TRY(upload_structure(c, n, nnz, host_indptr, host_indices));
    ARCTE_TRY(dev_reserve(c->adj, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    if (nnz > 0) ARCTE_TRY(copy_from_host(c, c->adj.p, host_data, sizeof(double) * (size_t)nnz));
    c->have_graph = true;
    return build_transition(c);
}

int arcte_cuda_build_transition(arcte_cuda_ctx *c)
{
    CHECK_CTX(c);
    if (!c->have_graph || !c->adj.p) { set_error("build_transition: adjacency weights not resident"); return ARCTE_E_ARG; }
    return build_transition(c);
}

int arcte_cuda_get_transition(arcte_cuda_ctx *c, double *host_w, double *host_d_out, double *host_d_in)
{
    CHECK_CTX(c);
    if (!c->have_transition) { set_error("get_transition: nothing built"); return ARCTE_E_ARG; }
    if (host_w && c->nnz > 0)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(host_w, c->w.p, sizeof(double) * (size_t)c->nnz, cudaMemcpyDeviceToHost, c->stream));
    if (host_d_out)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(host_d_out, c->d_out.p, sizeof(double) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    if (host_d_in)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(host_d_in, c->d_in.p, sizeof(double) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_get_seed_count(arcte_cuda_ctx *c, int64_t *n_seeds)
{
    CHECK_CTX(c);
    if (!c->have_transition || !n_seeds) { set_error("get_seed_count: no graph"); return ARCTE_E_ARG; }
    *n_seeds = c->n_seeds;
    return ARCTE_OK;
}

int arcte_cuda_get_seeds(arcte_cuda_ctx *c, int64_t *host_seeds)
{
    CHECK_CTX(c);
    if (!c->have_transition || !host_seeds) { set_error("get_seeds: no graph"); return ARCTE_E_ARG; }
    std::vector<int32_t> tmp((size_t)c->n_seeds);
    if (c->n_seeds > 0) {
        ARCTE_CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->seeds.p, sizeof(int32_t) * (size_t)c->n_seeds,
                                       cudaMemcpyDeviceToHost, c->stream));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    for (int64_t i = 0; i < c->n_seeds; ++i) host_seeds[i] = tmp[(size_t)i];
    return ARCTE_OK;
}

int arcte_cuda_epsilon_effective(arcte_cuda_ctx *c, double epsilon, int64_t n_seeds, const int64_t *host_seeds,
                                 double *host_eps_out)
{
    CHECK_CTX(c);
    if (!c->have_transition) { set_error("epsilon_effective: no graph"); return ARCTE_E_ARG; }
    if (n_seeds <= 0) return ARCTE_OK;
    std::vector<int32_t> tmp((size_t)n_seeds);
    for (int64_t i = 0; i < n_seeds; ++i) {
        if (host_seeds[i] < 0 || host_seeds[i] >= c->n) { set_error("epsilon_effective: seed out of range"); return ARCTE_E_ARG; }
        tmp[(size_t)i] = (int32_t)host_seeds[i];
    }
    ARCTE_TRY(dev_reserve(c->scratch[0], sizeof(int32_t) * (size_t)n_seeds));
    ARCTE_TRY(dev_reserve(c->scratch[2], sizeof(double) * (size_t)n_seeds));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(c->scratch[0].p, tmp.data(), sizeof(int32_t) * (size_t)n_seeds,
                                   cudaMemcpyHostToDevice, c->stream));
    ARCTE_TRY(compute_eps_effective(c, epsilon, c->scratch[0].as<int32_t>(), n_seeds, c->scratch[2].as<double>()));
    ARCTE_CUDA_TRY(cudaMemcpyAsync(host_eps_out, c->scratch[2].p, sizeof(double) * (size_t)n_seeds,
                                   cudaMemcpyDeviceToHost, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_push(arcte_cuda_ctx *c, int rule, int64_t seed, double rho, double eps_eff, double *host_s,
                    double *host_r, int64_t *n_push)
{
    CHECK_CTX(c);
    if (!host_s || !host_r) { set_error("push: null output"); return ARCTE_E_ARG; }
    return push_single(c, rule, seed, rho, eps_eff, host_s, host_r, n_push);
}

int arcte_cuda_extract(arcte_cuda_ctx *c, int rule, double rho, double epsilon, int shard_rank, int shard_count,
                       const double *host_eps_override, int64_t *n_segments, int64_t *n_members)
{
    CHECK_CTX(c);
    if (rule < 0 || rule > 2) { set_error("extract: unknown rule"); return ARCTE_E_ARG; }
    ARCTE_TRY(extract_shard(c, rule, rho, epsilon, shard_rank, shard_count, host_eps_override));
    if (n_segments) *n_segments = c->n_segments;
    if (n_members) *n_members = c->n_members;
    return ARCTE_OK;
}


/* The RCT centrality of arcte_and_centrality (embedding/arcte/cython_opt/arcte.pyx:125-241): every node is a
   seed (out_degree != 0 holds for all of them after transition.py:58), the walk uses the RAW epsilon (no
   epsilon-effective, arcte.pyx:164), and centrality[x] = sum over seeds of s[x] / d_in[x] (arcte.pyx:183-191). */
int arcte_cuda_centrality(arcte_cuda_ctx *c, double rho, double epsilon, double *host_centrality)
{
    CHECK_CTX(c);
    if (!c->have_transition || !host_centrality) { set_error("centrality: no graph / null output"); return ARCTE_E_ARG; }
    const int64_t n = c->n;
    const int64_t saved_seeds = c->n_seeds;
    DevBuf acc, eps;
    int rc = dev_reserve(acc, sizeof(unsigned long long) * (size_t)n);
    if (rc == ARCTE_OK) rc = dev_reserve(eps, sizeof(double) * (size_t)n);
    std::vector<double> h_eps((size_t)n, epsilon);
    if (rc == ARCTE_OK) {
        cudaMemsetAsync(acc.p, 0, sizeof(unsigned long long) * (size_t)n, c->stream);
        k_iota<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->seeds.as<int32_t>());   // every node, ascending
        c->n_seeds = n;
        c->centrality_acc = acc.as<unsigned long long>();
        rc = extract_shard(c, ARCTE_RULE_ABSORBING, rho, epsilon, 0, 1, h_eps.data());
        c->centrality_acc = nullptr;
    }
    if (rc == ARCTE_OK) {
        k_cent_finish<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, acc.as<unsigned long long>(), 1.0 / 274877906944.0,
                                                                         eps.as<double>());
        rc = copy_to_host(c, host_centrality, eps.p, sizeof(double) * (size_t)n);
    }
    dev_free(acc);
    dev_free(eps);
    // back to the seed list of arcte() (arcte.py:610-617); the segments of the centrality walk are not kept
    c->n_seeds = saved_seeds;
    c->have_segments = c->have_features = false;
    const int rc2 = select_seeds(c);
    return rc != ARCTE_OK ? rc : rc2;
}

int arcte_cuda_get_segments(arcte_cuda_ctx *c, int32_t *host_seg_seed, int32_t *host_seg_count,
                            int64_t *host_seg_offset, int32_t *host_members)
{
    CHECK_CTX(c);
    if (!c->have_segments) { set_error("get_segments: call extract first"); return ARCTE_E_ARG; }
    const size_t S = (size_t)c->n_segments;
    if (S > 0) {
        if (host_seg_seed) ARCTE_CUDA_TRY(cudaMemcpyAsync(host_seg_seed, c->work_seed.p, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, c->stream));
        if (host_seg_count) ARCTE_CUDA_TRY(cudaMemcpyAsync(host_seg_count, c->seg_count.p, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, c->stream));
        if (host_seg_offset) ARCTE_CUDA_TRY(cudaMemcpyAsync(host_seg_offset, c->seg_offset.p, sizeof(int64_t) * S, cudaMemcpyDeviceToHost, c->stream));
    }
    if (host_members && c->n_members > 0)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(host_members, c->members.p, sizeof(int32_t) * (size_t)c->n_members, cudaMemcpyDeviceToHost, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_segments_device(arcte_cuda_ctx *c, const int32_t **dev_seg_seed, const int32_t **dev_seg_count,
                               const int64_t **dev_seg_offset, const int32_t **dev_members)
{
    CHECK_CTX(c);
    if (!c->have_segments) { set_error("segments_device: call extract first"); return ARCTE_E_ARG; }
    if (dev_seg_seed) *dev_seg_seed = c->work_seed.as<int32_t>();
    if (dev_seg_count) *dev_seg_count = c->seg_count.as<int32_t>();
    if (dev_seg_offset) *dev_seg_offset = c->seg_offset.as<int64_t>();
    if (dev_members) *dev_members = c->members.as<int32_t>();
    return ARCTE_OK;
}

int arcte_cuda_export_segments(arcte_cuda_ctx *c, int32_t *dev_seg_seed, int32_t *dev_seg_count,
                               int64_t *dev_seg_offset, int32_t *dev_members)
{
    CHECK_CTX(c);
    if (!c->have_segments) { set_error("export_segments: call extract first"); return ARCTE_E_ARG; }
    const size_t S = (size_t)c->n_segments, M = (size_t)c->n_members;
    if (S > 0) {
        if (dev_seg_seed) ARCTE_CUDA_TRY(cudaMemcpyAsync(dev_seg_seed, c->work_seed.p, sizeof(int32_t) * S, cudaMemcpyDeviceToDevice, c->stream));
        if (dev_seg_count) ARCTE_CUDA_TRY(cudaMemcpyAsync(dev_seg_count, c->seg_count.p, sizeof(int32_t) * S, cudaMemcpyDeviceToDevice, c->stream));
        if (dev_seg_offset) ARCTE_CUDA_TRY(cudaMemcpyAsync(dev_seg_offset, c->seg_offset.p, sizeof(int64_t) * S, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (dev_members && M > 0)
        ARCTE_CUDA_TRY(cudaMemcpyAsync(dev_members, c->members.p, sizeof(int32_t) * M, cudaMemcpyDeviceToDevice, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

// If `p` lives on another device, stage `bytes` of it into `stage` on this context's
// device (peer DMA over NVLink) and return the local copy.
static int localise(arcte_cuda_ctx *c, const void *p, size_t bytes, DevBuf &stage, const void **out)
{
    *out = p;
    if (bytes == 0 || p == nullptr) return ARCTE_OK;
    cudaPointerAttributes at;
    ARCTE_CUDA_TRY(cudaPointerGetAttributes(&at, p));
    if (at.type == cudaMemoryTypeDevice && at.device == c->device) return ARCTE_OK;
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
        set_error("assemble: part pointers must be device pointers");
        return ARCTE_E_ARG;
    }
    ARCTE_TRY(dev_reserve(stage, bytes));
    {   // direct NVLink DMA needs peer access; "already enabled" / "not supported" are fine
        // (without it the copy is staged through the host by the driver)
        (void)cudaDeviceEnablePeerAccess(at.device, 0);
        (void)cudaGetLastError();
        (void)cudaSetDevice(at.device);
        (void)cudaDeviceEnablePeerAccess(c->device, 0);
        (void)cudaGetLastError();
        (void)cudaSetDevice(c->device);
    }
    ARCTE_CUDA_TRY(cudaMemcpyPeerAsync(stage.p, c->device, p, at.device, bytes, c->stream));
    *out = stage.p;
    return ARCTE_OK;
}

int arcte_cuda_assemble(arcte_cuda_ctx *c, int n_parts, const int64_t *part_n_segments,
                        const int64_t *part_n_members, const int32_t *const *dev_seg_seed,
                        const int32_t *const *dev_seg_count, const int64_t *const *dev_seg_offset,
                        const int32_t *const *dev_members, int64_t *nnz_out)
{
    if (!c) { set_error("null context"); return ARCTE_E_ARG; }
    return arcte_cuda_assemble_rows(c, n_parts, part_n_segments, part_n_members, dev_seg_seed, dev_seg_count,
                                    dev_seg_offset, dev_members, 0, c->n, nnz_out);
}

int arcte_cuda_assemble_rows(arcte_cuda_ctx *c, int n_parts, const int64_t *part_n_segments,
                             const int64_t *part_n_members, const int32_t *const *dev_seg_seed,
                             const int32_t *const *dev_seg_count, const int64_t *const *dev_seg_offset,
                             const int32_t *const *dev_members, int64_t row_lo, int64_t row_hi, int64_t *nnz_out)
{
    CHECK_CTX(c);
    std::vector<SegPart> parts;
    int rc = ARCTE_OK;
    if (n_parts == 0) {
        if (!c->have_segments) { set_error("assemble: call extract first"); return ARCTE_E_ARG; }
        parts.push_back(SegPart{c->n_segments, c->work_seed.as<int32_t>(), c->seg_count.as<int32_t>(),
                                c->seg_offset.as<int64_t>(), c->members.as<int32_t>()});
    } else {
        if (n_parts < 0 || !part_n_segments || !part_n_members || !dev_seg_seed || !dev_seg_count ||
            !dev_seg_offset || !dev_members) {
            set_error("assemble: bad part arrays");
            return ARCTE_E_ARG;
        }
        if (n_parts > 16) { set_error("assemble: at most 16 parts"); return ARCTE_E_ARG; }
        DevBuf *stage = c->peer_stage;
        for (int p = 0; p < n_parts && rc == ARCTE_OK; ++p) {
            const size_t S = (size_t)part_n_segments[p], M = (size_t)part_n_members[p];
            const void *a = nullptr, *b = nullptr, *d = nullptr, *e = nullptr;
            rc = localise(c, dev_seg_seed[p], S * sizeof(int32_t), stage[(size_t)p * 4 + 0], &a);
            if (rc == ARCTE_OK) rc = localise(c, dev_seg_count[p], S * sizeof(int32_t), stage[(size_t)p * 4 + 1], &b);
            if (rc == ARCTE_OK) rc = localise(c, dev_seg_offset[p], S * sizeof(int64_t), stage[(size_t)p * 4 + 2], &d);
            if (rc == ARCTE_OK) rc = localise(c, dev_members[p], M * sizeof(int32_t), stage[(size_t)p * 4 + 3], &e);
            parts.push_back(SegPart{part_n_segments[p], (const int32_t *)a, (const int32_t *)b,
                                    (const int64_t *)d, (const int32_t *)e});
        }
    }
    if (rc == ARCTE_OK) rc = assemble_parts(c, (int)parts.size(), parts.data(), row_lo, row_hi);
    if (rc != ARCTE_OK) return rc;
    if (nnz_out) *nnz_out = c->out_nnz;
    return ARCTE_OK;
}

int arcte_cuda_get_features(arcte_cuda_ctx *c, int64_t *host_indptr, int32_t *host_indices, double *host_data)
{
    CHECK_CTX(c);
    if (!c->have_features) { set_error("get_features: call assemble first"); return ARCTE_E_ARG; }
    return fetch_features(c, host_indptr, host_indices, host_data, 0, 0);
}

int arcte_cuda_fetch_features(arcte_cuda_ctx *c, int64_t *host_indptr, int32_t *host_indices, double *host_data,
                              int values_are_ones, int n_threads)
{
    CHECK_CTX(c);
    if (!c->have_features) { set_error("fetch_features: call assemble first"); return ARCTE_E_ARG; }
    return fetch_features(c, host_indptr, host_indices, host_data, values_are_ones, n_threads);
}

int arcte_cuda_fetch_features_to(arcte_cuda_ctx *c, int64_t pid, int64_t *dst_indptr, int32_t *dst_indices, double *dst_data,
                                 int values_are_ones, int n_threads)
{
    CHECK_CTX(c);
    if (!c->have_features) { set_error("fetch_features_to: call assemble first"); return ARCTE_E_ARG; }
    return fetch_features(c, dst_indptr, dst_indices, dst_data, values_are_ones, n_threads, (long)pid);
}

int arcte_cuda_host_write_to(int64_t pid, void *dst, const void *src, int64_t bytes)
{
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) { set_error("host_write_to: bad arguments"); return ARCTE_E_ARG; }
    return host_write_remote((long)pid, dst, src, (size_t)bytes);
}

int arcte_cuda_host_advise_huge(void *p, int64_t bytes)
{
    if (p && bytes > 0) host_advise_huge(p, (size_t)bytes);
    return ARCTE_OK;
}

int arcte_cuda_host_ones_alloc(int64_t count, void **out, int64_t *mapped_bytes)
{
    if (!out || !mapped_bytes || count < 0) { set_error("host_ones_alloc: bad arguments"); return ARCTE_E_ARG; }
    size_t mb = 0;
    const int rc = host_ones_alloc((size_t)count, out, &mb);
    *mapped_bytes = (int64_t)mb;
    return rc;
}

int arcte_cuda_host_ones_free(void *p, int64_t mapped_bytes)
{
    host_ones_free(p, (size_t)(mapped_bytes > 0 ? mapped_bytes : 0));
    return ARCTE_OK;
}

int arcte_cuda_host_alloc(void **out, int64_t bytes)
{
    if (!out || bytes <= 0) { set_error("host_alloc: bad arguments"); return ARCTE_E_ARG; }
    *out = nullptr;
    ARCTE_CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return ARCTE_OK;
}

int arcte_cuda_host_free(void *p)
{
    if (p) ARCTE_CUDA_TRY(cudaFreeHost(p));
    return ARCTE_OK;
}

int arcte_cuda_timer_start(arcte_cuda_ctx *c)
{
    CHECK_CTX(c);
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    ARCTE_CUDA_TRY(cudaEventRecord(c->tm0, c->stream));
    return ARCTE_OK;
}

int arcte_cuda_timer_stop(arcte_cuda_ctx *c, double *elapsed_ms)
{
    CHECK_CTX(c);
    ARCTE_CUDA_TRY(cudaEventRecord(c->tm1, c->stream));
    ARCTE_CUDA_TRY(cudaEventSynchronize(c->tm1));
    float ms = 0.f;
    ARCTE_CUDA_TRY(cudaEventElapsedTime(&ms, c->tm0, c->tm1));
    if (elapsed_ms) *elapsed_ms = ms;
    return ARCTE_OK;
}

int arcte_cuda_flush_l2(arcte_cuda_ctx *c)
{
    CHECK_CTX(c);
    const size_t bytes = size_t(512) << 20;
    ARCTE_TRY(dev_reserve(c->l2_flush, bytes));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->l2_flush.p, 0x5a, bytes, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ARCTE_OK;
}

int arcte_cuda_features_device(arcte_cuda_ctx *c, const int64_t **dev_indptr, const int32_t **dev_indices,
                               const double **dev_data, int64_t *n_rows, int64_t *nnz)
{
    CHECK_CTX(c);
    if (!c->have_features) { set_error("features_device: call assemble first"); return ARCTE_E_ARG; }
    if (dev_indptr) *dev_indptr = c->out_indptr.as<int64_t>();
    if (dev_indices) *dev_indices = c->out_indices.as<int32_t>();
    if (dev_data) *dev_data = c->out_data.as<double>();
    if (n_rows) *n_rows = c->out_rows;
    if (nnz) *nnz = c->out_nnz;
    return ARCTE_OK;
}

}  // extern "C"

namespace arcte {
// 64-bit content hash of an assembled block, additive over row blocks: every row start, column index and
// non-unit value is mixed with its GLOBAL position (splitmix64) and the terms are summed modulo 2^64, so the
// hashes of the row blocks of 1, 2, 4 or 8 GPUs add up to the same number exactly when the matrices agree.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256)
k_features_hash(int64_t rows, int64_t nnz, int64_t row_lo, int64_t nnz_lo, const int64_t *__restrict__ indptr,
                const int32_t *__restrict__ indices, const double *__restrict__ data, unsigned long long *__restrict__ out)
{
    unsigned long long h = 0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows + nnz; i += stride) {
        if (i < rows) {
            h += mix64(mix64(0x1000000000000000ull + (unsigned long long)(row_lo + i)) + (unsigned long long)(indptr[i] + nnz_lo));
        } else {
            const int64_t k = i - rows;
            const unsigned long long pos = (unsigned long long)(nnz_lo + k);
            h += mix64(mix64(0x2000000000000000ull + pos) + (unsigned long long)(uint32_t)indices[k]);
            const double v = data[k];
            if (v != 1.0) h += mix64(mix64(0x3000000000000000ull + pos) + (unsigned long long)__double_as_longlong(v));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(kFull, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}
}  // namespace arcte

extern "C" {

int arcte_cuda_features_hash(arcte_cuda_ctx *c, int64_t row_lo, int64_t nnz_lo, uint64_t *hash_out)
{
    CHECK_CTX(c);
    if (!c->have_features || !hash_out) { set_error("features_hash: call assemble first"); return ARCTE_E_ARG; }
    ARCTE_TRY(dev_reserve(c->scratch[4], sizeof(unsigned long long) * 2));
    ARCTE_CUDA_TRY(cudaMemsetAsync(c->scratch[4].p, 0, sizeof(unsigned long long), c->stream));
    const int64_t items = c->out_rows + c->out_nnz;
    unsigned grid = (unsigned)((items + 255) / 256);
    if (grid > (unsigned)c->sm_count * 16) grid = (unsigned)c->sm_count * 16;
    if (grid < 1) grid = 1;
    k_features_hash<<<grid, 256, 0, c->stream>>>(c->out_rows, c->out_nnz, row_lo, nnz_lo, c->out_indptr.as<int64_t>(),
                                                 c->out_indices.as<int32_t>(), c->out_data.as<double>(),
                                                 c->scratch[4].as<unsigned long long>());
    ++c->stats.launches;
    ARCTE_CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0ull;
    ARCTE_CUDA_TRY(cudaMemcpyAsync(&h, c->scratch[4].p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *hash_out = h;
    return ARCTE_OK;
}

int arcte_cuda_get_stats(arcte_cuda_ctx *c, arcte_cuda_stats *out)
{
    if (!c || !out) { set_error("get_stats: null argument"); return ARCTE_E_ARG; }
    *out = c->stats;
    return ARCTE_OK;
}

}  // extern "C"
