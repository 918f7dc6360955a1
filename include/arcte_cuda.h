/*
 * arcte_cuda.h -- C ABI of the B200-native ARCTE feature extractor.
 *
 * The reference (MKLab-ITI/reveal-graph-embedding) is pure Python and has no FFI
 * for this path; the boundary is the Python call
 *     arcte(adjacency_matrix, rho, epsilon, number_of_threads)
 *         reveal_graph_embedding/embedding/arcte/arcte.py:591
 * and, one level down, the raw-array hand-off its workers receive
 *     arcte_worker(iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon)
 *         reveal_graph_embedding/embedding/arcte/arcte.py:279-286
 * Every entry point below names the reference code it replaces.  A maintainer
 * binds them with ctypes; no torch types appear here.
 *
 * Conventions
 *   - every function returns 0 on success and a negative ARCTE_E_* code on
 *     failure; arcte_cuda_last_error() then returns a message (thread local).
 *   - "host" pointers are ordinary process memory (pinned or not); "dev"
 *     pointers are CUDA device pointers on the context's device.
 *   - one context drives one GPU and is not re-entrant; calls block until the
 *     result is complete (ctypes releases the GIL around them).
 *   - the adjacency CSR must be canonical (sorted column indices, no
 *     duplicates) with positive weights, which is what scipy hands the
 *     reference after csr_matrix()/sort_indices() (transition.py:52,65).
 */
#ifndef ARCTE_CUDA_H
#define ARCTE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct arcte_cuda_ctx arcte_cuda_ctx;

enum {
    ARCTE_OK = 0,
    ARCTE_E_CUDA = -1,      /* a CUDA runtime call failed                          */
    ARCTE_E_ARG = -2,       /* bad argument / call order                            */
    ARCTE_E_NOMEM = -3,     /* not enough device memory for the requested graph     */
    ARCTE_E_OVERFLOW = -4   /* a per-seed queue could not be grown any further      */
};

/* Push rules: which similarity.py driver + push.py rule the engine runs. */
enum {
    ARCTE_RULE_ABSORBING = 0, /* fast_approximate_cumulative_pagerank_difference, similarity.py:149;
                                 cumulative_pagerank_difference_limit_push, push.py:41   (arcte, arcte.py:591) */
    ARCTE_RULE_PAGERANK = 1,  /* fast_approximate_personalized_pagerank, similarity.py:11;
                                 pagerank_limit_push, push.py:4          (arcte_with_pagerank, arcte.py:491) */
    ARCTE_RULE_LAZY = 2       /* lazy_approximate_personalized_pagerank, similarity.py:66;
                                 pagerank_lazy_push, push.py:20     (arcte_with_lazy_pagerank, arcte.py:391) */
};

/* Walk schedules of the push engine (arcte_cuda_set_schedule). */
enum {
    ARCTE_SCHEDULE_FIFO = 0,     /* exact replay of the reference's FIFO (similarity.py:199-216): s, r, push
                                    counts and the thresholded support are bit-identical to the reference.  Default. */
    ARCTE_SCHEDULE_FRONTIER = 1  /* synchronous frontier rounds on 64-bit fixed-point state, one CTA per seed:
                                    every node over the threshold is pushed per round.  Same error bound
                                    0 <= (G - s)/d < eps (1-rho)/rho; support identical to the reference up to
                                    documented in-band ties; deterministic (independent of timing, sharding and
                                    launch geometry).  Absorbing rule (arcte) only. */
};

/* Engines of the exact FIFO schedule (arcte_cuda_set_engine).  All four replay the reference's queue
   discipline and arithmetic exactly (bit-identical s, r, push counts and supports); they differ in how
   the walk state is laid out and in how many queue entries one warp iteration takes. */
enum {
    ARCTE_ENGINE_AUTO = -1,          /* FIFO_COMPACT for graphs of 2^22 or more stored entries (all rules); below
                                        that, absorbing rule: BATCHED_DENSE for 4096 <= n <= 2^19, FIFO_DENSE
                                        otherwise; PageRank / lazy rules: FIFO_DENSE (measured cross-overs,
                                        profiles/r2_engines.md, r2_compact_state.md).  Default.             */
    ARCTE_ENGINE_FIFO_DENSE = 0,     /* one queue entry per warp iteration, dense 16-byte {s, r} per node and walk,
                                        32 walks per SM                                                     */
    ARCTE_ENGINE_BATCHED_DENSE = 1,  /* up to 32 queue entries (64 stored entries) per warp iteration, all pop checks
                                        and pushes of a batch applied at once, nodes referenced more than once
                                        accumulated in queue order by their first reference (shared memory);
                                        direct-mapped 32-byte {s, r, d_in, epoch} per node and walk: nothing is
                                        ever reset; 16 walks per SM                                          */
    ARCTE_ENGINE_BATCHED_HASH = 2,   /* the same batches, walk state in a growing open-addressing table of
                                        32-byte entries per walk (compact: the only engine whose memory does
                                        not grow with n per walk); experimental, slowest on the shapes measured */
    ARCTE_ENGINE_FIFO_COMPACT = 3    /* one queue entry per warp iteration; the pairs of a walk live in first-touch
                                        order in a compact array (first touches are coalesced writes, the threshold
                                        sweep is sequential, nothing is ever reset), found through a 4-byte
                                        epoch-tagged index map per node; all three rules                   */
};

/* Counters and device timings of the last arcte_cuda_extract/assemble on this context. */
typedef struct arcte_cuda_stats {
    int64_t n_seeds_total;   /* seeds selected on the graph (arcte.py:614-617)              */
    int64_t n_seeds_shard;   /* seeds this context processed                                 */
    int64_t pushes;          /* push operations (similarity.py return value, summed)          */
    int64_t edge_touches;    /* sum of deg(u) over pushes                                     */
    int64_t enqueues;        /* FIFO appends                                                  */
    int64_t max_queue;       /* longest live FIFO of any seed                                 */
    int64_t support;         /* sum over seeds of |{x : s[x] != 0}|                           */
    int64_t touched;         /* sum over seeds of nodes whose s or r left zero                */
    int64_t seed_degree;     /* sum over seeds of deg(seed)                                   */
    int64_t members;         /* sum over emitted seeds of community size                      */
    int64_t emitted;         /* seeds that emitted a community (arcte.py:370)                 */
    int64_t retries;         /* seeds re-run with a larger FIFO                               */
    int64_t n_slots;         /* concurrent per-warp walk states used                          */
    int64_t launches;        /* kernels launched by the last build/extract/assemble calls     */
    int64_t rounds;          /* frontier schedule: synchronous rounds, summed over seeds      */
    double ms_transition;    /* K1: degrees + row normalisation                               */
    double ms_seeds;         /* K2: seed selection/ordering + epsilon-effective               */
    double ms_push;          /* K3+K4: fused push / threshold / compaction kernel             */
    double ms_assemble;      /* K5: pack, transpose, splice                                    */
    double alg_bytes_push;   /* SURVEY 8(d) algorithmic bytes of the push kernel              */
    double slot_utilisation; /* mean busy time of a walk state / span of the push launch      */
    double ms_exchange;      /* multi-GPU: the all-to-all of community members (NCCL)          */
    int64_t engine;          /* ARCTE_ENGINE_* that walked the last extraction (-2: frontier)  */
} arcte_cuda_stats;

/* -- lifetime ------------------------------------------------------------- */
int arcte_cuda_device_count(int *count); /* CUDA devices visible to this process */
int arcte_cuda_create(arcte_cuda_ctx **out, int device_id);
void arcte_cuda_destroy(arcte_cuda_ctx *ctx);
const char *arcte_cuda_last_error(void);
/* Tuning knobs (0 keeps the default): concurrent walk states per SM, per-seed FIFO
   capacity in entries, fraction of free HBM the walk states may take (percent), initial
   capacity of the member buffer in entries.  FIFO rings and the member buffer grow on
   demand (affected seeds are re-run), so small values are safe, only slower. */
int arcte_cuda_configure(arcte_cuda_ctx *ctx, int warps_per_sm, int64_t queue_capacity,
                         int mem_percent, int64_t member_capacity);

/* Selects the walk schedule of arcte_cuda_extract / arcte_cuda_push on this context.  The
   remaining arguments tune the frontier schedule's launch geometry (0 or negative = default):
   the first heavy_permille/1000 of the count-descending seed list is walked by
   heavy_ctas_per_sm CTAs of heavy_threads threads per SM, the rest by light_ctas_per_sm CTAs of
   light_threads threads (threads: 32, 64, 128, 256, 512 or 1024). */
int arcte_cuda_set_schedule(arcte_cuda_ctx *ctx, int schedule, int heavy_permille, int heavy_threads,
                            int heavy_ctas_per_sm, int light_threads, int light_ctas_per_sm);

/* Selects the engine of the FIFO schedule (ARCTE_ENGINE_*); table_capacity > 0 bounds the entries of one
   walk's hash table half (rounded up to a power of two; a walk that outgrows it is re-run by
   ARCTE_ENGINE_FIFO_DENSE), 0 = as many as 2n or the memory budget allows.  The environment variable
   ARCTE_CUDA_ENGINE=fifo|dense|hash|compact overrides AUTO. */
int arcte_cuda_set_engine(arcte_cuda_ctx *ctx, int engine, int64_t table_capacity);

/* -- a11 + a1: graph upload and transition build --------------------------- */
/* Replaces the pickled (indices, indptr, data) hand-off of arcte.py:657-665 and
   get_natural_random_walk_matrix, eps_randomwalk/transition.py:43-68.  Copies the
   host CSR of the ADJACENCY matrix to the device and runs K1. */
int arcte_cuda_set_graph(arcte_cuda_ctx *ctx, int64_t n, int64_t nnz, const int64_t *host_indptr,
                         const int32_t *host_indices, const double *host_data);
/* Re-runs K1 (degrees, row normalisation) and K2a (seed selection) on the adjacency
   already resident in HBM -- the timed "inputs resident" form of the above. */
int arcte_cuda_build_transition(arcte_cuda_ctx *ctx);
/* W.data (nnz), out_degree (n), in_degree (n) to host: transition.py:99 return value. */
int arcte_cuda_get_transition(arcte_cuda_ctx *ctx, double *host_w, double *host_d_out,
                              double *host_d_in);

/* The same hand-off one level lower: the caller already holds W, out_degree and
   in_degree (exactly the raw arrays arcte_worker receives, arcte.py:279-286) and
   uploads them as they are; K1 is skipped, the seed list is still derived on the GPU. */
int arcte_cuda_set_transition(arcte_cuda_ctx *ctx, int64_t n, int64_t nnz, const int64_t *host_indptr,
                              const int32_t *host_indices, const double *host_w,
                              const double *host_d_out, const double *host_d_in);

/* -- a2: seed selection ---------------------------------------------------- */
/* arcte.py:610-617: nodes whose binarised column count is > 1, count-descending. */
int arcte_cuda_get_seed_count(arcte_cuda_ctx *ctx, int64_t *n_seeds);
int arcte_cuda_get_seeds(arcte_cuda_ctx *ctx, int64_t *host_seeds);
/* Replace the seed list by the caller's `iterate_nodes` (arcte_worker's first argument). */
int arcte_cuda_set_seeds(arcte_cuda_ctx *ctx, int64_t n_seeds, const int64_t *host_seeds);

/* -- a4: epsilon-effective -------------------------------------------------- */
/* calculate_epsilon_effective, arcte.py:26-50, for the given seed nodes. */
int arcte_cuda_epsilon_effective(arcte_cuda_ctx *ctx, double epsilon, int64_t n_seeds,
                                 const int64_t *host_seeds, double *host_eps_out);

/* -- a5-a7: operator seam, one seed ----------------------------------------- */
/* Same contract as similarity.py:149/:11/:66 on zeroed s and r: fills dense
   host_s[n], host_r[n] and the push count.  `rho` is used as given (the lazy
   worker passes lazy_rho, arcte.py:109); laziness factor is the reference's 0.5. */
int arcte_cuda_push(arcte_cuda_ctx *ctx, int rule, int64_t seed, double rho, double eps_eff,
                    double *host_s, double *host_r, int64_t *n_push);

/* -- a3 + a4-a8: extraction over a shard of the seed list -------------------- */
/* arcte_worker (arcte.py:279-388) for seed-list positions shard_rank,
   shard_rank + shard_count, ... (roundrobin_chunks, arcte.py:19-23).
   `rho` is the caller's restart probability; for ARCTE_RULE_LAZY the walk uses
   lazy_rho = 0.5 rho / (1 - 0.5 rho) exactly like the reference worker (arcte.py:109).
   host_eps_override: NULL, or one epsilon-effective per GLOBAL seed-list position
   (parity seam).  Results stay on the device as segments. */
int arcte_cuda_extract(arcte_cuda_ctx *ctx, int rule, double rho, double epsilon, int shard_rank,
                       int shard_count, const double *host_eps_override, int64_t *n_segments,
                       int64_t *n_members);
/* Segment k: seed node seg_seed[k], seg_count[k] members at members[seg_offset[k] ...]. */
int arcte_cuda_get_segments(arcte_cuda_ctx *ctx, int32_t *host_seg_seed, int32_t *host_seg_count,
                            int64_t *host_seg_offset, int32_t *host_members);
/* Device-resident views of the same four arrays (for an NCCL all-gather by the caller). */
int arcte_cuda_segments_device(arcte_cuda_ctx *ctx, const int32_t **dev_seg_seed,
                               const int32_t **dev_seg_count, const int64_t **dev_seg_offset,
                               const int32_t **dev_members);

/* Copies the same four arrays into caller-owned DEVICE buffers (e.g. tensors a
   collective library will send); sizes as returned by arcte_cuda_extract. */
int arcte_cuda_export_segments(arcte_cuda_ctx *ctx, int32_t *dev_seg_seed, int32_t *dev_seg_count,
                               int64_t *dev_seg_offset, int32_t *dev_members);

/* -- a9 + a10: assembly ------------------------------------------------------ */
/* arcte.py:379-386 and :670-688: features = hstack([I + pattern(A), local]) as a
   canonical n x 2n CSR.  Parts are segment sets (this context's own and/or those
   gathered from other GPUs), given as DEVICE pointers readable from this device.
   Parts living on another GPU of the box are first copied over NVLink (peer copy).
   n_parts == 0 assembles the context's own last extraction. */
int arcte_cuda_assemble(arcte_cuda_ctx *ctx, int n_parts, const int64_t *part_n_segments,
                        const int64_t *part_n_members, const int32_t *const *dev_seg_seed, const int32_t *const *dev_seg_count,
                        const int64_t *const *dev_seg_offset, const int32_t *const *dev_members,
                        int64_t *nnz_out);
/* Row-sharded form: assembles only rows [row_lo, row_hi) of the feature matrix (indptr has
   row_hi-row_lo+1 entries starting at 0).  With one process per GPU every rank builds its own
   row block from the all-gathered segments and the blocks are concatenated on one rank. */
int arcte_cuda_assemble_rows(arcte_cuda_ctx *ctx, int n_parts, const int64_t *part_n_segments,
                             const int64_t *part_n_members, const int32_t *const *dev_seg_seed,
                             const int32_t *const *dev_seg_count, const int64_t *const *dev_seg_offset,
                             const int32_t *const *dev_members, int64_t row_lo, int64_t row_hi,
                             int64_t *nnz_out);
int arcte_cuda_get_features(arcte_cuda_ctx *ctx, int64_t *host_indptr, int32_t *host_indices,
                            double *host_data);
/* The same copy for callers that know the values: with values_are_ones != 0 host_data is not copied but
   written as 1.0 by the host threads that stream the indices in (every stored value of the feature matrix
   is 1.0 except the identity entry of a self-loop row, which the caller patches to 2.0: arcte.py:379-381,
   :676-679).  Destinations are ordinary pageable memory (numpy arrays); the copy runs through a small
   fixed ring of pinned slots on n_threads host threads (<= 0: all, at most 16). */
int arcte_cuda_fetch_features(arcte_cuda_ctx *ctx, int64_t *host_indptr, int32_t *host_indices, double *host_data,
                              int values_are_ones, int n_threads);
/* Device-resident views of the assembled block (for a collective library). */
int arcte_cuda_features_device(arcte_cuda_ctx *ctx, const int64_t **dev_indptr, const int32_t **dev_indices,
                               const double **dev_data, int64_t *n_rows, int64_t *nnz);

/* -- f3: the centrality of arcte_and_centrality (embedding/arcte/cython_opt/arcte.pyx:125-241) ----------
   centrality[x] = sum over ALL nodes taken as seeds of s_seed[x] / d_in[x] with the RAW epsilon (arcte.pyx:164,
   :183-191).  Accumulated in 2^-38 fixed point (deterministic); the reference adds in seed order in floating
   point, so the two agree to about n * 2^-39 + rounding, far inside the push error bound. */
int arcte_cuda_centrality(arcte_cuda_ctx *ctx, double rho, double epsilon, double *host_centrality);

/* -- e: the multi-GPU exchange, inside the library (arcte.py:650-673) ------------------------
   Rank r of `world` walks shard r (arcte_cuda_extract with shard_rank = r, shard_count = world) and owns the
   rows [n r / world, n (r+1) / world) of the result.  arcte_cuda_exchange_assemble splits every community by
   destination row block on the device, exchanges the pieces with one grouped NCCL send/recv (an all-to-all:
   each member crosses NVLink at most once, values are never sent) and assembles this rank's row block
   (arcte_cuda_features_device / _fetch_features then see that block).  NCCL is loaded at run time
   (libnccl.so.2).  One process per GPU: rank 0 calls arcte_cuda_comm_unique_id, the caller distributes the
   128 bytes (e.g. torch.distributed.broadcast), every rank calls arcte_cuda_comm_init.  One process driving
   several GPUs: arcte_cuda_comm_init_all over its contexts (rank = position), then one host thread per
   context calls arcte_cuda_exchange_assemble concurrently. */
int arcte_cuda_comm_unique_id(void *id_out_128_bytes);
int arcte_cuda_comm_init(arcte_cuda_ctx *ctx, int world, int rank, const void *id_128_bytes);
int arcte_cuda_comm_init_all(arcte_cuda_ctx *const *ctxs, int n);
int arcte_cuda_comm_info(arcte_cuda_ctx *ctx, int *world, int *rank, int *nccl_version);
int arcte_cuda_exchange_assemble(arcte_cuda_ctx *ctx, int64_t *nnz_out);

/* -- after the path: column normalisation and community weighting ---------- */
/* These take and return HOST CSR arrays (canonical: sorted column indices) like the reference
   functions take and return scipy matrices; n_cols < 2^31.  Bit-exact against numpy/sklearn
   except for the logarithm inside normalize_columns / community_weighting (1-2 ulp). */

/* normalize_columns, embedding/common.py:49-67: every column with more than one stored entry
   is divided by sqrt(log(stored entries)).  Structure unchanged; host_data_out may alias
   host_data_in. */
int arcte_cuda_normalize_columns(arcte_cuda_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                                 const int32_t *host_indices, const double *host_data_in, double *host_data_out);
/* The same on the feature matrix the last arcte_cuda_assemble left on the device (all rows),
   in place, before arcte_cuda_get_features: arcte() followed by normalize_columns()
   (experiments/utility.py:207 + :66) without a host round trip. */
int arcte_cuda_normalize_features(arcte_cuda_ctx *ctx);

/* chi2_contingency_matrix, embedding/community_weighting.py:11-45.  X: n_rows x n_cols CSR
   (pattern only).  Y: n_rows x n_classes label matrix as CSR with integer-valued data, i.e.
   LabelBinarizer().fit_transform(y_train) (:19-21).  host_out: n_classes x n_cols, row-major. */
int arcte_cuda_chi2_contingency(arcte_cuda_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *host_x_indptr,
                                const int32_t *host_x_indices, int64_t n_classes, const int64_t *host_y_indptr,
                                const int32_t *host_y_indices, const double *host_y_data, double *host_out);
/* peak_snr_weight_aggregation, embedding/community_weighting.py:48-84.  The matrix is changed
   in place like the reference does (nan -> 0, :49); host_weights_out has n_cols entries. */
int arcte_cuda_peak_snr(arcte_cuda_ctx *ctx, int64_t n_classes, int64_t n_cols, double *host_cm_inout,
                        double *host_weights_out);
/* The two above back to back with the n_classes x n_cols matrix kept in HBM
   (chi2_psnr_community_weighting, community_weighting.py:128-131, first two lines). */
int arcte_cuda_chi2_psnr_weights(arcte_cuda_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *host_x_indptr,
                                 const int32_t *host_x_indices, int64_t n_classes, const int64_t *host_y_indptr,
                                 const int32_t *host_y_indices, const double *host_y_data, double *host_weights_out);
/* community_weighting, embedding/community_weighting.py:87-125, for ONE matrix (the reference
   treats X_train and X_test identically and independently): columns with more than one stored
   entry are multiplied by 0 if weight == 0 else log(1 + weight); explicit zeros are dropped;
   rows are l2-normalised with sklearn's left-to-right sum of squares.  Output buffers are
   caller-allocated with room for the input's nnz; *out_nnz is the number of entries kept. */
int arcte_cuda_community_weighting(arcte_cuda_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                                   const int32_t *host_indices, const double *host_data,
                                   const double *host_weights, int64_t *host_out_indptr, int32_t *host_out_indices,
                                   double *host_out_data, int64_t *out_nnz);

/* The experiment loop (experiments/utility.py:83-104) slices the SAME feature matrix into
   train/test rows for every fold and weights both blocks.  These calls keep the matrix in HBM:
   store it once, then per fold gather the two row blocks on the device, compute the chi2 /
   peak-SNR weights from the training block and weight both blocks; only the weighted blocks
   travel back.  Results are identical to slicing on the host and calling
   arcte_cuda_chi2_psnr_weights + arcte_cuda_community_weighting. */
int arcte_cuda_store_features(arcte_cuda_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *host_indptr,
                              const int32_t *host_indices, const double *host_data);
/* The same, adopting the matrix arcte_cuda_assemble (+ arcte_cuda_normalize_features) left on
   the device: no host round trip between extraction and the experiment loop. */
int arcte_cuda_store_assembled(arcte_cuda_ctx *ctx);
/* host_y_*: label matrix of the TRAINING rows, in the order of host_train_rows (n_train x n_classes). */
int arcte_cuda_weighted_fold(arcte_cuda_ctx *ctx, int64_t n_train, const int64_t *host_train_rows, int64_t n_test,
                             const int64_t *host_test_rows, int64_t n_classes, const int64_t *host_y_indptr,
                             const int32_t *host_y_indices, const double *host_y_data, int64_t *train_nnz,
                             int64_t *test_nnz);
/* which: 0 = weighted training block, 1 = weighted test block (sizes from arcte_cuda_weighted_fold). */
int arcte_cuda_get_fold(arcte_cuda_ctx *ctx, int which, int64_t *host_indptr, int32_t *host_indices,
                        double *host_data);

/* -- text I/O of the console script (host only) ------------------------ */
/* read_adjacency_matrix, datautil/datarw.py:54-120: parses `src<sep>dst<sep>weight` rows
   ('#' comments) with all host threads (n_threads <= 0: every hardware thread), renumbers the
   node ids 0..n-1 in first-seen order (source before target) and, if `undirected`, appends the
   reciprocal entry after each edge (self loops once).  The entries come back in file order as
   COO triplets; node_ids[i] is the original id of node i (the reference's node_to_id). */
typedef struct arcte_cuda_edge_list arcte_cuda_edge_list;
int arcte_cuda_io_read_edge_list(const char *path, const char *separator, int undirected, int n_threads,
                                 arcte_cuda_edge_list **out, int64_t *n_nodes, int64_t *n_entries);
int arcte_cuda_io_edge_list_copy(const arcte_cuda_edge_list *el, int64_t *row, int64_t *col, double *data,
                                 int64_t *node_ids);
void arcte_cuda_io_edge_list_free(arcte_cuda_edge_list *el);
/* write_features, datautil/datarw.py:123-143: one `node_id<sep>column<sep>int(value)` line per
   stored entry of the CSR, row-major, formatted by all host threads and written with pwrite at
   precomputed offsets. */
int arcte_cuda_io_write_features(const char *path, const char *separator, int64_t n_rows, const int64_t *indptr,
                                 const int32_t *indices, const double *data, const int64_t *node_ids,
                                 int n_threads, int64_t *bytes_written);

/* One process per GPU: the same copy with the destination arrays living in ANOTHER process (`pid`, e.g. rank 0 of
   a torchrun job; the three pointers are virtual addresses of that process).  The pinned slots are written there
   with process_vm_writev, so every rank brings its own row block home through its own PCIe link and nothing is
   staged in shared memory.  Needs ptrace permission on the target (same user, Yama scope <= 1); pid = own pid or
   0 is the local copy.  arcte_cuda_host_write_to copies a small host buffer the same way;
   arcte_cuda_host_advise_huge asks for huge pages for a range of THIS process (the owner of the destination
   calls it before the first touch). */
int arcte_cuda_fetch_features_to(arcte_cuda_ctx *ctx, int64_t pid, int64_t *dst_indptr, int32_t *dst_indices,
                                 double *dst_data, int values_are_ones, int n_threads);
int arcte_cuda_host_write_to(int64_t pid, void *dst, const void *src, int64_t bytes);
int arcte_cuda_host_advise_huge(void *p, int64_t bytes);

/* A writable array of `count` doubles, all 1.0, that costs no fill: a 32 MB in-memory file of ones (created once
   per process) mapped copy-on-write back to back over the whole range.  *mapped_bytes is what to hand to
   arcte_cuda_host_ones_free.  The feature matrix's value array is built this way (every stored value is 1.0 except
   self-loop diagonals, which the caller patches: those pages are copied on that write). */
int arcte_cuda_host_ones_alloc(int64_t count, void **out, int64_t *mapped_bytes);
int arcte_cuda_host_ones_free(void *p, int64_t mapped_bytes);

/* -- page-locked host memory for results ------------------------------------- */
/* cudaHostAlloc / cudaFreeHost: result buffers handed to arcte_cuda_get_features can be
   page-locked so the device-to-host copy runs at PCIe rate instead of through the
   driver's staging buffer.  Not tied to a context. */
int arcte_cuda_host_alloc(void **out, int64_t bytes);
int arcte_cuda_host_free(void *p);
/* Fills count doubles with `value` using n_threads host threads (<= 0: all).  Every stored value
   of a feature matrix is 1.0 except self-loop diagonals (arcte.py:379-381, :676-679), so a pooled
   page-locked block that already holds ones saves the device-to-host copy of the value array. */
int arcte_cuda_host_fill_f64(double *p, int64_t count, double value, int n_threads);

/* -- measurement helpers ------------------------------------------------------ */
/* CUDA events on the context's own stream (the stream every kernel of this library is
   launched on), so a caller can time a whole step on the device. */
int arcte_cuda_timer_start(arcte_cuda_ctx *ctx);
int arcte_cuda_timer_stop(arcte_cuda_ctx *ctx, double *elapsed_ms);
/* Overwrites a scratch buffer larger than the 126 MB L2 (cold-cache timing). */
int arcte_cuda_flush_l2(arcte_cuda_ctx *ctx);

/* 64-bit content hash of the assembled block (row starts, column indices, non-unit values, each mixed with its
   global position; terms summed modulo 2^64).  row_lo / nnz_lo: position of the block in the whole matrix, so
   that the hashes of the row blocks of several GPUs ADD UP to the hash of the matrix one GPU assembles.
   reveal_graph_embedding_b200.engine.csr_hash computes the same number from a scipy matrix on the host. */
int arcte_cuda_features_hash(arcte_cuda_ctx *ctx, int64_t row_lo, int64_t nnz_lo, uint64_t *hash_out);

/* -- stats ------------------------------------------------------------------- */
int arcte_cuda_get_stats(arcte_cuda_ctx *ctx, arcte_cuda_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* ARCTE_CUDA_H */
