"""ARCTE feature extraction on B200 GPUs -- drop-in for
reveal_graph_embedding/embedding/arcte/arcte.py of the reference.

Same entry points, argument meaning and return value:

    arcte(adjacency_matrix, rho, epsilon, number_of_threads=None)                  arcte.py:591
    arcte_with_pagerank(adjacency_matrix, rho, epsilon, number_of_threads=None)    arcte.py:491
    arcte_with_lazy_pagerank(adjacency_matrix, rho, epsilon, number_of_threads=None) arcte.py:391
    arcte_worker(iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon)  arcte.py:279

each returning the scipy.sparse CSR the reference returns (n x 2n, float64, canonical;
the workers n x n).  `number_of_threads` used to be the size of the multiprocessing pool
(arcte.py:650); here it caps the number of GPUs the seeds are sharded over (None = all
visible GPUs).  Inside a torch.distributed job (one process per GPU, e.g. torchrun) every
rank calls the function with the same matrix and processes its round-robin shard of the
seeds; after an NCCL all-gather of the per-GPU segments rank 0 assembles and returns the
full matrix (the other ranks return None; ARCTE_CUDA_RESULT_ON_ALL_RANKS=1 returns it
everywhere).

Differences from the reference, all on purpose:
  * no silent degradation: the reference prints and returns the base features when the
    final hstack fails (arcte.py:684-686) and drops worker exceptions (arcte.py:657-666);
    here every failure raises.
  * number_of_threads larger than the number of seeds works (the reference crashes on an
    empty chunk, arcte.py:19-23).
"""
import os
import threading
import time

import numpy as np
import scipy.sparse as sparse

from ... import distributed, hostmem
from ...engine import (RULE_ABSORBING, RULE_LAZY, RULE_PAGERANK, Engine, canonical_csr, device_count,
                       get_engine)


_SIDE_THREAD_MIN_NNZ = 1 << 20   # below this a thread costs more than the scan it hides


def _quietly(fn):
    try:
        fn()
    except Exception:
        pass


def _extract_features(adjacency_matrix, rho, epsilon, number_of_threads, rule):
    A = canonical_csr(adjacency_matrix)
    rho_eff = rho  # the lazy_rho substitution of arcte.py:109 happens inside the library

    if distributed.is_active():
        return distributed.arcte_distributed(A, rule, rho_eff, epsilon)

    visible = device_count()
    if visible < 1:
        raise RuntimeError("arcte: no CUDA device visible (this package has no CPU path)")
    n_gpus = visible if number_of_threads is None else max(1, min(int(number_of_threads), visible))

    if n_gpus == 1:
        eng = get_engine(0)
        t = [time.perf_counter()]
        eng.set_graph(A, canonical=True); t.append(time.perf_counter())
        # the rows with a stored diagonal (their identity entry is 2.0, arcte.py:676-679) depend on the input only:
        # found on a side thread while the device walks (a failure there is met again, and raised, by features())
        side = threading.Thread(target=_quietly, args=(eng.self_loop_rows,)) if A.nnz >= _SIDE_THREAD_MIN_NNZ else None
        if side:
            side.start()
        try:
            eng.extract(rule, rho_eff, epsilon); t.append(time.perf_counter())
            eng.assemble(); t.append(time.perf_counter())
        finally:
            if side:
                side.join()
        X = eng.features(); t.append(time.perf_counter())
        if os.environ.get("ARCTE_CUDA_DEBUG"):
            import sys
            print("[arcte] 1 GPU: set_graph %.1f ms, extract %.1f ms, assemble %.1f ms, features %.1f ms"
                  % tuple(1e3 * (b - a) for a, b in zip(t[:-1], t[1:])), file=sys.stderr)
        return X

    # single process, several GPUs: graph replicated, seeds dealt round-robin (arcte.py:651),
    # one host thread per GPU (ctypes drops the GIL).  After the walks every GPU pulls all
    # segments over NVLink (peer copies), assembles its own block of rows and copies it
    # straight into its slice of the result arrays -- one PCIe link per GPU.
    engines = [get_engine(d) for d in range(n_gpus)]
    n = A.shape[0]

    def run_parallel(fn):
        errors = []

        def call(rank):
            try:
                fn(rank)
            except BaseException as exc:  # re-raised on the caller's thread
                errors.append(exc)

        threads = [threading.Thread(target=call, args=(r,)) for r in range(n_gpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def walk(rank):
        e = engines[rank]
        e.set_graph(A, canonical=True)
        if rank == 0:   # self-loop rows of the input (arcte.py:676-679), once, while the devices walk; shared below
            side = threading.Thread(target=_quietly, args=(e.self_loop_rows,))
            side.start()
        try:
            e.extract(rule, rho_eff, epsilon, shard_rank=rank, shard_count=n_gpus)
        finally:
            if rank == 0:
                side.join()

    dbg = os.environ.get("ARCTE_CUDA_DEBUG")
    t0 = time.perf_counter()
    run_parallel(walk)
    for e in engines[1:]:
        e._loops = engines[0]._loops   # same input matrix on every GPU
    t1 = time.perf_counter()
    # one exchange step inside the library: NCCL all-to-all of the communities split by row block, then
    # every GPU assembles its own rows (csrc/exchange.cu)
    if any(e.comm_info()[:2] != (n_gpus, r) for r, e in enumerate(engines)):
        Engine.comm_init_all(engines)
    run_parallel(lambda rank: engines[rank].exchange_assemble())
    t2 = time.perf_counter()
    nnz = np.array([e.out_nnz for e in engines], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(nnz)])
    total = int(offsets[-1])
    # plain numpy result arrays; every GPU streams its row block into its slice over its own PCIe link
    # (csrc/hostcopy.cu).  Every stored value is 1.0 except self-loop diagonals (arcte.py:379-381, :676-679):
    # the value array is copy-on-write mappings of a block of ones (hostmem.ones), neither copied nor written.
    indices = np.empty(total, dtype=np.int32)
    data = hostmem.ones(total)
    indptr = np.empty(n + 1, dtype=np.int64)
    threads_each = max(2, (os.cpu_count() or 8) // n_gpus)

    def fetch(rank):
        lo, hi = distributed.row_range(n, rank, n_gpus)
        ip = np.empty(hi - lo + 1, dtype=np.int64)
        o0, o1 = int(offsets[rank]), int(offsets[rank + 1])
        engines[rank].fetch_block(ip, indices[o0:o1], None, n_threads=threads_each)
        engines[rank].patch_self_loops(data[o0:o1], ip, lo, hi)
        indptr[lo:hi + 1] = ip + offsets[rank]   # neighbouring blocks write the same value at their common row

    run_parallel(fetch)
    t3 = time.perf_counter()
    if dbg:
        import sys
        print("[arcte] in-process %d GPUs: upload+walk %.1f ms, exchange+assemble %.1f ms, fetch %.1f ms"
              % (n_gpus, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), file=sys.stderr)
    if max(2 * n, total) < 2 ** 31:
        indptr = indptr.astype(np.int32)
    else:
        indices = indices.astype(np.int64)
    return sparse.csr_matrix((data, indices, indptr), shape=(n, 2 * n), copy=False)


def arcte(adjacency_matrix, rho, epsilon, number_of_threads=None):
    """Local-community features from absorbing regularised commute times (arcte.py:591-688)."""
    return _extract_features(adjacency_matrix, rho, epsilon, number_of_threads, RULE_ABSORBING)


def arcte_with_pagerank(adjacency_matrix, rho, epsilon, number_of_threads=None):
    """Same with personalised PageRank vectors (arcte.py:491-588)."""
    return _extract_features(adjacency_matrix, rho, epsilon, number_of_threads, RULE_PAGERANK)


def arcte_with_lazy_pagerank(adjacency_matrix, rho, epsilon, number_of_threads=None):
    """Same with lazy personalised PageRank vectors (arcte.py:391-488)."""
    return _extract_features(adjacency_matrix, rho, epsilon, number_of_threads, RULE_LAZY)


def arcte_and_centrality(adjacency_matrix, rho, epsilon):
    """The reference's Cython-only variant (embedding/arcte/cython_opt/arcte.pyx:125-241): ARCTE features plus
    the RCT centrality vector.

    centrality[x] = sum over ALL nodes taken as seeds of s_seed[x] / in_degree[x], the walks run with the RAW
    epsilon (arcte.pyx:164, :183-191; nodes without out-edges get 1.0, :222 -- none after transition.py:58).
    It is computed on the GPU in 2^-38 fixed point (deterministic) and agrees with a restatement of those lines
    to ~1e-9 relative (tests/test_gpu_parity.py::test_centrality_vs_restatement).

    The feature half of that function is NOT reproduced: it numbers communities compactly in an order that
    depends on an unstable argsort of tied values (arcte.pyx:194-212), its `centrality += s_sparse` no longer
    type-checks under current scipy, and its build is commented out in the reference's setup.py:4-12, so
    there is no runnable behaviour to be identical to.  The features returned here are arcte()'s (arcte.py:591),
    which is what the reference's console script and experiments use."""
    A = canonical_csr(adjacency_matrix)
    eng = get_engine(0)
    eng.set_graph(A, canonical=True)
    centrality = eng.centrality(rho, epsilon)
    eng.extract(RULE_ABSORBING, rho, epsilon)
    eng.assemble()
    return eng.features(), centrality


def _worker(rule, iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon):
    eng = get_engine(0)
    eng.set_transition(indptr_c, indices_c, data_c, out_degree, in_degree)
    eng.set_seeds(np.asarray(iterate_nodes, dtype=np.int64))
    eng.extract(rule, rho, epsilon)
    eng.assemble()
    n = eng.n
    return sparse.csr_matrix(eng.features()[:, n:])


def arcte_worker(iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon):
    """The reference's per-process worker (arcte.py:279-388): n x n local-community block
    for the given seed nodes, from the raw transition-matrix arrays."""
    return _worker(RULE_ABSORBING, iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon)


def arcte_with_pagerank_worker(iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon):
    """arcte.py:167-276."""
    return _worker(RULE_PAGERANK, iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon)


def arcte_with_lazy_pagerank_worker(iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho,
                                    epsilon):
    """arcte.py:53-164."""
    return _worker(RULE_LAZY, iterate_nodes, indices_c, indptr_c, data_c, out_degree, in_degree, rho, epsilon)


def calculate_epsilon_effective(rho, epsilon, seed_degree, neighbor_degrees, mean_degree):
    """arcte.py:26-50 on the GPU.  The reference signature works on bare degree values; the
    device kernel works on a graph, so a star graph with those degrees is built on the fly
    (node 0 = the seed).  Meant for parity tests, not for speed."""
    nb = np.asarray(neighbor_degrees, dtype=np.float64)
    k = nb.size
    indptr = np.concatenate([[0, k], k + np.arange(1, k + 1)]).astype(np.int64)
    indices = np.concatenate([np.arange(1, k + 1), np.zeros(k)]).astype(np.int32)
    w = np.ones(2 * k)
    d_out = np.concatenate([[float(seed_degree)], nb])
    eng = get_engine(0)
    eng.set_transition(indptr, indices, w, d_out, d_out.copy())
    return float(eng.epsilon_effective(epsilon, [0])[0])
