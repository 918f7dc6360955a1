"""Host-side driver of one GPU's libarcte_cuda context.

`Engine` is a thin object over the C ABI (include/arcte_cuda.h): it owns one
context (one GPU), converts scipy/numpy containers to the plain arrays the ABI
takes and turns status codes into exceptions.  No computation happens here.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sparse

from . import _lib, hostmem
from ._lib import (RULE_ABSORBING, RULE_LAZY, RULE_PAGERANK, SCHEDULE_FIFO, SCHEDULE_FRONTIER,  # noqa: F401
                   ArcteCudaError, check, ptr)

_SCHEDULES = {"fifo": SCHEDULE_FIFO, "exact": SCHEDULE_FIFO, "frontier": SCHEDULE_FRONTIER,
              SCHEDULE_FIFO: SCHEDULE_FIFO, SCHEDULE_FRONTIER: SCHEDULE_FRONTIER}
# "auto": frontier when a shard has at most this many seeds (the FIFO launch is then as long as its
# longest walk and the frontier schedule is 2-12x faster, profiles/r1_frontier_schedule.md), FIFO
# above it and for the PageRank rules.  Never the default: the two schedules differ in tie entries.
AUTO_FRONTIER_MAX_SEEDS = 40000
_default_schedule = [None]  # set_default_schedule(); None = ARCTE_CUDA_SCHEDULE or "fifo"


def set_default_schedule(schedule):
    """Walk schedule of every engine created or fetched from now on: "fifo" (exact replay of
    the reference's queue, the default) or "frontier" (synchronous fixed-point rounds:
    deterministic, same error bound, support equal up to in-band ties; several times faster).
    The environment variable ARCTE_CUDA_SCHEDULE sets the same default."""
    if schedule is not None and schedule != "auto" and schedule not in _SCHEDULES:
        raise ValueError("unknown schedule %r" % (schedule,))
    _default_schedule[0] = schedule
    for dev in list(_ENGINES):
        if _ENGINES[dev]._h is not None:
            get_engine(dev)


def _resolve_schedule():
    import os
    name = _default_schedule[0]
    if name is None:
        name = os.environ.get("ARCTE_CUDA_SCHEDULE", "fifo").strip().lower() or "fifo"
    if name != "auto" and name not in _SCHEDULES:
        raise ValueError("ARCTE_CUDA_SCHEDULE must be 'fifo', 'frontier' or 'auto', got %r" % (name,))
    return name


def canonical_csr(adjacency_matrix):
    """What the reference feeds its workers: float64 CSR with sorted indices
    (transition.py:52,65).  Duplicate entries are summed (the reference's numpy
    fancy-index `+=` is ill-defined on them)."""
    if sparse.isspmatrix_csr(adjacency_matrix) and adjacency_matrix.dtype == np.float64:
        A = adjacency_matrix  # as is: scipy caches the canonical-format check on the object
    else:
        A = sparse.csr_matrix(adjacency_matrix, dtype=np.float64)
    if A.shape[0] != A.shape[1]:
        raise ValueError("adjacency matrix must be square, got %r" % (A.shape,))
    if not A.has_canonical_format:
        A = A.copy()  # never touch the caller's arrays
        A.sum_duplicates()
        A.sort_indices()
    return A


def _mix64(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def csr_hash(X, chunk=1 << 24):
    """The number arcte_cuda_features_hash gives for the same matrix, computed on the host from a scipy CSR:
    sum modulo 2^64 of splitmix64 terms over the row starts, the column indices and the values other than 1.0,
    each mixed with its position."""
    with np.errstate(over="ignore"):
        h = np.uint64(0)
        n, nnz = X.shape[0], int(X.indptr[-1])
        for lo in range(0, n, chunk):
            i = np.arange(lo, min(n, lo + chunk), dtype=np.uint64)
            h += _mix64(_mix64(np.uint64(0x1000000000000000) + i) + X.indptr[lo:lo + i.size].astype(np.uint64)).sum(dtype=np.uint64)
        for lo in range(0, nnz, chunk):
            k = np.arange(lo, min(nnz, lo + chunk), dtype=np.uint64)
            h += _mix64(_mix64(np.uint64(0x2000000000000000) + k) + X.indices[lo:lo + k.size].astype(np.uint64)).sum(dtype=np.uint64)
            v = X.data[lo:lo + k.size]
            m = v != 1.0
            if m.any():
                h += _mix64(_mix64(np.uint64(0x3000000000000000) + k[m]) + v[m].view(np.uint64)).sum(dtype=np.uint64)
        return int(h)


class Engine:
    def __init__(self, device=0):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.arcte_cuda_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        self.n = 0
        self.nnz = 0
        self._loops = None
        self._values_structural = False
        self.schedule = SCHEDULE_FIFO
        self.auto_schedule = False

    def close(self):
        if getattr(self, "_h", None):
            self._L.arcte_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration -----------------------------------------------------------------
    def configure(self, warps_per_sm=0, queue_capacity=0, mem_percent=0, member_capacity=0):
        check(self._L.arcte_cuda_configure(self._h, int(warps_per_sm), int(queue_capacity), int(mem_percent),
                                           int(member_capacity)))

    def set_schedule(self, schedule, heavy_permille=-1, heavy_threads=0, heavy_ctas_per_sm=0, light_threads=0,
                     light_ctas_per_sm=0):
        """"fifo" / "frontier" (include/arcte_cuda.h: ARCTE_SCHEDULE_*); the other arguments tune the
        frontier schedule's launch geometry (defaults when 0 / negative)."""
        if schedule not in _SCHEDULES:
            raise ValueError("unknown schedule %r" % (schedule,))
        check(self._L.arcte_cuda_set_schedule(self._h, _SCHEDULES[schedule], int(heavy_permille), int(heavy_threads),
                                              int(heavy_ctas_per_sm), int(light_threads), int(light_ctas_per_sm)))
        self.schedule = _SCHEDULES[schedule]

    def set_engine(self, engine="auto", table_capacity=0):
        """Engine of the exact FIFO schedule (include/arcte_cuda.h: ARCTE_ENGINE_*): "auto", "fifo" (one
        queue entry per warp iteration, dense state), "dense" (batched, dense state), "hash" (batched,
        compact per-walk hash tables) or "compact" (one entry per iteration, pairs in first-touch order behind an
        epoch-tagged index map).  All of them give bit-identical results."""
        names = {"auto": _lib.ENGINE_AUTO, "fifo": _lib.ENGINE_FIFO_DENSE, "dense": _lib.ENGINE_BATCHED_DENSE,
                 "hash": _lib.ENGINE_BATCHED_HASH, "compact": _lib.ENGINE_FIFO_COMPACT}
        if engine not in names:
            raise ValueError("unknown engine %r" % (engine,))
        check(self._L.arcte_cuda_set_engine(self._h, names[engine], int(table_capacity)))

    # -- a11 + a1 -------------------------------------------------------------------------
    def set_graph(self, adjacency_matrix, canonical=False):
        """Upload the adjacency CSR and build the transition matrix (K1) and seed list (K2a)."""
        A = adjacency_matrix if canonical else canonical_csr(adjacency_matrix)
        self.n, self.nnz = int(A.shape[0]), int(A.nnz)
        self._loops = None
        self._indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self._indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        self._data = np.ascontiguousarray(A.data, dtype=np.float64)
        check(self._L.arcte_cuda_set_graph(self._h, self.n, self.nnz, ptr(self._indptr), ptr(self._indices),
                                           ptr(self._data)))

    def set_transition(self, indptr, indices, w, d_out, d_in):
        """Upload W, out_degree, in_degree as given (arcte_worker's raw arrays, arcte.py:279-286)."""
        self._loops = None
        self._indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self._indices = np.ascontiguousarray(indices, dtype=np.int32)
        self._data = np.ascontiguousarray(w, dtype=np.float64)
        d_out = np.ascontiguousarray(d_out, dtype=np.float64)
        d_in = np.ascontiguousarray(d_in, dtype=np.float64)
        self.n, self.nnz = int(self._indptr.size - 1), int(self._indices.size)
        if d_out.size != self.n or d_in.size != self.n or self._data.size != self.nnz:
            raise ValueError("inconsistent transition arrays")
        check(self._L.arcte_cuda_set_transition(self._h, self.n, self.nnz, ptr(self._indptr), ptr(self._indices),
                                                ptr(self._data), ptr(d_out), ptr(d_in)))

    def set_seeds(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        check(self._L.arcte_cuda_set_seeds(self._h, seeds.size, ptr(seeds)))

    def build_transition(self):
        check(self._L.arcte_cuda_build_transition(self._h))

    def transition(self):
        """(W.data, out_degree, in_degree) like transition.py:99."""
        w = np.empty(max(self.nnz, 1), dtype=np.float64)
        d_out = np.empty(self.n, dtype=np.float64)
        d_in = np.empty(self.n, dtype=np.float64)
        check(self._L.arcte_cuda_get_transition(self._h, ptr(w), ptr(d_out), ptr(d_in)))
        return w[:self.nnz], d_out, d_in

    # -- a2 -----------------------------------------------------------------------------------
    def seeds(self):
        k = C.c_int64()
        check(self._L.arcte_cuda_get_seed_count(self._h, C.byref(k)))
        out = np.empty(max(k.value, 1), dtype=np.int64)
        check(self._L.arcte_cuda_get_seeds(self._h, ptr(out)))
        return out[:k.value]

    # -- a4 -----------------------------------------------------------------------------------
    def epsilon_effective(self, epsilon, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        out = np.empty(max(seeds.size, 1), dtype=np.float64)
        check(self._L.arcte_cuda_epsilon_effective(self._h, float(epsilon), seeds.size, ptr(seeds), ptr(out)))
        return out[:seeds.size]

    # -- a5-a7 ----------------------------------------------------------------------------------
    def push(self, rule, seed, rho, eps_eff):
        s = np.empty(self.n, dtype=np.float64)
        r = np.empty(self.n, dtype=np.float64)
        nop = C.c_int64()
        check(self._L.arcte_cuda_push(self._h, int(rule), int(seed), float(rho), float(eps_eff), ptr(s), ptr(r),
                                      C.byref(nop)))
        return s, r, nop.value

    # -- a3-a8 ----------------------------------------------------------------------------------
    def extract(self, rule, rho, epsilon, shard_rank=0, shard_count=1, eps_override=None):
        ov = None
        if eps_override is not None:
            ov = np.ascontiguousarray(eps_override, dtype=np.float64)
        if self.auto_schedule:
            k = C.c_int64()
            check(self._L.arcte_cuda_get_seed_count(self._h, C.byref(k)))
            per_shard = -(-k.value // max(int(shard_count), 1))
            want = (SCHEDULE_FRONTIER if int(rule) == RULE_ABSORBING and per_shard <= AUTO_FRONTIER_MAX_SEEDS
                    else SCHEDULE_FIFO)
            if want != self.schedule:
                self.set_schedule(want)
        ns, nm = C.c_int64(), C.c_int64()
        check(self._L.arcte_cuda_extract(self._h, int(rule), float(rho), float(epsilon), int(shard_rank),
                                         int(shard_count), ptr(ov), C.byref(ns), C.byref(nm)))
        self.n_segments, self.n_members = ns.value, nm.value
        return ns.value, nm.value

    def centrality(self, rho, epsilon):
        """RCT centrality of arcte_and_centrality (cython_opt/arcte.pyx:125-241) for the uploaded graph: sum over
        all nodes taken as seeds of s/d_in, raw epsilon.  float64 [n]."""
        out = np.empty(self.n, dtype=np.float64)
        check(self._L.arcte_cuda_centrality(self._h, float(rho), float(epsilon), ptr(out)))
        return out

    def segments(self):
        S, M = self.n_segments, self.n_members
        seed = np.empty(max(S, 1), dtype=np.int32)
        cnt = np.empty(max(S, 1), dtype=np.int32)
        off = np.empty(max(S, 1), dtype=np.int64)
        mem = np.empty(max(M, 1), dtype=np.int32)
        check(self._L.arcte_cuda_get_segments(self._h, ptr(seed), ptr(cnt), ptr(off), ptr(mem)))
        return seed[:S], cnt[:S], off[:S], mem[:M]

    def segments_device(self):
        """Raw device addresses (ints) of (seg_seed, seg_count, seg_offset, members)."""
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._L.arcte_cuda_segments_device(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value or 0, b.value or 0, c.value or 0, d.value or 0

    def export_segments(self, seed_ptr, count_ptr, offset_ptr, members_ptr):
        """D2D copy of the segments into caller-owned device buffers (addresses as ints)."""
        check(self._L.arcte_cuda_export_segments(self._h, C.c_void_p(seed_ptr), C.c_void_p(count_ptr),
                                                 C.c_void_p(offset_ptr), C.c_void_p(members_ptr)))

    # -- a9-a10 ---------------------------------------------------------------------------------
    def assemble(self, parts=None, row_lo=0, row_hi=None):
        """parts: None (own segments) or a list of (n_segments, n_members, seed_ptr, count_ptr,
        offset_ptr, members_ptr) with device addresses as ints.  row_lo/row_hi: assemble only
        that row block of the feature matrix (default: all rows)."""
        nnz = C.c_int64()
        row_hi = self.n if row_hi is None else int(row_hi)
        if not parts:
            check(self._L.arcte_cuda_assemble_rows(self._h, 0, None, None, None, None, None, None, int(row_lo),
                                                   row_hi, C.byref(nnz)))
        else:
            P = len(parts)
            ns = np.array([p[0] for p in parts], dtype=np.int64)
            nm = np.array([p[1] for p in parts], dtype=np.int64)
            cols = [np.array([p[i] for p in parts], dtype=np.uint64) for i in (2, 3, 4, 5)]
            check(self._L.arcte_cuda_assemble_rows(self._h, P, ptr(ns), ptr(nm), ptr(cols[0]), ptr(cols[1]),
                                                   ptr(cols[2]), ptr(cols[3]), int(row_lo), row_hi, C.byref(nnz)))
        self.out_nnz = nnz.value
        self.out_rows = row_hi - int(row_lo)
        self.out_row_lo = int(row_lo)
        self._values_structural = True
        return nnz.value

    # -- (e) multi-GPU exchange inside the library (csrc/exchange.cu) ---------------------------
    def comm_unique_id(self):
        """128-byte NCCL id (rank 0 creates it, the caller distributes it)."""
        buf = C.create_string_buffer(128)
        check(self._L.arcte_cuda_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, world, rank, unique_id):
        if len(unique_id) != 128:
            raise ValueError("the NCCL id is 128 bytes")
        check(self._L.arcte_cuda_comm_init(self._h, int(world), int(rank), C.create_string_buffer(unique_id, 128)))

    def comm_info(self):
        """(world, rank, nccl_version) of this context's communicator; world 0 = none."""
        w, r, v = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self._L.arcte_cuda_comm_info(self._h, C.byref(w), C.byref(r), C.byref(v)))
        return w.value, r.value, v.value

    @staticmethod
    def comm_init_all(engines):
        """One process driving several GPUs: a communicator over `engines`, rank = position."""
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        check(_lib.load().arcte_cuda_comm_init_all(arr, len(engines)))

    def exchange_assemble(self):
        """After extract(shard_rank = comm rank, shard_count = comm world): split the communities by
        destination row block, all-to-all over NCCL, assemble this rank's rows.  Returns their nnz."""
        nnz = C.c_int64(0)
        check(self._L.arcte_cuda_exchange_assemble(self._h, C.byref(nnz)))
        world, rank, _ = self.comm_info()
        self.out_nnz = nnz.value
        self.out_row_lo = (self.n * rank) // world
        self.out_rows = (self.n * (rank + 1)) // world - self.out_row_lo
        self._values_structural = True
        return nnz.value

    def features_hash(self, nnz_lo=0):
        """64-bit content hash of the assembled block (include/arcte_cuda.h); nnz_lo: stored entries of the row
        blocks before this one.  Hashes of row blocks add up (mod 2^64) to csr_hash() of the whole matrix."""
        h = C.c_uint64(0)
        check(self._L.arcte_cuda_features_hash(self._h, int(getattr(self, "out_row_lo", 0)), int(nnz_lo), C.byref(h)))
        return int(h.value)

    def features_device(self):
        """(indptr_ptr, indices_ptr, data_ptr, n_rows, nnz) of the assembled block on the device."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        r, z = C.c_int64(), C.c_int64()
        check(self._L.arcte_cuda_features_device(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(r),
                                                 C.byref(z)))
        return a.value or 0, b.value or 0, c.value or 0, r.value, z.value

    def features_block(self):
        """Raw (indptr, indices, data) of the assembled row block (indptr starts at 0)."""
        nnz, rows = self.out_nnz, getattr(self, "out_rows", self.n)
        indptr = np.empty(rows + 1, dtype=np.int64)
        indices = np.empty(max(nnz, 1), dtype=np.int32)
        data = np.empty(max(nnz, 1), dtype=np.float64)
        check(self._L.arcte_cuda_get_features(self._h, ptr(indptr), ptr(indices), ptr(data)))
        return indptr, indices[:nnz], data[:nnz]

    def features_into(self, indptr, indices, data):
        """Copy the assembled block into caller-owned host arrays (slices of larger arrays are
        fine): indptr int64[rows+1], indices int32[nnz], data float64[nnz]."""
        rows = getattr(self, "out_rows", self.n)
        assert indptr.dtype == np.int64 and indptr.size == rows + 1 and indptr.flags.c_contiguous
        assert indices.dtype == np.int32 and indices.size == self.out_nnz and indices.flags.c_contiguous
        assert data is None or (data.dtype == np.float64 and data.size == self.out_nnz and data.flags.c_contiguous)
        check(self._L.arcte_cuda_get_features(self._h, ptr(indptr), ptr(indices) if self.out_nnz else None,
                                              ptr(data) if (self.out_nnz and data is not None) else None))

    def fetch_block(self, indptr, indices, data, values_are_ones=False, n_threads=0):
        """Copy the assembled block into caller-owned plain numpy arrays (slices of larger arrays are fine):
        indptr int64[rows+1], indices int32[nnz], data float64[nnz] (each may be None).  With values_are_ones
        the values are not copied but written as 1.0 by the copy threads (the caller patches the 2.0s)."""
        rows = getattr(self, "out_rows", self.n)
        assert indptr is None or (indptr.dtype == np.int64 and indptr.size == rows + 1 and indptr.flags.c_contiguous)
        assert indices is None or (indices.dtype == np.int32 and indices.size == self.out_nnz and indices.flags.c_contiguous)
        assert data is None or (data.dtype == np.float64 and data.size == self.out_nnz and data.flags.c_contiguous)
        check(self._L.arcte_cuda_fetch_features(self._h, ptr(indptr), ptr(indices) if self.out_nnz else None,
                                                ptr(data) if self.out_nnz else None, int(bool(values_are_ones)),
                                                int(n_threads)))

    def fetch_block_to(self, pid, addr_indptr, addr_indices, addr_data, values_are_ones=False, n_threads=0):
        """fetch_block with the destination arrays in another process: `pid` and three virtual addresses of that
        process (0 = skip the array).  See arcte_cuda_fetch_features_to."""
        check(self._L.arcte_cuda_fetch_features_to(self._h, int(pid), C.c_void_p(addr_indptr or None),
                                                   C.c_void_p(addr_indices or None) if self.out_nnz else None,
                                                   C.c_void_p(addr_data or None) if self.out_nnz else None,
                                                   int(bool(values_are_ones)), int(n_threads)))

    def features(self):
        """The n x 2n CSR of arcte.py:683, index dtype chosen like scipy (int32 if it fits)."""
        nnz = self.out_nnz
        if getattr(self, "out_rows", self.n) != self.n:
            raise ArcteCudaError("features(): only a row block is assembled; use the distributed path")
        indptr = np.empty(self.n + 1, dtype=np.int64)
        indices = np.empty(nnz, dtype=np.int32)
        # every stored value is 1.0 (arcte.py:379-381, :676-679) except the identity entry of a self-loop row:
        # unless the values were changed on the device they are neither copied nor written (hostmem.ones:
        # copy-on-write mappings of a block of ones), only the 2.0 entries are patched in
        import os, sys, time
        t0 = time.perf_counter()
        if self._values_structural:
            data = hostmem.ones(nnz)
            self.fetch_block(indptr, indices, None)
        else:
            data = np.empty(nnz, dtype=np.float64)
            self.fetch_block(indptr, indices, data)
        t1 = time.perf_counter()
        if self._values_structural:
            self.patch_self_loops(data, indptr)
        if os.environ.get("ARCTE_CUDA_DEBUG"):
            print("[arcte] features: fetch %.1f ms (%.2f GB indices%s), self-loop patch %.1f ms"
                  % (1e3 * (t1 - t0), 4e-9 * nnz, " + %.2f GB of ones mapped" % (8e-9 * nnz) if self._values_structural else
                     " + %.2f GB values" % (8e-9 * nnz), 1e3 * (time.perf_counter() - t1)), file=sys.stderr)
        if max(2 * self.n, nnz) < 2 ** 31:
            indptr = indptr.astype(np.int32)
        else:
            indices = indices.astype(np.int64)
        X = sparse.csr_matrix((data, indices, indptr), shape=(self.n, 2 * self.n), copy=False)
        return X

    # -- after the path: column normalisation / community weighting ----------------
    @staticmethod
    def _csr_arrays(X):
        X = sparse.csr_matrix(X, dtype=np.float64)
        if not X.has_sorted_indices:
            X = X.copy()
            X.sort_indices()
        if X.shape[1] >= 2 ** 31:
            raise ValueError("matrices with 2^31 or more columns are not supported")
        return (X.shape, np.ascontiguousarray(X.indptr, dtype=np.int64),
                np.ascontiguousarray(X.indices, dtype=np.int32), np.ascontiguousarray(X.data, dtype=np.float64))

    def normalize_columns(self, features):
        """embedding/common.py:49-67 on any scipy sparse matrix; returns a new CSR."""
        shape, indptr, indices, data = self._csr_arrays(features)
        out = np.empty(data.size, np.float64)   # page-locked for large results
        check(self._L.arcte_cuda_normalize_columns(self._h, shape[0], shape[1], ptr(indptr),
                                                   ptr(indices) if data.size else None,
                                                   ptr(data) if data.size else None,
                                                   ptr(out) if data.size else None))
        idx_t = np.int32 if max(shape[1], data.size) < 2 ** 31 else np.int64
        return sparse.csr_matrix((out, indices.astype(idx_t, copy=False), indptr.astype(idx_t)), shape=shape)

    def normalize_features(self):
        """normalize_columns on the assembled feature matrix resident on the device, in place."""
        check(self._L.arcte_cuda_normalize_features(self._h))
        self._values_structural = False

    def self_loop_rows(self):
        """(rows, rank): rows whose diagonal entry is stored in the uploaded matrix (their identity
        entry is 2.0) and the number of stored entries before it in the row."""
        if self._loops is None:
            ip, ix = self._indptr, self._indices
            # stored diagonal entries: scipy's C++ csr_diagonal on the pattern (no O(nnz) temporaries in numpy)
            pattern = sparse.csr_matrix((np.ones(ix.size, dtype=np.int8), ix, ip), shape=(self.n, self.n), copy=False)
            rows = np.flatnonzero(pattern.diagonal()).astype(np.int64)
            # position of the diagonal inside each of those rows: only their entries are expanded
            lens = (ip[rows + 1] - ip[rows]).astype(np.int64)
            first = np.concatenate([[0], np.cumsum(lens)[:-1]]) if rows.size else np.zeros(0, dtype=np.int64)
            within = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(first, lens)
            hit = ix[np.repeat(ip[rows], lens) + within] == np.repeat(rows, lens)
            self._loops = (rows, within[hit])
        return self._loops

    def patch_self_loops(self, data, out_indptr, row_lo=0, row_hi=None):
        """Sets the identity entries of self-loop rows to 2.0 in a value array of ones (arcte.py:676-679:
        I + pattern(A) has 2.0 where A stores its diagonal).  out_indptr: row pointers of the block
        [row_lo, row_hi), starting at 0; the diagonal is the rank-th base-block column of its row."""
        row_hi = self.n if row_hi is None else row_hi
        rows, rank = self.self_loop_rows()
        m = (rows >= row_lo) & (rows < row_hi)
        if m.any():
            data[np.asarray(out_indptr)[rows[m] - row_lo].astype(np.int64) + rank[m]] = 2.0

    def chi2_contingency(self, X_train, Y):
        """embedding/community_weighting.py:11-45; Y = binarised label matrix (sparse)."""
        shape, indptr, indices, _ = self._csr_arrays(X_train)
        yshape, y_indptr, y_indices, y_data = self._csr_arrays(Y)
        if yshape[0] != shape[0]:
            raise ValueError("X_train and y_train have different numbers of rows")
        out = np.zeros((yshape[1], shape[1]), dtype=np.float64)
        check(self._L.arcte_cuda_chi2_contingency(self._h, shape[0], shape[1], ptr(indptr),
                                                  ptr(indices) if indices.size else None, yshape[1], ptr(y_indptr),
                                                  ptr(y_indices) if y_indices.size else None,
                                                  ptr(y_data) if y_data.size else None, ptr(out)))
        return out

    def peak_snr(self, contingency_matrix):
        """embedding/community_weighting.py:48-84; changes the matrix in place (nan -> 0)."""
        cm = contingency_matrix
        if not (isinstance(cm, np.ndarray) and cm.dtype == np.float64 and cm.ndim == 2 and cm.flags.c_contiguous):
            raise ValueError("contingency matrix must be a C-contiguous 2-D float64 array")
        w = np.zeros(cm.shape[1], dtype=np.float64)
        check(self._L.arcte_cuda_peak_snr(self._h, cm.shape[0], cm.shape[1], ptr(cm), ptr(w)))
        return w

    def chi2_psnr_weights(self, X_train, Y):
        """chi2_contingency followed by peak_snr with the K x F matrix kept in HBM."""
        shape, indptr, indices, _ = self._csr_arrays(X_train)
        yshape, y_indptr, y_indices, y_data = self._csr_arrays(Y)
        if yshape[0] != shape[0]:
            raise ValueError("X_train and y_train have different numbers of rows")
        w = np.zeros(shape[1], dtype=np.float64)
        check(self._L.arcte_cuda_chi2_psnr_weights(self._h, shape[0], shape[1], ptr(indptr),
                                                   ptr(indices) if indices.size else None, yshape[1], ptr(y_indptr),
                                                   ptr(y_indices) if y_indices.size else None,
                                                   ptr(y_data) if y_data.size else None, ptr(w)))
        return w

    def community_weighting(self, X, community_weights):
        """embedding/community_weighting.py:87-125 for one matrix; returns a new CSR."""
        shape, indptr, indices, data = self._csr_arrays(X)
        w = np.ascontiguousarray(community_weights, dtype=np.float64)
        if w.size != shape[1]:
            raise ValueError("one community weight per column is required")
        out_indptr = np.zeros(shape[0] + 1, dtype=np.int64)
        out_indices = np.empty(max(data.size, 1), np.int32)
        out_data = np.empty(max(data.size, 1), np.float64)
        nnz = C.c_int64()
        check(self._L.arcte_cuda_community_weighting(self._h, shape[0], shape[1], ptr(indptr),
                                                     ptr(indices) if data.size else None,
                                                     ptr(data) if data.size else None, ptr(w) if w.size else None,
                                                     ptr(out_indptr), ptr(out_indices), ptr(out_data), C.byref(nnz)))
        k = nnz.value
        idx_t = np.int32 if max(shape[1], k) < 2 ** 31 else np.int64
        return sparse.csr_matrix((out_data[:k], out_indices[:k].astype(idx_t, copy=False), out_indptr.astype(idx_t)),
                                 shape=shape)

    # -- feature matrix resident across the folds of an experiment ---------------------------------
    def store_features(self, features=None):
        """Keep `features` (any scipy sparse matrix) in HBM for weighted_fold(); None adopts the
        matrix the last assemble() (+ normalize_features()) left on the device."""
        if features is None:
            check(self._L.arcte_cuda_store_assembled(self._h))
            self.stored_shape = (self.n, 2 * self.n)
            return
        shape, indptr, indices, data = self._csr_arrays(features)
        check(self._L.arcte_cuda_store_features(self._h, shape[0], shape[1], ptr(indptr),
                                                ptr(indices) if data.size else None, ptr(data) if data.size else None))
        self.stored_shape = shape

    def weighted_fold(self, train, test, Y_train):
        """X[train], X[test] of the stored matrix, weighted by the chi2 / peak-SNR weights of the
        training block (utility.py:94-104), everything on the device.  Y_train: binarised labels of
        the training rows.  Returns (X_train_weighted, X_test_weighted) as scipy CSR."""
        train = np.ascontiguousarray(train, dtype=np.int64)
        test = np.ascontiguousarray(test, dtype=np.int64)
        yshape, y_indptr, y_indices, y_data = self._csr_arrays(Y_train)
        if yshape[0] != train.size:
            raise ValueError("one label row per training row is required")
        a, b = C.c_int64(), C.c_int64()
        check(self._L.arcte_cuda_weighted_fold(self._h, train.size, ptr(train) if train.size else None, test.size,
                                               ptr(test) if test.size else None, yshape[1], ptr(y_indptr),
                                               ptr(y_indices) if y_indices.size else None,
                                               ptr(y_data) if y_data.size else None, C.byref(a), C.byref(b)))
        out = []
        for which, rows, nnz in ((0, train.size, a.value), (1, test.size, b.value)):
            indptr = np.zeros(rows + 1, dtype=np.int64)
            indices = np.empty(max(nnz, 1), np.int32)
            data = np.empty(max(nnz, 1), np.float64)
            check(self._L.arcte_cuda_get_fold(self._h, which, ptr(indptr), ptr(indices), ptr(data)))
            idx_t = np.int32 if max(self.stored_shape[1], nnz) < 2 ** 31 else np.int64
            out.append(sparse.csr_matrix((data[:nnz], indices[:nnz].astype(idx_t, copy=False), indptr.astype(idx_t)),
                                         shape=(rows, self.stored_shape[1])))
        return out[0], out[1]

    def timer_start(self):
        check(self._L.arcte_cuda_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        check(self._L.arcte_cuda_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(self._L.arcte_cuda_flush_l2(self._h))

    def stats(self):
        st = _lib.Stats()
        check(self._L.arcte_cuda_get_stats(self._h, C.byref(st)))
        return st.as_dict()


_ENGINES = {}


def get_engine(device=0):
    """Process-wide engine per GPU: keeps the walk-state pool allocated between calls."""
    e = _ENGINES.get(device)
    if e is None or e._h is None:
        e = Engine(device)
        _ENGINES[device] = e
    name = _resolve_schedule()
    e.auto_schedule = name == "auto"
    if not e.auto_schedule and e.schedule != _SCHEDULES[name]:
        e.set_schedule(_SCHEDULES[name])
    return e


def host_write_to(pid, address, array):
    """Copy a contiguous numpy array to virtual address `address` of process `pid` (process_vm_writev)."""
    a = np.ascontiguousarray(array)
    check(_lib.load().arcte_cuda_host_write_to(int(pid), C.c_void_p(int(address)), ptr(a), a.nbytes))


def advise_huge(array):
    """Ask for huge pages for a (yet untouched) numpy array of this process."""
    if array.nbytes:
        check(_lib.load().arcte_cuda_host_advise_huge(ptr(array), array.nbytes))


def device_count():
    """Number of CUDA devices visible to the library (CUDA runtime, not torch)."""
    n = C.c_int(0)
    check(_lib.load().arcte_cuda_device_count(C.byref(n)))
    return n.value
