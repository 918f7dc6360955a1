"""Seed sharding across GPUs, one process per GPU (torchrun / torch.distributed).

The reference fans seeds out to a multiprocessing pool and sums the workers' matrices on the parent
(arcte.py:650-673).  Here every rank holds the whole graph, walks the seeds at positions rank,
rank + world, ... of the degree-sorted seed list (roundrobin_chunks, arcte.py:19-23) and owns a contiguous
block of rows of the result.  The ONE exchange step runs inside libarcte_cuda (csrc/exchange.cu): every
community is split by destination row block on the device and the pieces cross NVLink once, in a single
grouped NCCL send/recv; each rank then assembles its own rows.  torch.distributed is used for the
plumbing only: the process group tells the ranks apart, carries the 128-byte NCCL id once and a handful of
integers per call.

The rows come home through every rank's own PCIe link: rank 0 allocates ordinary numpy arrays and every rank
streams its row block straight into them (process_vm_writev from the pinned staging slots of csrc/hostcopy.cu);
rank 0 returns the matrix, the other ranks None.  Where one process may not write into another, or every rank
wants the matrix (ARCTE_CUDA_RESULT_ON_ALL_RANKS=1), the arrays live in /dev/shm and every rank maps them.
"""
import os
import sys
import uuid

import numpy as np


def _dist():
    t = sys.modules.get("torch")
    if t is None:
        return None
    d = t.distributed
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        return d
    return None


def is_active():
    """True inside an initialised torch.distributed job with more than one rank."""
    return _dist() is not None


def shard_positions(n_seeds, rank, world):
    """Seed-list positions of one rank: rank, rank + world, ... (arcte.py:19-23)."""
    return range(rank, n_seeds, world)


def row_range(n, rank, world):
    """Rows of the feature matrix rank `rank` assembles: equal contiguous blocks."""
    return (n * rank) // world, (n * (rank + 1)) // world


def result_on_all_ranks():
    """ARCTE_CUDA_RESULT_ON_ALL_RANKS=1: every rank returns the matrix (all of them map the same shared
    memory).  Default: rank 0 only, the others return None."""
    return os.environ.get("ARCTE_CUDA_RESULT_ON_ALL_RANKS", "0") == "1"


# ------------------------------------------------------------------------------------------------
# host-side collectives: a few integers / bytes per call, over whatever backend the group has
# ------------------------------------------------------------------------------------------------
def _device_for_group(dist):
    import torch
    backend = dist.get_backend()
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_int64(values):
    """values: sequence of ints, the same length on every rank -> int64 array [world, len]."""
    import torch
    dist = _dist()
    dev = _device_for_group(dist)
    mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    out = torch.empty(dist.get_world_size() * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().reshape(dist.get_world_size(), -1)


def broadcast_bytes(payload, n_bytes, src=0):
    """`payload` (bytes of length n_bytes on rank src, ignored elsewhere) -> the same bytes on every rank."""
    import torch
    dist = _dist()
    dev = _device_for_group(dist)
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tolist())


# ------------------------------------------------------------------------------------------------
# the result in shared memory: one file per array under /dev/shm, mapped by every rank
# ------------------------------------------------------------------------------------------------
SHM_DIR = os.environ.get("ARCTE_CUDA_SHM_DIR", "/dev/shm")


def block_offsets(block_nnz):
    """Exclusive prefix of the per-rank stored-entry counts (+ total)."""
    return np.concatenate([[0], np.cumsum(np.asarray(block_nnz, dtype=np.int64))])


class SharedResult:
    """indptr int64[n+1], indices int32[total], data float64[total] backed by files in SHM_DIR.
    Rank 0 creates them, everybody maps them; the files are unlinked as soon as every rank has them open
    (the memory lives as long as a mapping does)."""

    def __init__(self, tag, n, total, create):
        self.paths = [os.path.join(SHM_DIR, "arcte_%s_%s" % (tag, k)) for k in ("indptr", "indices", "data")]
        shapes = [((n + 1,), np.int64), ((max(total, 1),), np.int32), ((max(total, 1),), np.float64)]
        mode = "w+" if create else "r+"
        self.arrays = [np.memmap(p, dtype=dt, mode=mode, shape=sh) for p, (sh, dt) in zip(self.paths, shapes)]
        self.n, self.total = n, total

    def unlink(self):
        for p in self.paths:
            try:
                os.unlink(p)
            except FileNotFoundError:
                pass

    def csr(self):
        import scipy.sparse as sparse
        indptr, indices, data = (np.asarray(a) for a in self.arrays)
        indices, data = indices[:self.total], data[:self.total]
        if max(2 * self.n, self.total) < 2 ** 31:
            indptr = indptr.astype(np.int32)
        else:
            indices = indices.astype(np.int64)
        return sparse.csr_matrix((data, indices, indptr), shape=(self.n, 2 * self.n), copy=False)


def place_block(result, rank, world, offsets, blk_indptr):
    """Row pointers of one rank's block into the shared indptr: neighbouring blocks write the same value
    at the row they share."""
    lo, hi = row_range(result.n, rank, world)
    result.arrays[0][lo:hi + 1] = blk_indptr + offsets[rank]
    return lo, hi


# ------------------------------------------------------------------------------------------------
def ensure_communicator(eng):
    """NCCL communicator of this engine's context over all ranks of the default process group (created on
    first use; the id comes from rank 0's library and travels through the process group)."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    if eng.comm_info()[:2] == (world, rank):
        return
    uid = eng.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    eng.comm_init(world, rank, uid)


_remote_ok = {}   # world size -> bool: can every rank of this job write into rank 0's memory?


def remote_writes_work(rank, world):
    """One probe per job: every rank writes 8 bytes into a buffer of rank 0 with process_vm_writev (needs ptrace
    permission on rank 0: same user and Yama scope <= 1, the usual case for the ranks of one torchrun)."""
    from .engine import host_write_to
    if os.environ.get("ARCTE_CUDA_NO_REMOTE_WRITES") == "1":
        return False
    key = world
    if key not in _remote_ok:
        probe = np.zeros(world, dtype=np.int64)
        info = all_gather_int64([os.getpid(), probe.ctypes.data])
        ok = 1
        try:
            host_write_to(int(info[0, 0]), int(info[0, 1]) + 8 * rank, np.array([rank + 1], dtype=np.int64))
        except Exception:
            ok = 0
        _dist().barrier()
        if rank == 0:
            ok = int(ok and np.array_equal(probe, np.arange(1, world + 1)))
        _remote_ok[key] = bool(all_gather_int64([ok])[:, 0].min())
    return _remote_ok[key]


def arcte_distributed(A, rule, rho_eff, epsilon, engine=None, upload=True, all_ranks=None):
    """One rank's share of arcte() inside a torch.distributed job.  Rank 0 returns the matrix, the other
    ranks None (all_ranks / ARCTE_CUDA_RESULT_ON_ALL_RANKS=1: everybody, through shared memory).

    The rows come home through every rank's own PCIe link.  Default: rank 0 allocates ordinary numpy arrays
    and every rank streams its row block straight into them (process_vm_writev from its pinned staging slots,
    csrc/hostcopy.cu); the 1.0 values of a rank's rows are written the same way, never copied from the device.
    Where writing into another process is not permitted, or every rank wants the matrix, the arrays live in
    /dev/shm instead (SharedResult)."""
    import scipy.sparse as sparse
    from . import hostmem
    from .engine import advise_huge, get_engine, host_write_to
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    if world > 16:
        raise RuntimeError("arcte: at most 16 ranks (one box); got %d" % world)
    if engine is None:
        local = os.environ.get("LOCAL_RANK")
        if local is None:
            import torch
            local = torch.cuda.current_device()
        engine = get_engine(int(local))
    eng = engine
    import time
    dbg = os.environ.get("ARCTE_CUDA_DEBUG") and rank == 0
    tm = [("start", time.perf_counter())]

    def mark(name):
        if dbg:
            tm.append((name, time.perf_counter()))
    if upload:
        eng.set_graph(A, canonical=True)
    if all_ranks is None:
        all_ranks = result_on_all_ranks()
    ensure_communicator(eng)
    mark("set_graph")
    n = eng.n
    # the self-loop rows (identity entry 2.0, arcte.py:676-679) depend on the input only: found on a side thread
    # while the device walks (a failure there is met again, and raised, by patch_self_loops)
    import threading

    def _loops():
        try:
            eng.self_loop_rows()
        except Exception:
            pass
    side = threading.Thread(target=_loops)
    side.start()
    try:
        eng.extract(rule, rho_eff, epsilon, shard_rank=rank, shard_count=world)
    finally:
        side.join()
    mark("extract")
    nnz = eng.exchange_assemble()
    mark("exchange+assemble")
    structural = eng._values_structural
    lo, hi = row_range(n, rank, world)
    threads = max(2, (os.cpu_count() or 8) // world)

    if not all_ranks and remote_writes_work(rank, world):
        offsets = block_offsets(all_gather_int64([nnz])[:, 0])
        total = int(offsets[-1])
        addr = [0, 0, 0, 0]
        if rank == 0:
            indptr = np.empty(n + 1, dtype=np.int64)
            indices = np.empty(max(total, 1), dtype=np.int32)
            advise_huge(indices)
            # values: all ones but for the patched diagonals -> copy-on-write mappings of a block of ones
            data = hostmem.ones(max(total, 1)) if structural else np.empty(max(total, 1), dtype=np.float64)
            addr = [os.getpid(), indptr.ctypes.data, indices.ctypes.data, data.ctypes.data]
        pid, a_ptr, a_idx, a_dat = (int(x) for x in all_gather_int64(addr)[0])
        mark("sizes+alloc")
        o0 = int(offsets[rank])
        ip = np.empty(hi - lo + 1, dtype=np.int64)
        eng.fetch_block(ip, None, None)                       # this block's row pointers (small), made global below
        eng.fetch_block_to(pid, 0, a_idx + 4 * o0, 0 if structural else a_dat + 8 * o0, n_threads=threads)
        host_write_to(pid, a_ptr + 8 * lo, ip + offsets[rank])   # neighbouring blocks write the same value at their common row
        mark("own block home")
        dist.barrier()
        mark("barrier (slowest rank)")
        if rank != 0:
            return None
        if structural:
            eng.patch_self_loops(data, indptr)                # rank 0 holds the same graph: all rows at once
        indices, data = indices[:total], data[:total]
        if max(2 * n, total) < 2 ** 31:
            indptr = indptr.astype(np.int32)
        else:
            indices = indices.astype(np.int64)
        X = sparse.csr_matrix((data, indices, indptr), shape=(n, 2 * n), copy=False)
        mark("patch+csr")
        if dbg:
            print("[arcte] rank 0 of %d: " % world + ", ".join("%s %.1f ms" % (b[0], 1e3 * (b[1] - a[1])) for a, b in zip(tm[:-1], tm[1:])),
                  file=sys.stderr)
        return X

    # ---- shared-memory variant ----
    tag_bits = uuid.uuid4().int & ((1 << 62) - 1) if rank == 0 else 0
    table = all_gather_int64([nnz, tag_bits])
    offsets = block_offsets(table[:, 0])
    total = int(offsets[-1])
    tag = "%016x_%d" % (int(table[0, 1]), os.getuid())
    res = None
    if rank == 0:
        res = SharedResult(tag, n, total, create=True)
    dist.barrier()
    if rank != 0:
        res = SharedResult(tag, n, total, create=False)
    o0, o1 = int(offsets[rank]), int(offsets[rank + 1])
    ip = np.empty(hi - lo + 1, dtype=np.int64)
    indices, data = np.asarray(res.arrays[1]), np.asarray(res.arrays[2])
    eng.fetch_block(ip, indices[o0:o1], data[o0:o1], values_are_ones=structural, n_threads=threads)
    if structural:
        eng.patch_self_loops(data[o0:o1], ip, lo, hi)
    place_block(res, rank, world, offsets, ip)
    dist.barrier()          # every block is in place (the mappings share pages: nothing to flush)
    if rank == 0:
        res.unlink()
    return res.csr() if (rank == 0 or all_ranks) else None
