"""Seed sharding across GPUs, one process per GPU (torchrun / torch.distributed).

The reference fans seeds out to a multiprocessing pool and sums the workers' matrices
on the parent (arcte.py:650-673).  Here every rank holds the whole graph, walks the
seeds at positions rank, rank + world, ... of the degree-sorted seed list
(roundrobin_chunks, arcte.py:19-23) and the per-rank member segments are joined by ONE
exchange step: an NCCL all-gather of (seed, count, offset, members).  torch is used for
the process group and the collective only; the arrays it moves are filled and consumed
by libarcte_cuda through raw device pointers.
"""
import sys


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


def allgather_segments(seg_seed, seg_count, seg_offset, members, group=None):
    """All-gather four 1-D tensors of rank-dependent length.

    seg_seed/seg_count: int32 [S_r]; seg_offset: int64 [S_r]; members: int32 [M_r].
    Returns a list over ranks of (seg_seed, seg_count, seg_offset, members) tensor views
    (on the same device as the inputs).  Sizes are exchanged first, then every array is
    padded to the largest rank's size so a single fixed-size all-gather per array moves it.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = seg_seed.device
    sizes = torch.tensor([seg_seed.numel(), members.numel()], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 2).cpu()
    s_max = max(int(all_sizes[:, 0].max()), 1)
    m_max = max(int(all_sizes[:, 1].max()), 1)

    def gather(x, width, dtype):
        send = torch.zeros(width, dtype=dtype, device=dev)
        send[:x.numel()] = x
        recv = torch.empty(world * width, dtype=dtype, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        return recv.view(world, width)

    g_seed = gather(seg_seed, s_max, torch.int32)
    g_count = gather(seg_count, s_max, torch.int32)
    g_off = gather(seg_offset, s_max, torch.int64)
    g_mem = gather(members, m_max, torch.int32)
    parts = []
    for r in range(world):
        S, M = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        parts.append((g_seed[r, :S], g_count[r, :S], g_off[r, :S], g_mem[r, :M]))
    return parts


def result_on_all_ranks():
    """ARCTE_CUDA_RESULT_ON_ALL_RANKS=1: every rank assembles and returns the matrix.
    Default: rank 0 only (the others return None) -- copying a multi-gigabyte matrix to
    the host once per rank would dominate the call."""
    import os
    return os.environ.get("ARCTE_CUDA_RESULT_ON_ALL_RANKS", "0") == "1"


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, address, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(address), False), "version": 2}


def _view(address, count, typestr, device):
    import torch
    if count == 0:
        dt = {"<i8": torch.int64, "<i4": torch.int32, "<f8": torch.float64}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(_DeviceArray(address, count, typestr), device=device)


def row_range(n, rank, world):
    """Rows of the feature matrix rank `rank` assembles: equal contiguous blocks."""
    return (n * rank) // world, (n * (rank + 1)) // world


def extract_and_concatenate(eng, rule, rho_eff, epsilon):
    """Steps 1-3 of the distributed call, everything staying in HBM.

    1. walk this rank's round-robin shard of the seeds (K2b-K4);
    2. ONE exchange of the walk results: NCCL all-gather of the member segments;
    3. every rank assembles the row block [n*r/G, n*(r+1)/G) of the feature matrix (K5 on 1/G
       of the entries); the blocks are concatenated on rank 0 with NCCL send/recv.
    Returns (indptr, indices, data, nnz) device tensors on rank 0, None on the other ranks.
    """
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", eng.device)
    eng.extract(rule, rho_eff, epsilon, shard_rank=rank, shard_count=world)
    parts, keep = gather_engine_segments(eng)
    n = eng.n
    lo, hi = row_range(n, rank, world)
    eng.assemble(parts, row_lo=lo, row_hi=hi)
    del keep
    p_indptr, p_indices, p_data, n_rows, nnz = eng.features_device()
    sizes = torch.tensor([nnz], dtype=torch.int64, device=dev)
    all_nnz = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_nnz, sizes)
    all_nnz = all_nnz.cpu().numpy()
    offsets = np.concatenate([[0], np.cumsum(all_nnz)])
    total = int(offsets[-1])
    blk_indptr = _view(p_indptr, n_rows + 1, "<i8", dev)
    blk_indices = _view(p_indices, nnz, "<i4", dev)
    blk_data = _view(p_data, nnz, "<f8", dev)
    torch.cuda.synchronize(dev)
    if rank != 0:
        ops = [dist.P2POp(dist.isend, blk_indptr, 0)]
        if nnz > 0:
            ops += [dist.P2POp(dist.isend, blk_indices, 0), dist.P2POp(dist.isend, blk_data, 0)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        torch.cuda.synchronize(dev)
        return None
    indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
    tmp_ptr = []
    ops = []
    for r in range(1, world):
        rlo, rhi = row_range(n, r, world)
        t = torch.empty(rhi - rlo + 1, dtype=torch.int64, device=dev)
        tmp_ptr.append((r, rlo, rhi, t))
        ops.append(dist.P2POp(dist.irecv, t, r))
        if all_nnz[r] > 0:
            ops.append(dist.P2POp(dist.irecv, indices[int(offsets[r]):int(offsets[r + 1])], r))
            ops.append(dist.P2POp(dist.irecv, data[int(offsets[r]):int(offsets[r + 1])], r))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    indptr[lo:hi + 1] = blk_indptr
    indices[:nnz] = blk_indices
    data[:nnz] = blk_data
    for r in reqs:
        r.wait()
    for r, rlo, rhi, t in tmp_ptr:
        indptr[rlo:rhi + 1] = t + int(offsets[r])
    torch.cuda.synchronize(dev)
    return indptr, indices, data, total


def arcte_distributed(A, rule, rho_eff, epsilon, engine=None, upload=True, all_ranks=None):
    """One rank's share of arcte() inside a torch.distributed job: see extract_and_concatenate.
    Rank 0 copies the matrix to the host and returns it; the other ranks return None.  With
    all_ranks (ARCTE_CUDA_RESULT_ON_ALL_RANKS=1) every rank assembles and returns the matrix."""
    import numpy as np
    import scipy.sparse as sparse
    import torch
    import torch.distributed as dist
    from . import hostmem
    from .engine import get_engine
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = engine or get_engine(torch.cuda.current_device())
    if upload:
        eng.set_graph(A, canonical=True)
    if all_ranks is None:
        all_ranks = result_on_all_ranks()
    if all_ranks:
        eng.extract(rule, rho_eff, epsilon, shard_rank=rank, shard_count=world)
        parts, keep = gather_engine_segments(eng)
        eng.assemble(parts)
        del keep
        return eng.features()
    out = extract_and_concatenate(eng, rule, rho_eff, epsilon)
    if out is None:
        return None
    indptr, indices, data, total = out
    n = eng.n
    # device -> host into pooled page-locked buffers
    h_indices = hostmem.empty(max(total, 1), np.int32)
    torch.from_numpy(h_indices).copy_(indices)
    h_indptr = indptr.cpu().numpy()
    # values: ones pre-filled on the host + self-loop diagonals patched, else a device-to-host copy
    h_data = hostmem.ones(max(total, 1))
    if h_data is not None:
        rows, rank_in_row = eng.self_loop_rows()
        if rows.size:
            h_data[h_indptr[rows].astype(np.int64) + rank_in_row] = 2.0
    else:
        h_data = hostmem.empty(max(total, 1), np.float64)
        torch.from_numpy(h_data).copy_(data)
    hostmem.start_pending()
    h_indices, h_data = h_indices[:total], h_data[:total]
    if max(2 * n, total) < 2 ** 31:
        h_indptr = h_indptr.astype(np.int32)
    else:
        h_indices = h_indices.astype(np.int64)
    return sparse.csr_matrix((h_data, h_indices, h_indptr), shape=(n, 2 * n), copy=False)


def gather_engine_segments(eng):
    """Copy the engine's device-resident segments into torch tensors, all-gather them over
    NCCL and return them as raw-pointer parts for Engine.assemble (plus the tensors that
    must stay alive until assemble returns)."""
    import torch
    dev = torch.device("cuda", eng.device)
    S, M = eng.n_segments, eng.n_members
    seg_seed = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
    seg_count = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
    seg_offset = torch.empty(max(S, 1), dtype=torch.int64, device=dev)
    members = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    eng.export_segments(seg_seed.data_ptr(), seg_count.data_ptr(), seg_offset.data_ptr(), members.data_ptr())
    gathered = allgather_segments(seg_seed[:S], seg_count[:S], seg_offset[:S], members[:M])
    torch.cuda.synchronize(dev)
    parts = [(int(a.numel()), int(d.numel()), a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr())
             for (a, b, c, d) in gathered]
    return parts, gathered
