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


def arcte_distributed(A, rule, rho_eff, epsilon, engine=None, upload=True, all_ranks=None):
    """One rank's share of arcte(): walk this rank's seeds, all-gather the segments, and
    (rank 0, or every rank when all_ranks) assemble and return the complete n x 2n CSR."""
    import torch
    import torch.distributed as dist
    from .engine import get_engine
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device()
    eng = engine or get_engine(device)
    if upload:
        eng.set_graph(A, canonical=True)
    eng.extract(rule, rho_eff, epsilon, shard_rank=rank, shard_count=world)
    parts, keep = gather_engine_segments(eng)
    if all_ranks is None:
        all_ranks = result_on_all_ranks()
    if rank != 0 and not all_ranks:
        del keep
        return None
    eng.assemble(parts)
    del keep
    return eng.features()


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
