"""world_size-2 (and 3) gloo tests of the host-side sharding logic (no GPU):
round-robin seed shards (arcte.py:19-23) + the all-gather of ragged segment arrays used
to join the per-GPU results.  Segments are produced by the oracle here; on the GPU box the
same functions move device tensors over NCCL (tests/test_gpu_parity.py covers the CUDA side)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import EPS, RHO, load_golden
        from oracle import arcte_oracle as O
        from reveal_graph_embedding_b200 import distributed as ardist
        assert ardist.is_active()
        A, z = load_golden(name)
        g = O.Graph(A)
        seeds = g.seeds()
        mine = seeds[list(ardist.shard_positions(seeds.size, rank, world))]
        sd, seg, mem, eff, st = O.extract(g, 0, RHO, EPS, mine, 1)
        off = np.concatenate([[0], np.cumsum(seg)[:-1]]).astype(np.int64) if seg.size else np.zeros(0, np.int64)
        parts = ardist.allgather_segments(torch.from_numpy(sd.astype(np.int32)),
                                          torch.from_numpy(seg.astype(np.int32)),
                                          torch.from_numpy(off), torch.from_numpy(mem.astype(np.int32)))
        assert len(parts) == world
        # every rank rebuilds the full matrix from the gathered parts
        all_seed = np.concatenate([p[0].numpy() for p in parts]).astype(np.int64)
        all_cnt = np.concatenate([p[1].numpy() for p in parts]).astype(np.int64)
        all_mem = np.concatenate([np.concatenate([p[3].numpy()[o:o + c] for o, c in zip(p[2].numpy(), p[1].numpy())] or
                                                 [np.zeros(0, np.int32)]) for p in parts]).astype(np.int32)
        X = O.assemble(g, all_seed, all_cnt, all_mem)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), indptr=X.indptr, indices=X.indices, data=X.data,
                 n_mine=mine.size)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "ba300"), (3, "edgecases160"), (2, "planted419")])
def test_gloo_sharded_extraction_matches_reference(tmp_path, world, name):
    from helpers import golden_features, load_golden
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    A, z = load_golden(name)
    want = golden_features(z, 0, A.shape[0])
    total = 0
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["indptr"], want.indptr)
        assert np.array_equal(got["indices"], want.indices)
        assert np.array_equal(got["data"], want.data)
        total += int(got["n_mine"])
    assert total == z["seeds"].size  # shards partition the seed list


def test_shard_positions_round_robin():
    from reveal_graph_embedding_b200.distributed import shard_positions
    for n in (0, 1, 7, 8, 9):
        for world in (1, 2, 3, 8, 16):
            got = sorted(p for r in range(world) for p in shard_positions(n, r, world))
            assert got == list(range(n))
            for r in range(world):
                assert list(shard_positions(n, r, world)) == list(range(n))[r::world]  # arcte.py:19-23 islice
