"""world_size-2 (and 3) gloo tests of the host side of the one-process-per-GPU path (no GPU):
round-robin seed shards (arcte.py:19-23), the integer / byte collectives that carry the NCCL id and the block
sizes, and the shared-memory result every rank writes its row block into.  The row blocks are produced by
the oracle here; on the GPU box libarcte_cuda produces them (walk, in-library NCCL exchange, assembly:
tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
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
        n = A.shape[0]
        g = O.Graph(A)
        seeds = g.seeds()
        # what the ranks' walks produce together: every rank's shard, joined (the GPU path joins them with the
        # NCCL all-to-all); this rank then owns the rows of its block
        sds, segs, mems = [], [], []
        for r in range(world):
            mine = seeds[list(ardist.shard_positions(seeds.size, r, world))]
            sd, seg, mem, eff, st = O.extract(g, 0, RHO, EPS, mine, 1)
            sds.append(sd), segs.append(seg), mems.append(mem)
        X = O.assemble(g, np.concatenate(sds), np.concatenate(segs), np.concatenate(mems))
        lo, hi = ardist.row_range(n, rank, world)
        blk = X[lo:hi]
        # the collectives of arcte_distributed
        uid = bytes(range(128)) if rank == 0 else None
        assert ardist.broadcast_bytes(uid, 128, src=0) == bytes(range(128))
        table = ardist.all_gather_int64([blk.nnz, 4242 if rank == 0 else 0])
        assert table.shape == (world, 2) and table[rank, 0] == blk.nnz
        offsets = ardist.block_offsets(table[:, 0])
        tag = "test%d_%d" % (int(table[0, 1]), port)
        res = ardist.SharedResult(tag, n, int(offsets[-1]), create=True) if rank == 0 else None
        dist.barrier()
        if rank != 0:
            res = ardist.SharedResult(tag, n, int(offsets[-1]), create=False)
        o0, o1 = int(offsets[rank]), int(offsets[rank + 1])
        res.arrays[1][o0:o1] = blk.indices
        res.arrays[2][o0:o1] = blk.data
        ardist.place_block(res, rank, world, offsets, blk.indptr.astype(np.int64))
        dist.barrier()
        if rank == 0:
            res.unlink()
            assert not any(os.path.exists(p) for p in res.paths)
        # the default way home: every rank writes its block into ordinary memory of rank 0 (process_vm_writev)
        if ardist.remote_writes_work(rank, world):
            from reveal_graph_embedding_b200.engine import host_write_to
            direct = np.zeros(int(offsets[-1]), dtype=np.int32)
            info = ardist.all_gather_int64([os.getpid(), direct.ctypes.data])
            host_write_to(int(info[0, 0]), int(info[0, 1]) + 4 * o0, blk.indices.astype(np.int32))
            dist.barrier()
            if rank == 0:
                assert np.array_equal(direct, np.asarray(res.arrays[1])[:int(offsets[-1])])
        got = res.csr()   # every rank maps the same memory
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), indptr=got.indptr, indices=got.indices, data=got.data)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "ba300"), (3, "edgecases160"), (2, "planted419")])
def test_gloo_row_blocks_in_shared_memory_match_reference(tmp_path, world, name):
    from helpers import golden_features, load_golden
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    A, z = load_golden(name)
    want = golden_features(z, 0, A.shape[0])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["indptr"], want.indptr)
        assert np.array_equal(got["indices"], want.indices)
        assert np.array_equal(got["data"], want.data)


def test_shard_positions_round_robin():
    from reveal_graph_embedding_b200.distributed import shard_positions
    for n in (0, 1, 7, 8, 9):
        for world in (1, 2, 3, 8, 16):
            got = sorted(p for r in range(world) for p in shard_positions(n, r, world))
            assert got == list(range(n))
            for r in range(world):
                assert list(shard_positions(n, r, world)) == list(range(n))[r::world]  # arcte.py:19-23 islice


def test_row_blocks_partition_the_rows_and_match_the_device_rule():
    """row_range() against a restatement of exchange.cu's dest_of(): every row has exactly one owner."""
    from reveal_graph_embedding_b200.distributed import row_range

    def dest_of(row, n, G):
        d = (row * G) // n
        while d + 1 < G and (n * (d + 1)) // G <= row:
            d += 1
        while d > 0 and (n * d) // G > row:
            d -= 1
        return d

    for n in (1, 2, 5, 7, 160, 419, 1000, 1138499):
        for G in (1, 2, 3, 7, 8, 16):
            owner = np.full(n, -1)
            for r in range(G):
                lo, hi = row_range(n, r, G)
                assert (owner[lo:hi] == -1).all()
                owner[lo:hi] = r
            assert (owner >= 0).all()
            rows = np.unique(np.concatenate([np.arange(min(n, 50)), np.arange(max(0, n - 50), n),
                                             np.random.default_rng(n + G).integers(0, n, 200)]))
            for x in rows:
                assert dest_of(int(x), n, G) == owner[x]
