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


# ------------------------------------------------------------------------------------------------
# arcte_distributed() end to end on CPU: the host flow of one rank (side thread for the self-loop rows, size
# exchange, both ways home, patching of the 2.0 diagonals) around an engine whose device work is done by the oracle
# ------------------------------------------------------------------------------------------------
_ORACLE_LOCK = __import__("threading").Lock()   # the oracle library is not re-entrant; the in-process path calls from threads


def _oracle_engine(rank, world):
    from oracle import arcte_oracle as O
    from reveal_graph_embedding_b200 import distributed as ardist
    from reveal_graph_embedding_b200.engine import Engine, host_write_to

    class OracleEngine(Engine):   # patch_self_loops / self_loop_rows are the product's own (pure numpy)
        def __init__(self):
            self._values_structural = True
            self.calls = []

        def set_graph(self, A, canonical=False):
            self.A, self.n, self.nnz, self._loops = A, int(A.shape[0]), int(A.nnz), None
            self._indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
            self._indices = np.ascontiguousarray(A.indices, dtype=np.int32)
            self.calls.append("set_graph")

        def comm_info(self):
            return (world, rank, 0)

        def extract(self, rule, rho, eps, shard_rank=0, shard_count=1):
            assert (shard_rank, shard_count) == (rank, world)
            self.rule, self.rho, self.eps = rule, rho, eps
            self.calls.append("extract")

        def exchange_assemble(self):
            with _ORACLE_LOCK:
                return self._exchange_assemble()

        def _exchange_assemble(self):
            g = O.Graph(self.A)
            seeds = g.seeds()
            sds, segs, mems = [], [], []
            for r in range(world):   # what the all-to-all delivers: every shard's communities
                mine = seeds[list(ardist.shard_positions(seeds.size, r, world))]
                sd, seg, mem, eff, st = O.extract(g, self.rule, self.rho, self.eps, mine, 1)
                sds.append(sd), segs.append(seg), mems.append(mem)
            X = O.assemble(g, np.concatenate(sds), np.concatenate(segs), np.concatenate(mems))
            lo, hi = ardist.row_range(self.n, rank, world)
            self.blk = X[lo:hi].tocsr()
            self.out_nnz, self.out_rows = int(self.blk.nnz), hi - lo
            self.calls.append("exchange_assemble")
            return self.out_nnz

        def fetch_block(self, indptr, indices, data, values_are_ones=False, n_threads=0):
            if indptr is not None:
                indptr[:] = self.blk.indptr
            if indices is not None:
                indices[:] = self.blk.indices
            if data is not None:
                data[:] = 1.0 if values_are_ones else self.blk.data

        def fetch_block_to(self, pid, a_ip, a_idx, a_dat, values_are_ones=False, n_threads=0):
            assert not a_ip and not a_dat   # structural result: only the column indices travel
            host_write_to(pid, a_idx, self.blk.indices.astype(np.int32))

    return OracleEngine()


def _dist_worker(rank, world, port, name, out_dir, no_remote, everywhere):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if no_remote:
        os.environ["ARCTE_CUDA_NO_REMOTE_WRITES"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import EPS, RHO, load_golden
        from reveal_graph_embedding_b200 import distributed as ardist
        A, z = load_golden(name)
        eng = _oracle_engine(rank, world)
        X = ardist.arcte_distributed(A, 0, RHO, EPS, engine=eng, all_ranks=everywhere)
        assert eng.calls == ["set_graph", "extract", "exchange_assemble"]
        assert eng._loops is not None          # found by the side thread (or, failing that, by the patch)
        if rank == 0 or everywhere:
            np.savez(os.path.join(out_dir, "rank%d.npz" % rank), indptr=X.indptr, indices=X.indices, data=X.data)
        else:
            assert X is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,no_remote,everywhere",
                         [(2, "edgecases160", False, False), (3, "edgecases160", True, False),
                          (2, "ba300", False, True), (3, "planted419", False, False)])
def test_arcte_distributed_host_flow_with_an_oracle_engine(tmp_path, world, name, no_remote, everywhere):
    """edgecases160 has self loops: the 2.0 diagonals are patched into a value array of ones on the way home."""
    from helpers import golden_features, load_golden
    port = _free_port()
    mp.spawn(_dist_worker, args=(world, port, name, str(tmp_path), no_remote, everywhere), nprocs=world, join=True)
    A, z = load_golden(name)
    want = golden_features(z, 0, A.shape[0])
    ranks = range(world) if everywhere else [0]
    for r in ranks:
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["indptr"], want.indptr)
        assert np.array_equal(got["indices"], want.indices)
        assert np.array_equal(got["data"], want.data)


@pytest.mark.parametrize("n_gpus,name", [(2, "edgecases160"), (3, "ba300")])
def test_in_process_multi_gpu_host_flow_with_oracle_engines(monkeypatch, n_gpus, name):
    """One process driving several GPUs (embedding/arcte/arcte.py): per-GPU threads, one self-loop scan shared by
    all engines, every engine's row block fetched into its slice of the result, 2.0 diagonals patched per block."""
    from helpers import EPS, RHO, golden_features, load_golden
    import reveal_graph_embedding_b200.embedding.arcte.arcte as mod
    engines = [_oracle_engine(r, n_gpus) for r in range(n_gpus)]
    scans = []
    for e in engines:
        real = e.self_loop_rows
        e.self_loop_rows = (lambda real=real, e=e: (scans.append(1) if e._loops is None else None, real())[1])
    monkeypatch.setattr(mod, "get_engine", lambda d=0: engines[d])
    monkeypatch.setattr(mod, "device_count", lambda: n_gpus)
    monkeypatch.setattr(mod.Engine, "comm_init_all", staticmethod(lambda engs: None))
    A, z = load_golden(name)
    X = mod.arcte(A, RHO, EPS)
    want = golden_features(z, 0, A.shape[0])
    assert np.array_equal(X.indptr, want.indptr) and np.array_equal(X.indices, want.indices)
    assert np.array_equal(X.data, want.data)
    assert all(e.calls == ["set_graph", "extract", "exchange_assemble"] for e in engines)
    assert len(scans) == 1 and all(e._loops is engines[0]._loops for e in engines)   # scanned once, shared
