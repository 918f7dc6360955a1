"""torchrun target: every rank runs arcte() on a golden graph inside an NCCL process group
and checks the full matrix bit-for-bit against the reference fixture.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from helpers import EPS, RHO, GOLDEN_NAMES, golden_features, load_golden, assert_csr_identical

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ["ARCTE_CUDA_RESULT_ON_ALL_RANKS"] = "1" if os.environ.get("DIST_CHECK_ALL", "1") == "1" else "0"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte, arcte_with_lazy_pagerank
for name in GOLDEN_NAMES:
    A, z = load_golden(name)
    everywhere = os.environ["ARCTE_CUDA_RESULT_ON_ALL_RANKS"] == "1"
    X = arcte(A, RHO, EPS)
    assert (X is not None) == (everywhere or dist.get_rank() == 0)
    if X is not None:
        assert_csr_identical(X, golden_features(z, 0, A.shape[0]))
    if "X2_data" in z:
        X = arcte_with_lazy_pagerank(A, RHO, EPS)
        if X is not None:
            assert_csr_identical(X, golden_features(z, 2, A.shape[0]))
# a larger result with self loops: the values are written on the host, the 2.0 diagonals patched per row block
if os.environ["ARCTE_CUDA_RESULT_ON_ALL_RANKS"] == "0":
    import scipy.sparse as sparse
    from oracle import arcte_oracle
    from reveal_graph_embedding_b200 import graphs
    B = graphs.barabasi_albert(20000, 3, seed=2).tolil()
    for i in (0, 5, 777, 19999):
        B[i, i] = 1.0
    B = sparse.csr_matrix(B)
    want = arcte_oracle.arcte(B, RHO, EPS, 8) if dist.get_rank() == 0 else None
    for rep in range(2):
        X = arcte(B, RHO, EPS)
        if dist.get_rank() == 0:
            assert_csr_identical(X, want)
            X.data[:] = -1.0
        del X
        dist.barrier()
dist.barrier()
if dist.get_rank() == 0:
    print("dist_check ok: world=%d, %d graphs" % (dist.get_world_size(), len(GOLDEN_NAMES)))
dist.destroy_process_group()
