"""Multi-GPU paths (skipped unless the box exposes >= 2 GPUs): one process driving several contexts and
one process per GPU (torchrun), both through the in-library NCCL exchange (csrc/exchange.cu)."""
import os
import subprocess
import sys

import pytest

from helpers import EPS, GOLDEN_NAMES, RHO, assert_csr_identical, golden_features, load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    from reveal_graph_embedding_b200.engine import device_count
    return device_count()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_in_process_multi_gpu_identical(name):
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    A, z = load_golden(name)
    X = arcte(A, RHO, EPS, number_of_threads=n_gpus())
    assert_csr_identical(X, golden_features(z, 0, A.shape[0]))


@pytest.mark.parametrize("everywhere", ["1", "0"])
def test_torchrun_nccl_identical(everywhere):
    """everywhere=1: every rank returns the matrix; 0 (default): rank 0 only, others None."""
    g = n_gpus()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g),
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, DIST_CHECK_ALL=everywhere))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "dist_check ok" in out.stdout


def test_in_process_multi_gpu_large_result_with_self_loops(oracle):
    """Large result over several GPUs: the value array is written as ones by the copy threads and the 2.0
    entries of self-loop diagonals are patched across the per-GPU row blocks."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import scipy.sparse as sparse
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    A = graphs.barabasi_albert(20000, 3, seed=2).tolil()
    for i in (0, 5, 777, 19999):
        A[i, i] = 1.0
    A = sparse.csr_matrix(A)
    want = oracle.arcte(A, RHO, EPS, 8)
    for rep in range(2):
        X = arcte(A, RHO, EPS, number_of_threads=n_gpus())
        assert_csr_identical(X, want)
        X.data[:] = -1.0   # the result arrays are the caller's: nothing of them is reused by the next call
