"""Shared helpers for the parity tests."""
import os

import numpy as np
import scipy.sparse as sparse

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["ba300", "weighted200", "planted419", "edgecases160", "ba2000"]
RHO, EPS = 0.1, 1e-5


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = z["A_indptr"].size - 1
    A = sparse.csr_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    return A, z


def golden_features(z, rule, n):
    return sparse.csr_matrix((z["X%d_data" % rule], z["X%d_indices" % rule],
                              z["X%d_indptr" % rule]), shape=(n, 2 * n))


def assert_csr_identical(X, Y):
    """Bit-exact structural and value equality of two canonical CSR matrices."""
    X = sparse.csr_matrix(X)
    Y = sparse.csr_matrix(Y)
    assert X.shape == Y.shape
    assert X.has_sorted_indices or X.sort_indices() is None
    assert np.array_equal(X.indptr.astype(np.int64), Y.indptr.astype(np.int64))
    assert np.array_equal(X.indices.astype(np.int64), Y.indices.astype(np.int64))
    assert np.array_equal(X.data, Y.data)


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ia = a.view(np.int64)
    ib = b.view(np.int64)
    return np.abs(ia - ib)
