"""Shared helpers for the parity tests."""
import os

import numpy as np
import scipy.sparse as sparse

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["ba300", "weighted200", "planted419", "edgecases160", "ba2000", "ba5000"]
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


def load_npz_csr(z, prefix):
    shape = tuple(int(x) for x in z[prefix + "_shape"])
    return sparse.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=shape)


def load_weighting(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


# The reference's np.log goes through numpy's SIMD log on the generating host; glibc's and
# CUDA's log each stay within 1 ulp of the true value but not of each other.  A column scale
# 1/sqrt(log(df)) or log(1+w) therefore moves by at most 1-2 ulp, and an l2-normalised row
# (hundreds of such terms) by a few more.
LOG_ULP = 2
ROW_NORM_ULP = 8


def weighting_chain_cases(zexp, zgen):
    """(tag, X_train, X_test, y_train) of every fold stored in the two weighting fixtures."""
    n = int(zexp["X_shape"][0])
    X = load_npz_csr(zexp, "X")
    Xn = sparse.csr_matrix((zexp["Xn_data"], X.indices, X.indptr), shape=X.shape)
    Y = load_npz_csr(zexp, "Y")
    cases = []
    for k in range(2):
        tr, te = zexp["t%d_train" % k], zexp["t%d_test" % k]
        cases.append((zexp, "t%d" % k, Xn[tr, :], Xn[te, :], Y[tr, :], Y[te, :]))
    G = load_npz_csr(zgen, "G")
    Gn = sparse.csr_matrix((zgen["Gn_data"], G.indices, G.indptr), shape=G.shape)
    Gy = load_npz_csr(zgen, "Gy")
    tr, te = zgen["g_train"], zgen["g_test"]
    cases.append((zgen, "g", Gn[tr, :], Gn[te, :], Gy[tr, :], Gy[te, :]))
    assert n == 600
    return cases
