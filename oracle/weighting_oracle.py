"""ctypes front-end of oracle/weighting_oracle.c (column normalisation and the chi2 /
peak-SNR community weighting that follow the ARCTE path in the reference's experiments).

TEST INFRASTRUCTURE ONLY -- see the header of weighting_oracle.c.  The function names and
signatures are the reference's (embedding/common.py:49, embedding/community_weighting.py).

Parity status: PINNED against the unmodified Python reference through
tests/golden/weighting600.npz and generic_weighting.npz (tests/test_oracle_weighting.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libweighting_oracle.so")
        src = os.path.join(_HERE, "weighting_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", "libweighting_oracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.oracle_normalize_columns.argtypes = [C.c_int64, C.c_int64, _i64p, _i32p, _f64p]
        L.oracle_var.argtypes = [_f64p, C.c_int64, _f64p]
        L.oracle_var.restype = C.c_double
        L.oracle_chi2_contingency.argtypes = [C.c_int64, C.c_int64, _i64p, _i32p, C.c_int64, _i64p, _i32p, _f64p,
                                              _f64p]
        L.oracle_peak_snr.argtypes = [C.c_int64, C.c_int64, _f64p, _f64p]
        L.oracle_community_weighting.argtypes = [C.c_int64, C.c_int64, _i64p, _i32p, _f64p, _f64p, _i64p, _i32p,
                                                 _f64p]
        L.oracle_community_weighting.restype = C.c_int64
        _LIB = L
    return _LIB


def _csr(X):
    X = sparse.csr_matrix(X, dtype=np.float64)
    if not X.has_sorted_indices:
        X = X.copy()
        X.sort_indices()
    return (X.shape, np.ascontiguousarray(X.indptr, dtype=np.int64), np.ascontiguousarray(X.indices, dtype=np.int32),
            np.ascontiguousarray(X.data, dtype=np.float64))


def label_matrix(y_train):
    """LabelBinarizer().fit_transform(y) as chi2_contingency_matrix uses it
    (community_weighting.py:19-21): an indicator matrix passes through, a single column
    becomes [1 - Y, Y]."""
    if sparse.issparse(y_train) or (isinstance(y_train, np.ndarray) and y_train.ndim == 2):
        Y = sparse.csr_matrix(y_train)
    else:
        from sklearn.preprocessing import LabelBinarizer
        Y = sparse.csr_matrix(LabelBinarizer().fit_transform(y_train))
    if Y.shape[1] == 1:
        d = np.asarray(Y.todense())
        Y = sparse.csr_matrix(np.append(1 - d, d, axis=1))
    return Y


def normalize_columns(features):
    """embedding/common.py:49-67."""
    shape, indptr, indices, data = _csr(features)
    data = data.copy()
    lib().oracle_normalize_columns(shape[0], shape[1], indptr, indices, data)
    return sparse.csr_matrix((data, indices, indptr.astype(np.int32)), shape=shape)


def var(row):
    row = np.ascontiguousarray(row, dtype=np.float64)
    return float(lib().oracle_var(row, row.size, np.empty_like(row)))


def chi2_contingency_matrix(X_train, y_train):
    """embedding/community_weighting.py:11-45."""
    shape, indptr, indices, _ = _csr(X_train)
    _, y_indptr, y_indices, y_data = _csr(label_matrix(y_train))
    K = label_matrix(y_train).shape[1]
    out = np.empty((K, shape[1]), dtype=np.float64)
    lib().oracle_chi2_contingency(shape[0], shape[1], indptr, indices, K, y_indptr, y_indices, y_data, out)
    return out


def peak_snr_weight_aggregation(contingency_matrix):
    """embedding/community_weighting.py:48-84 (mutates its argument like the reference)."""
    assert contingency_matrix.dtype == np.float64 and contingency_matrix.flags.c_contiguous
    K, F = contingency_matrix.shape
    w = np.empty(F, dtype=np.float64)
    lib().oracle_peak_snr(K, F, contingency_matrix, w)
    return w


def _weight_one(X, community_weights):
    shape, indptr, indices, data = _csr(X)
    out_indptr = np.empty(shape[0] + 1, dtype=np.int64)
    out_indices = np.empty(max(data.size, 1), dtype=np.int32)
    out_data = np.empty(max(data.size, 1), dtype=np.float64)
    w = np.ascontiguousarray(community_weights, dtype=np.float64)
    nnz = lib().oracle_community_weighting(shape[0], shape[1], indptr, indices, data, w, out_indptr, out_indices,
                                           out_data)
    return sparse.csr_matrix((out_data[:nnz], out_indices[:nnz], out_indptr.astype(np.int32)), shape=shape)


def community_weighting(X_train, X_test, community_weights):
    """embedding/community_weighting.py:87-125 (sparse branch)."""
    return _weight_one(X_train, community_weights), _weight_one(X_test, community_weights)
