"""Generate the golden fixtures of the steps that FOLLOW the ARCTE hot path in every
experiment of the reference: column normalisation, the chi2 /
peak-SNR community weighting, and the macro/micro-F1 the weighted features give.

Run (only in the build container; the GPU box has no /root/reference):

    python tests/golden/make_golden_weighting.py

Everything below is computed by importing the UNMODIFIED Python reference:

  normalize_columns             reveal_graph_embedding/embedding/common.py:49-67
  chi2_contingency_matrix       reveal_graph_embedding/embedding/community_weighting.py:11-45
  peak_snr_weight_aggregation   reveal_graph_embedding/embedding/community_weighting.py:48-84
  community_weighting           reveal_graph_embedding/embedding/community_weighting.py:87-125
  generate_folds                reveal_graph_embedding/learning/holdout.py:80-111
  form_node_label_prediction_matrix / calculate_measures   learning/evaluation.py:9-74
  experiment loop               reveal_graph_embedding/experiments/utility.py:66-140

weighting600.npz
  A_*            planted-partition graph, 600 nodes, 6 groups (canonical CSR)
  Y_*            multi-label node-label matrix (int64 CSR), every node labelled
  X_*            arcte(A, 0.1, 1e-5, 1)                       (the hot path's output)
  Xn_data        normalize_columns(X).data  (same structure as X)
  t{k}_train/test                 fold k of generate_folds(Y, all nodes, 6, 10 %, 2 folds)
  t{k}_cm, t{k}_weights           chi2_contingency_matrix / peak_snr_weight_aggregation on the fold
  t{k}_Xtr_*, t{k}_Xte_*          community_weighting(X_train, X_test, weights)
  t{k}_macro_f1, t{k}_micro_f1    LinearSVC(C=1, dual=False) one-vs-rest, evaluation.py measures
generic_weighting.npz
  G_*            a generic float CSR (values != 1, empty / singleton columns, explicit zeros)
  Gn_data        normalize_columns(G).data
  Gy, G_cm, G_weights, Gw_*       the same chain on G with a dense-ish label matrix
  var_rows, var_out               np.var on long rows (pins numpy's pairwise order at F > 128)
"""
import os
import sys

import numpy as np
import scipy.sparse as sparse

sys.path.insert(0, "/root/reference")

from reveal_graph_embedding.embedding.arcte import arcte as ref_arcte  # noqa: E402
from reveal_graph_embedding.embedding.common import normalize_columns  # noqa: E402
from reveal_graph_embedding.embedding.community_weighting import (  # noqa: E402
    chi2_contingency_matrix, community_weighting, peak_snr_weight_aggregation)
from reveal_graph_embedding.learning import evaluation  # noqa: E402
from reveal_graph_embedding.learning.holdout import generate_folds  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RHO, EPS = 0.1, 1e-5


def put_csr(out, prefix, X, index_dtype=np.int64):
    X = sparse.csr_matrix(X)
    X.sort_indices()
    out[prefix + "_indptr"] = X.indptr.astype(np.int64)
    out[prefix + "_indices"] = X.indices.astype(index_dtype)
    out[prefix + "_data"] = X.data.copy()
    out[prefix + "_shape"] = np.array(X.shape, dtype=np.int64)


def planted_with_labels(n, groups, p_in, p_out, seed):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, groups, size=n)
    P = np.where(lab[:, None] == lab[None, :], p_in, p_out)
    M = np.triu(rng.random((n, n)) < P, k=1)
    r, c = np.nonzero(M)
    A = sparse.coo_matrix((np.ones(2 * r.size), (np.concatenate([r, c]), np.concatenate([c, r]))),
                          shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    # multi-label: the planted group, plus a second random label for 30 % of the nodes,
    # and 10 % of the primary labels replaced by noise
    rows, cols = [], []
    for v in range(n):
        g = int(lab[v]) if rng.random() > 0.10 else int(rng.integers(0, groups))
        rows.append(v); cols.append(g)
        if rng.random() < 0.30:
            h = int(rng.integers(0, groups))
            if h != g:
                rows.append(v); cols.append(h)
    Y = sparse.coo_matrix((np.ones(len(rows), dtype=np.int64), (rows, cols)), shape=(n, groups)).tocsr()
    return A, Y


def run_chain(out, tag, X_train, X_test, y_train, y_test=None):
    cm = chi2_contingency_matrix(X_train, y_train)
    out[tag + "_cm"] = cm.copy()
    w = peak_snr_weight_aggregation(cm)          # mutates cm (nan -> 0) like the reference
    out[tag + "_weights"] = w.copy()
    Xtr, Xte = community_weighting(X_train, X_test, w)
    put_csr(out, tag + "_Xtr", Xtr)
    put_csr(out, tag + "_Xte", Xte)
    return Xtr, Xte


def fixture_experiment():
    from sklearn import svm
    from sklearn.multiclass import OneVsRestClassifier
    out = {}
    n, groups = 600, 6
    A, Y = planted_with_labels(n, groups, 0.05, 0.004, 600)
    put_csr(out, "A", A, np.int32)
    put_csr(out, "Y", Y, np.int32)
    X = sparse.csr_matrix(ref_arcte.arcte(A.copy(), RHO, EPS, 1))
    X.sort_indices()
    put_csr(out, "X", X)
    Xn = normalize_columns(X)                     # experiments/utility.py:66
    assert np.array_equal(Xn.indices, X.indices) and np.array_equal(Xn.indptr, X.indptr)
    out["Xn_data"] = Xn.data.copy()
    folds = generate_folds(Y, np.arange(n), groups, 10, 2)   # utility.py:83
    for k in range(2):
        train, test = next(folds)
        out["t%d_train" % k], out["t%d_test" % k] = np.asarray(train, np.int64), np.asarray(test, np.int64)
        X_train, X_test, y_train, y_test = Xn[train, :], Xn[test, :], Y[train, :], Y[test, :]
        Xtr, Xte = run_chain(out, "t%d" % k, X_train, X_test, y_train)
        model = OneVsRestClassifier(svm.LinearSVC(C=1.0, random_state=None, dual=False, fit_intercept=True))
        model.fit(Xtr, y_train)
        y_pred = model.decision_function(Xte)
        y_pred = evaluation.form_node_label_prediction_matrix(y_pred, y_test)
        m = evaluation.calculate_measures(y_pred, y_test)
        out["t%d_macro_f1" % k], out["t%d_micro_f1" % k] = np.float64(m[4]), np.float64(m[5])
        print("fold %d: train %d test %d  macro-F1 %.6f micro-F1 %.6f" % (k, train.size, test.size, m[4], m[5]))
    path = os.path.join(HERE, "weighting600.npz")
    np.savez_compressed(path, **out)
    print("weighting600: X nnz=%d local nnz=%d  %.1f KB" % (X.nnz, X.nnz - A.nnz - n, os.path.getsize(path) / 1024))


def fixture_generic():
    out = {}
    rng = np.random.default_rng(2024)
    n_rows, n_cols, K = 300, 700, 7
    G = sparse.random(n_rows, n_cols, density=0.03, random_state=rng, format="csr",
                      data_rvs=lambda k: rng.uniform(0.2, 4.0, size=k))
    G = sparse.lil_matrix(G)
    G[:, 5] = 0.0            # empty column
    G[:, 6] = 0.0
    G[17, 6] = 2.5           # singleton column (document frequency 1: left untouched)
    G[:, 7] = 1.0            # full column
    G = sparse.csr_matrix(G)
    G.sort_indices()
    G.data[3] = 0.0          # an explicit stored zero (counts toward the document frequency)
    put_csr(out, "G", G, np.int32)
    Gn = normalize_columns(G.copy())
    Gn.sort_indices()
    assert np.array_equal(Gn.indices, G.indices)
    out["Gn_data"] = Gn.data.copy()
    Yd = (rng.random((n_rows, K)) < 0.2).astype(np.int64)
    Yd[:, 3] = 0             # a class nobody has
    Y = sparse.csr_matrix(Yd)
    put_csr(out, "Gy", Y, np.int32)
    train = np.sort(rng.choice(n_rows, size=120, replace=False))
    test = np.setdiff1d(np.arange(n_rows), train)
    out["g_train"], out["g_test"] = train, test
    run_chain(out, "g", Gn[train, :], Gn[test, :], Y[train, :])
    # np.var on long contiguous rows: pins the pairwise tree for F > 128
    lens = [1, 5, 8, 9, 127, 128, 129, 1000, 4097, 33333]
    rows, var_out = [], []
    for L in lens:
        v = rng.uniform(0.0, 50.0, size=L) * rng.choice([1.0, 1e-4, 1e2], size=L)
        rows.append(v)
        var_out.append(np.var(v))
    out["var_lens"] = np.array(lens, dtype=np.int64)
    out["var_rows"] = np.concatenate(rows)
    out["var_out"] = np.array(var_out)
    path = os.path.join(HERE, "generic_weighting.npz")
    np.savez_compressed(path, **out)
    print("generic_weighting: %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    fixture_experiment()
    fixture_generic()
