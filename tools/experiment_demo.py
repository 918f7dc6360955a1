"""The reference's experiment (experiments/demo.py -> experiments/utility.py:24-155) end to end on
this package: ARCTE features -> normalize_columns -> folds -> chi2/peak-SNR community weighting ->
one-vs-rest LinearSVC -> macro/micro-F1, with the wall time of every stage.  The graph is a
planted-partition stand-in with noisy multi-labels (the reference's datasets are not
redistributable and SNOW2014Graph's edge file is absent from its tree).

    python tools/experiment_demo.py [n_nodes] [n_groups] [train percent] [trials]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sparse
from sklearn import svm
from sklearn.metrics import f1_score
from sklearn.multiclass import OneVsRestClassifier

from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
from reveal_graph_embedding_b200.embedding.common import normalize_columns
from reveal_graph_embedding_b200.embedding.community_weighting import chi2_psnr_community_weighting

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 20
percent = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
trials = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(2014)
lab = rng.integers(0, groups, size=n)
# sparse planted partition: ~12 intra-group and ~4 inter-group neighbours per node
m_in, m_out = 6 * n, 2 * n
order = np.argsort(lab, kind="stable")
starts = np.searchsorted(lab[order], np.arange(groups + 1))
u = rng.integers(0, n, size=m_in)
gu = lab[u]
v = order[starts[gu] + (rng.random(m_in) * (starts[gu + 1] - starts[gu])).astype(np.int64)]
u2, v2 = rng.integers(0, n, size=m_out), rng.integers(0, n, size=m_out)
r, c = np.concatenate([u, u2]), np.concatenate([v, v2])
keep = r != c
A = sparse.coo_matrix((np.ones(2 * keep.sum()), (np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]]))),
                      shape=(n, n)).tocsr()
A.sum_duplicates(); A.data[:] = 1.0
rows = np.arange(n)
noisy = np.where(rng.random(n) < 0.1, rng.integers(0, groups, size=n), lab)
extra = rng.random(n) < 0.3
Y = sparse.coo_matrix((np.ones(n + extra.sum(), dtype=np.int64),
                       (np.concatenate([rows, rows[extra]]), np.concatenate([noisy, rng.integers(0, groups, size=extra.sum())]))),
                      shape=(n, groups)).tocsr()
Y.sum_duplicates(); Y.data[:] = 1
print("graph: n=%d nnz=%d, %d labels, %.2f labels/node" % (n, A.nnz, groups, Y.nnz / n))
t = time.perf_counter(); X = arcte(A, 0.1, 1e-5); t_first = time.perf_counter() - t   # CUDA context + walk-state pool
t = time.perf_counter(); X = arcte(A, 0.1, 1e-5); t_arcte = time.perf_counter() - t
t = time.perf_counter(); X = normalize_columns(X); t_norm = time.perf_counter() - t
print("arcte %.3f s (first call in the process %.3f s; features %d x %d, nnz %d), normalize_columns %.3f s"
      % (t_arcte, t_first, X.shape[0], X.shape[1], X.nnz, t_norm))
macro, micro = [], []
for trial in range(trials):
    perm = np.random.default_rng(trial).permutation(n)
    k = int(np.ceil(percent * n / 100))
    train, test = np.sort(perm[:k]), np.sort(perm[k:])
    t = time.perf_counter()
    X_train, X_test, y_train, y_test = X[train, :], X[test, :], Y[train, :], Y[test, :]
    t_slice = time.perf_counter() - t
    t = time.perf_counter(); X_train, X_test = chi2_psnr_community_weighting(X_train, X_test, y_train); t_w = time.perf_counter() - t
    t = time.perf_counter()
    model = OneVsRestClassifier(svm.LinearSVC(C=1.0, random_state=None, dual=False, fit_intercept=True))
    model.fit(X_train, y_train); t_fit = time.perf_counter() - t
    t = time.perf_counter(); scores = model.decision_function(X_test); t_pred = time.perf_counter() - t
    counts = np.asarray(y_test.sum(axis=1)).ravel()           # learning/evaluation.py:9-43
    idx = np.argsort(scores, axis=1)
    pred = np.zeros(scores.shape, dtype=np.int8)
    for i, cnt in enumerate(counts):
        if cnt:
            pred[i, idx[i, -1:-cnt - 1:-1]] = 1
    truth = y_test.toarray()
    macro.append(f1_score(truth, pred, average="macro")); micro.append(f1_score(truth, pred, average="micro"))
    print("trial %d: slice %.3f s, community weighting %.3f s, LinearSVC fit %.3f s, predict %.3f s, macro-F1 %.4f micro-F1 %.4f"
          % (trial, t_slice, t_w, t_fit, t_pred, macro[-1], micro[-1]))
print("%.0f %% training nodes: macro-F1 %.4f +- %.4f, micro-F1 %.4f +- %.4f" % (percent, np.mean(macro), np.std(macro), np.mean(micro), np.std(micro)))
