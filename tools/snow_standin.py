"""BASELINE.json config 1 without the missing edge file: a 533,874-node stand-in for SNOW2014Graph whose
community structure is planted from the dataset's REAL label matrix (tests/golden/snow_labels.npz, copied
from the reference's snow2014graph/user_label_matrix.tsv: 10,992 labelled nodes, 90 labels), run through
ARCTE and through the reference's experiment chain (experiments/utility.py:66-140, experiments/demo.py:67-68:
training sets of 1..10 % of the labelled nodes, 10 trials each).

THE SUBSTITUTION: the reference's men_ret_graph.tsv is absent from its tree (.MISSING_LARGE_BLOBS:1); the graph
here is synthetic (seed 2014): every label is a community of its labelled nodes plus a share of the unlabelled
ones, edges are drawn inside communities and, with a heavy-tailed node propensity, across them.  Only the
labels are real.  F1 numbers are therefore about this stand-in, not about the dataset.

    python tools/snow_standin.py gpu  [--ref-sample K]   # on the GPU box: extract, time, hash, compare
    python tools/snow_standin.py f1   [--trials T]       # CPU: the experiment chain on the same features

`gpu`  extracts the features with the library, prints their content hash and timings, checks the WHOLE matrix
       against the oracle port (all host threads) and K sampled columns against the UNMODIFIED Python reference
       (baseline/_ref: arcte_worker, arcte.py:279).
`f1`   computes the same features with the oracle port (same hash, printed), then normalize_columns ->
       generate_folds (the reference's, learning/holdout.py) -> chi2 / peak-SNR community weighting ->
       OneVsRest LinearSVC(C=1, dual=False) -> the reference's evaluation measures, per training percentage.
       With --use-gpu the features and the weighting chain come from the library instead (GPU box).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import scipy.sparse as sparse  # noqa: E402

RHO, EPS = 0.1, 1e-5


def load_labels():
    z = np.load(os.path.join(ROOT, "tests", "golden", "snow_labels.npz"))
    n, k = int(z["n_rows"]), int(z["n_cols"])
    Y = sparse.coo_matrix((np.ones(z["rows"].size, dtype=np.int64), (z["rows"].astype(np.int64), z["cols"].astype(np.int64))),
                          shape=(n, k)).tocsr()
    Y.sum_duplicates()
    Y.data[:] = 1
    return Y


def standin_graph(Y, seed=2014, intra_per_member=1.5, noise_edges=300000):
    """Undirected, unweighted.  Community c = labelled nodes of label c + unlabelled nodes drawn with a heavy-tailed
    propensity (so that a few accounts are in many conversations, like a mention/retweet graph)."""
    rng = np.random.default_rng(seed)
    n, k = Y.shape
    labelled = np.unique(Y.nonzero()[0])
    unl = np.setdiff1d(np.arange(n), labelled)
    prop = np.minimum(rng.pareto(1.8, size=n) + 1.0, 300.0)   # node propensity, heavy tail (capped)
    p_unl = prop[unl] / prop[unl].sum()
    Yc = Y.tocsc()
    rows, cols = [], []
    for c in range(k):
        core = Yc.indices[Yc.indptr[c]:Yc.indptr[c + 1]]
        extra = rng.choice(unl, size=min(unl.size, 40 * core.size), replace=False, p=p_unl)
        members = np.concatenate([core, extra])
        w = np.concatenate([np.full(core.size, 8.0), prop[extra]])
        w /= w.sum()
        m = int(intra_per_member * members.size)
        rows.append(rng.choice(members, size=m, p=w))
        cols.append(rng.choice(members, size=m, p=w))
    p_all = prop / prop.sum()
    rows.append(rng.choice(n, size=noise_edges, p=p_all))
    cols.append(rng.integers(0, n, size=noise_edges))
    r, c = np.concatenate(rows), np.concatenate(cols)
    keep = r != c
    r, c = r[keep], c[keep]
    A = sparse.coo_matrix((np.ones(2 * r.size), (np.concatenate([r, c]), np.concatenate([c, r]))), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.data[:] = 1.0
    A.sort_indices()
    return A


def describe(A, Y):
    deg = np.diff(A.indptr)
    return {"nodes": int(A.shape[0]), "nnz": int(A.nnz), "isolated": int((deg == 0).sum()), "max_degree": int(deg.max()),
            "labelled_nodes": int(np.unique(Y.nonzero()[0]).size), "labels": int(Y.shape[1]), "label_entries": int(Y.nnz),
            "substitution": "synthetic graph planted from the REAL SNOW2014Graph label matrix; the dataset's edge file is "
                            "absent from the reference tree"}


def port_features(A, threads):
    from oracle import arcte_oracle as O
    O.build()
    return O.arcte(A, RHO, EPS, threads)


def reference_imports():
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(d, "reveal_graph_embedding")):
            sys.path.insert(0, d)
            return d
    raise SystemExit("the unmodified reference is needed (baseline/_ref or /root/reference)")


def run_gpu(args):
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.engine import csr_hash, get_engine
    Y = load_labels()
    A = standin_graph(Y)
    out = {"config": describe(A, Y)}
    t = time.perf_counter(); X = arcte(A, RHO, EPS, 1); out["arcte_first_call_s"] = time.perf_counter() - t
    t = time.perf_counter(); X = arcte(A, RHO, EPS, 1); out["arcte_s"] = time.perf_counter() - t
    st = get_engine(0).stats()
    out.update({"seeds": st["n_seeds_total"], "features_nnz": int(X.nnz), "features_hash": "%016x" % csr_hash(X),
                "stage_ms": {k: st[k] for k in ("ms_transition", "ms_seeds", "ms_push", "ms_assemble")},
                "seeds_per_s": st["n_seeds_total"] / out["arcte_s"], "engine": st["engine"]})
    threads = os.cpu_count() or 1
    t = time.perf_counter(); Xp = port_features(A, threads); out["port_s"] = time.perf_counter() - t
    out["port_threads"] = threads
    out["identical_to_port"] = bool(np.array_equal(X.indptr, Xp.indptr) and np.array_equal(X.indices, Xp.indices)
                                    and np.array_equal(X.data, Xp.data))
    # K columns against the UNMODIFIED Python reference
    ref_dir = reference_imports()
    from reveal_graph_embedding.embedding.arcte.arcte import arcte_worker
    from oracle import arcte_oracle as O
    g = O.Graph(A)
    seeds = g.seeds()
    idx = np.unique(np.linspace(0, seeds.size - 1, args.ref_sample).astype(np.int64))
    sample = seeds[idx]
    t = time.perf_counter()
    local = arcte_worker(sample, g.indices.astype(np.int64), g.indptr, g.w, g.d_out, g.d_in, RHO, EPS)
    out["python_reference"] = {"seeds": int(sample.size), "seconds_one_process": time.perf_counter() - t, "from": ref_dir}
    n = A.shape[0]
    got = X[:, n:].tocsc()
    want = sparse.csc_matrix(local)
    same = all(np.array_equal(np.sort(got.indices[got.indptr[s]:got.indptr[s + 1]]),
                              np.sort(want.indices[want.indptr[s]:want.indptr[s + 1]])) for s in sample)
    out["python_reference"]["columns_identical"] = bool(same)
    print(json.dumps(out))


def run_f1(args):
    ref_dir = reference_imports()
    from reveal_graph_embedding.learning import evaluation
    from reveal_graph_embedding.learning.holdout import generate_folds
    from sklearn import svm
    from sklearn.multiclass import OneVsRestClassifier
    Y = load_labels()
    A = standin_graph(Y)
    out = {"config": describe(A, Y), "reference_functions_from": ref_dir}
    if args.use_gpu:
        from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
        from reveal_graph_embedding_b200.embedding.common import normalize_columns
        from reveal_graph_embedding_b200.embedding.community_weighting import chi2_psnr_community_weighting
        t = time.perf_counter(); X = arcte(A, RHO, EPS, 1); out["features_s"] = time.perf_counter() - t
    else:
        from oracle import weighting_oracle as WO
        t = time.perf_counter(); X = port_features(A, os.cpu_count() or 1); out["features_s"] = time.perf_counter() - t
        normalize_columns = WO.normalize_columns

        def chi2_psnr_community_weighting(X_train, X_test, y_train):   # utility.py:101-104
            w = WO.peak_snr_weight_aggregation(WO.chi2_contingency_matrix(X_train, y_train))
            return WO.community_weighting(X_train, X_test, w)
    from reveal_graph_embedding_b200.engine import csr_hash
    out["features_nnz"], out["features_hash"] = int(X.nnz), "%016x" % csr_hash(X)
    X = normalize_columns(X)                                        # utility.py:66
    labelled = np.unique(Y.nonzero()[0]).astype(np.int64)           # utility.py:40: folds over labelled nodes only
    table = []
    for pct in range(1, 11):                                        # demo.py:67: 1..10 %
        folds = generate_folds(Y, labelled, Y.shape[1], float(pct), args.trials)
        macro, micro = [], []
        for trial in range(args.trials):
            train, test = next(folds)
            X_train, X_test, y_train, y_test = X[train, :], X[test, :], Y[train, :], Y[test, :]
            X_train, X_test = chi2_psnr_community_weighting(X_train, X_test, y_train)    # utility.py:101-104
            model = OneVsRestClassifier(svm.LinearSVC(C=1.0, random_state=None, dual=False, fit_intercept=True),
                                        n_jobs=args.jobs)                                # utility.py:116-120
            model.fit(X_train, y_train)
            y_pred = evaluation.form_node_label_prediction_matrix(model.decision_function(X_test), y_test)
            m = evaluation.calculate_measures(y_pred, y_test)
            macro.append(float(m[4])), micro.append(float(m[5]))
        table.append({"train_percent": pct, "macro_f1": float(np.mean(macro)), "macro_f1_std": float(np.std(macro)),
                      "micro_f1": float(np.mean(micro)), "micro_f1_std": float(np.std(micro)), "trials": args.trials})
        print("train %2d %%: macro-F1 %.4f +- %.4f, micro-F1 %.4f +- %.4f" % (pct, table[-1]["macro_f1"], table[-1]["macro_f1_std"],
                                                                              table[-1]["micro_f1"], table[-1]["micro_f1_std"]),
              file=sys.stderr, flush=True)
    out["f1"] = table
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["gpu", "f1"])
    ap.add_argument("--ref-sample", type=int, default=300)
    ap.add_argument("--trials", type=int, default=10)
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--use-gpu", action="store_true")
    a = ap.parse_args()
    (run_gpu if a.mode == "gpu" else run_f1)(a)
