"""GPU parity of the steps after the ARCTE path, through the
reference-shaped Python API over the C ABI:

  normalize_columns             vs embedding/common.py:49            (reference fixtures + oracle)
  chi2_contingency_matrix       vs embedding/community_weighting.py:11   BIT-EXACT
  peak_snr_weight_aggregation   vs embedding/community_weighting.py:48   BIT-EXACT
  community_weighting           vs embedding/community_weighting.py:87   <= ROW_NORM_ULP
  the experiment loop of experiments/utility.py:66-140 end to end: macro/micro-F1 EQUAL

Tolerances (helpers.LOG_ULP / ROW_NORM_ULP) exist only because of log(): numpy's SIMD log,
glibc's and CUDA's are each within 1 ulp of the true value but not of each other.
"""
import numpy as np
import pytest
import scipy.sparse as sparse

from helpers import (EPS, LOG_ULP, RHO, ROW_NORM_ULP, assert_csr_identical, load_npz_csr, load_weighting, ulp_diff,
                     weighting_chain_cases)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fixtures():
    return load_weighting("weighting600"), load_weighting("generic_weighting")


@pytest.fixture(scope="module")
def wo():
    from oracle import weighting_oracle
    weighting_oracle.lib()
    return weighting_oracle


def test_normalize_columns_vs_reference(fixtures):
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    z, zg = fixtures
    for zz, p, q in ((z, "X", "Xn_data"), (zg, "G", "Gn_data")):
        X = load_npz_csr(zz, p)
        before = X.data.copy()
        Xn = normalize_columns(X)
        assert sparse.isspmatrix_csr(Xn) and Xn.dtype == np.float64 and Xn.has_sorted_indices
        assert np.array_equal(Xn.indptr, X.indptr) and np.array_equal(Xn.indices, X.indices)
        assert ulp_diff(Xn.data, zz[q]).max() <= LOG_ULP
        assert np.array_equal(X.data, before)  # the caller's matrix is not modified


def test_normalize_columns_accepts_any_sparse_format(fixtures):
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    _, zg = fixtures
    G = load_npz_csr(zg, "G")
    for conv in (sparse.coo_matrix, sparse.csc_matrix, sparse.lil_matrix):
        M = conv(G)
        # lil/coo conversions drop nothing here except that lil removes the explicit zero
        Xn = normalize_columns(M)
        want = sparse.csr_matrix(M)
        want.sort_indices()
        assert np.array_equal(Xn.indices, want.indices)


def test_normalize_features_resident_equals_separate_call(fixtures):
    """arcte() + normalize_columns() fused on the device == the two public calls."""
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    from reveal_graph_embedding_b200.engine import get_engine
    z, _ = fixtures
    A = load_npz_csr(z, "A")
    X = arcte(A, RHO, EPS, 1)
    assert_csr_identical(X, load_npz_csr(z, "X"))
    Xn = normalize_columns(X)
    eng = get_engine(0)
    eng.set_graph(A)
    eng.extract(0, RHO, EPS)
    eng.assemble()
    eng.normalize_features()
    Xf = eng.features()
    assert_csr_identical(Xf, Xn)
    assert ulp_diff(Xf.data, z["Xn_data"]).max() <= LOG_ULP


def test_chi2_and_peak_snr_bit_exact_vs_reference(fixtures):
    from reveal_graph_embedding_b200.embedding.community_weighting import (chi2_contingency_matrix,
                                                                            peak_snr_weight_aggregation)
    for z, tag, Xtr, Xte, ytr, yte in weighting_chain_cases(*fixtures):
        cm = chi2_contingency_matrix(Xtr, ytr)
        assert cm.dtype == np.float64 and cm.shape == z[tag + "_cm"].shape
        assert np.array_equal(cm, z[tag + "_cm"])
        w = peak_snr_weight_aggregation(cm)
        assert np.array_equal(w, z[tag + "_weights"])


def test_fused_chi2_psnr_weights(fixtures):
    from reveal_graph_embedding_b200.embedding.community_weighting import _label_matrix
    from reveal_graph_embedding_b200.engine import get_engine
    for z, tag, Xtr, Xte, ytr, yte in weighting_chain_cases(*fixtures):
        w = get_engine(0).chi2_psnr_weights(Xtr, _label_matrix(ytr))
        assert np.array_equal(w, z[tag + "_weights"])


def test_community_weighting_vs_reference(fixtures):
    from reveal_graph_embedding_b200.embedding.community_weighting import (chi2_psnr_community_weighting,
                                                                            community_weighting)
    for z, tag, Xtr, Xte, ytr, yte in weighting_chain_cases(*fixtures):
        for a, b in (community_weighting(Xtr, Xte, z[tag + "_weights"]),
                     chi2_psnr_community_weighting(Xtr, Xte, ytr)):
            for got, name in ((a, "_Xtr"), (b, "_Xte")):
                want = load_npz_csr(z, tag + name)
                assert sparse.isspmatrix_csr(got) and got.shape == want.shape
                assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
                assert ulp_diff(got.data, want.data).max() <= ROW_NORM_ULP


def test_peak_snr_edge_cases():
    from reveal_graph_embedding_b200.embedding.community_weighting import peak_snr_weight_aggregation
    cm = np.array([[np.nan, 0.0, 3.0, 1.0], [2.0, 0.0, 0.0, 4.0], [0.5, 0.0, 0.0, 9.0]])
    w = peak_snr_weight_aggregation(cm)
    assert cm[0, 0] == 0.0
    noise = np.sqrt(np.mean([np.var(r) for r in cm]))
    assert np.array_equal(w, np.array([(2.0 - 0.5) / noise, 0.0, 3.0 / noise, (9.0 - 1.0) / noise]))


def test_weighting_at_scale_vs_oracle(wo):
    """ARCTE features of BA(20000, 3) (40,000 columns: the variance rows span many pairwise
    leaves and all five warp levels), 12 classes, 10 % training rows."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    from reveal_graph_embedding_b200.embedding import community_weighting as cw
    A = graphs.barabasi_albert(20000, 3, seed=2)
    X = arcte(A, RHO, EPS, 1)
    Xn = normalize_columns(X)
    On = wo.normalize_columns(X)
    assert np.array_equal(Xn.indices, On.indices) and ulp_diff(Xn.data, On.data).max() <= LOG_ULP
    rng = np.random.default_rng(3)
    n, K = A.shape[0], 12
    Y = sparse.csr_matrix((rng.random((n, K)) < 0.15).astype(np.int64))
    perm = rng.permutation(n)
    train, test = np.sort(perm[:2000]), np.sort(perm[2000:])
    Xtr, Xte, ytr = On[train, :], On[test, :], Y[train, :]   # identical inputs for both sides
    cm = cw.chi2_contingency_matrix(Xtr, ytr)
    cm_o = wo.chi2_contingency_matrix(Xtr, ytr)
    assert np.array_equal(cm, cm_o)
    w = cw.peak_snr_weight_aggregation(cm)
    w_o = wo.peak_snr_weight_aggregation(cm_o)
    assert np.array_equal(w, w_o)
    a, b = cw.community_weighting(Xtr, Xte, w)
    ao, bo = wo.community_weighting(Xtr, Xte, w_o)
    for got, want in ((a, ao), (b, bo)):
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert ulp_diff(got.data, want.data).max() <= ROW_NORM_ULP
    # the rows really are unit vectors and zero-weight columns are gone
    norms = np.sqrt(np.asarray(a.multiply(a).sum(axis=1))).ravel()
    assert np.allclose(norms[norms > 0], 1.0, rtol=0, atol=1e-12)
    assert not np.any(a.data == 0.0)


def test_downstream_f1_equals_reference(fixtures):
    """BASELINE.json config 1 in miniature (experiments/utility.py:66-140): arcte ->
    normalize_columns -> folds -> chi2/PSNR community weighting -> one-vs-rest LinearSVC ->
    macro/micro-F1, all on the GPU build except the classifier; the F1 scores must EQUAL the
    ones the unmodified reference produced on the same graph, labels and folds."""
    from sklearn import svm
    from sklearn.metrics import f1_score
    from sklearn.multiclass import OneVsRestClassifier
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    from reveal_graph_embedding_b200.embedding.community_weighting import chi2_psnr_community_weighting
    z, _ = fixtures
    A, Y = load_npz_csr(z, "A"), load_npz_csr(z, "Y")
    X = normalize_columns(arcte(A, RHO, EPS, 1))
    for k in range(2):
        train, test = z["t%d_train" % k], z["t%d_test" % k]
        X_train, X_test, y_train, y_test = X[train, :], X[test, :], Y[train, :], Y[test, :]
        X_train, X_test = chi2_psnr_community_weighting(X_train, X_test, y_train)
        model = OneVsRestClassifier(svm.LinearSVC(C=1.0, random_state=None, dual=False, fit_intercept=True))
        model.fit(X_train, y_train)
        scores = model.decision_function(X_test)
        # learning/evaluation.py:9-43: predict as many labels per node as it truly has
        true_counts = np.asarray(y_test.sum(axis=1)).ravel()
        order = np.argsort(scores, axis=1)
        pred = np.zeros(scores.shape, dtype=np.int8)
        for i, c in enumerate(true_counts):
            if c:
                pred[i, order[i, -1:-c - 1:-1]] = 1
        truth = y_test.toarray()
        assert f1_score(truth, pred, average="macro") == pytest.approx(float(z["t%d_macro_f1" % k]), abs=1e-12)
        assert f1_score(truth, pred, average="micro") == pytest.approx(float(z["t%d_micro_f1" % k]), abs=1e-12)


def test_resident_fold_equals_host_sliced_chain(fixtures, wo):
    """ResidentFeatures (row gather + chi2/PSNR + weighting on the device) == slicing on the host and
    calling the reference-shaped functions, for the fixture folds and for a larger matrix."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.embedding.common import normalize_columns
    from reveal_graph_embedding_b200.embedding.community_weighting import (ResidentFeatures,
                                                                            chi2_psnr_community_weighting)
    from reveal_graph_embedding_b200.engine import get_engine
    z, _ = fixtures
    X = load_npz_csr(z, "X")
    Xn = sparse.csr_matrix((z["Xn_data"], X.indices, X.indptr), shape=X.shape)
    Y = load_npz_csr(z, "Y")
    res = ResidentFeatures(Xn)
    for k in range(2):
        tr, te = z["t%d_train" % k], z["t%d_test" % k]
        a, b = res.chi2_psnr_community_weighting(tr, te, Y[tr, :])
        for got, name in ((a, "_Xtr"), (b, "_Xte")):
            want = load_npz_csr(z, "t%d" % k + name)
            assert got.shape == want.shape
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert ulp_diff(got.data, want.data).max() <= ROW_NORM_ULP
    # larger, unsorted fold indices, features adopted straight from the device
    A = graphs.barabasi_albert(8000, 3, seed=4)
    Xh = normalize_columns(arcte(A, RHO, EPS, 1))
    eng = get_engine(0)
    eng.set_graph(A)
    eng.extract(0, RHO, EPS)
    eng.assemble()
    eng.normalize_features()
    res = ResidentFeatures(None)
    rng = np.random.default_rng(8)
    n, K = A.shape[0], 9
    Yb = sparse.csr_matrix((rng.random((n, K)) < 0.2).astype(np.int64))
    perm = rng.permutation(n)
    tr, te = perm[:700], perm[700:5000]                      # not sorted, not covering every row
    a, b = res.chi2_psnr_community_weighting(tr, te, Yb[tr, :])
    a0, b0 = chi2_psnr_community_weighting(Xh[tr, :], Xh[te, :], Yb[tr, :])
    assert_csr_identical(a, a0)
    assert_csr_identical(b, b0)
    # empty test block
    a, b = res.chi2_psnr_community_weighting(tr, np.zeros(0, dtype=np.int64), Yb[tr, :])
    assert_csr_identical(a, a0) and b.shape == (0, 2 * n) and b.nnz == 0
