"""Pins oracle/weighting_oracle.c against fixtures produced by the unmodified Python
reference (tests/golden/make_golden_weighting.py): normalize_columns
(embedding/common.py:49), chi2_contingency_matrix / peak_snr_weight_aggregation /
community_weighting (embedding/community_weighting.py:11-125).  Bit-exact except where a
logarithm is involved (helpers.LOG_ULP / ROW_NORM_ULP)."""
import numpy as np
import pytest
import scipy.sparse as sparse

from helpers import (LOG_ULP, ROW_NORM_ULP, load_npz_csr, load_weighting, ulp_diff, weighting_chain_cases)


@pytest.fixture(scope="module")
def wo():
    from oracle import weighting_oracle
    weighting_oracle.lib()
    return weighting_oracle


@pytest.fixture(scope="module")
def fixtures():
    return load_weighting("weighting600"), load_weighting("generic_weighting")


def test_normalize_columns_arcte_features(wo, fixtures):
    z, _ = fixtures
    X = load_npz_csr(z, "X")
    Xn = wo.normalize_columns(X)
    assert np.array_equal(Xn.indptr, X.indptr) and np.array_equal(Xn.indices, X.indices)
    assert ulp_diff(Xn.data, z["Xn_data"]).max() <= LOG_ULP
    assert np.array_equal(X.data, z["X_data"])  # input untouched


def test_normalize_columns_generic(wo, fixtures):
    _, z = fixtures
    G = load_npz_csr(z, "G")
    Gn = wo.normalize_columns(G)
    assert ulp_diff(Gn.data, z["Gn_data"]).max() <= LOG_ULP
    # singleton and empty columns are left alone (document frequency <= 1, common.py:61)
    k = np.where(G.indices == 6)[0]
    assert k.size == 1 and Gn.data[k[0]] == 2.5


def test_var_matches_numpy(wo, fixtures):
    _, z = fixtures
    off = 0
    for L, v in zip(z["var_lens"], z["var_out"]):
        row = z["var_rows"][off:off + L]
        off += L
        assert wo.var(row) == v
        assert wo.var(row) == np.var(row)  # and the numpy installed here


def test_chi2_contingency_bit_exact(wo, fixtures):
    for z, tag, Xtr, Xte, ytr, yte in weighting_chain_cases(*fixtures):
        cm = wo.chi2_contingency_matrix(Xtr, ytr)
        assert cm.shape == z[tag + "_cm"].shape
        assert np.array_equal(cm, z[tag + "_cm"])


def test_peak_snr_bit_exact(wo, fixtures):
    for z, tag, *_ in weighting_chain_cases(*fixtures):
        cm = z[tag + "_cm"].copy()
        w = wo.peak_snr_weight_aggregation(cm)
        assert np.array_equal(w, z[tag + "_weights"])


def test_peak_snr_nan_and_empty_columns(wo):
    cm = np.array([[np.nan, 0.0, 3.0, 1.0], [2.0, 0.0, 0.0, 4.0], [0.5, 0.0, 0.0, 9.0]])
    w = wo.peak_snr_weight_aggregation(cm)
    assert cm[0, 0] == 0.0                      # nan -> 0 in place (community_weighting.py:49)
    noise = np.sqrt(np.mean([np.var(r) for r in cm]))
    assert np.array_equal(w, np.array([(2.0 - 0.5) / noise, 0.0, 3.0 / noise, (9.0 - 1.0) / noise]))


def test_community_weighting(wo, fixtures):
    for z, tag, Xtr, Xte, ytr, yte in weighting_chain_cases(*fixtures):
        a, b = wo.community_weighting(Xtr, Xte, z[tag + "_weights"])
        for got, name in ((a, "_Xtr"), (b, "_Xte")):
            want = load_npz_csr(z, tag + name)
            assert got.shape == want.shape
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert ulp_diff(got.data, want.data).max() <= ROW_NORM_ULP


def test_single_label_column_is_expanded(wo):
    rng = np.random.default_rng(1)
    X = sparse.random(40, 30, density=0.2, random_state=rng, format="csr")
    y = sparse.csr_matrix((rng.random((40, 1)) < 0.4).astype(np.int64))
    cm = wo.chi2_contingency_matrix(X, y)
    assert cm.shape == (2, 30)                  # community_weighting.py:20-21
