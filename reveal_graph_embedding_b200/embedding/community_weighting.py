"""Supervised community weighting on the GPU -- drop-in for
reveal_graph_embedding/embedding/community_weighting.py of the reference (same names,
arguments and return values):

    chi2_contingency_matrix(X_train, y_train)                       community_weighting.py:11
    peak_snr_weight_aggregation(contingency_matrix)                 community_weighting.py:48
    community_weighting(X_train, X_test, community_weights)         community_weighting.py:87
    chi2_psnr_community_weighting(X_train, X_test, y_train)         community_weighting.py:128

The label matrix is binarised on the host exactly like the reference does
(sklearn LabelBinarizer, community_weighting.py:19-21); everything numeric runs on the
device through the C ABI (include/arcte_cuda.h).  There is no CPU path.
"""
import numpy as np
import scipy.sparse as sparse
from scipy.sparse import issparse

from ..engine import get_engine


def _label_matrix(y_train):
    """LabelBinarizer().fit_transform(y_train) as a sparse matrix; a single column becomes
    [1 - Y, Y] (community_weighting.py:19-21)."""
    from sklearn.preprocessing import LabelBinarizer
    Y = LabelBinarizer(sparse_output=True).fit_transform(y_train)
    Y = sparse.csr_matrix(Y, dtype=np.float64)
    if Y.shape[1] == 1:
        d = np.asarray(Y.todense())
        Y = sparse.csr_matrix(np.append(1 - d, d, axis=1))
    return Y


def chi2_contingency_matrix(X_train, y_train):
    if np.any((X_train.data if issparse(X_train) else X_train) < 0):
        raise ValueError("Input X must be non-negative.")  # community_weighting.py:16-17
    return get_engine(0).chi2_contingency(X_train, _label_matrix(y_train))


def peak_snr_weight_aggregation(contingency_matrix):
    return get_engine(0).peak_snr(contingency_matrix)


def community_weighting(X_train, X_test, community_weights):
    if issparse(X_train):
        eng = get_engine(0)
        X_train = eng.community_weighting(X_train, community_weights)
        X_test = eng.community_weighting(X_test, community_weights)
    return X_train, X_test


def chi2_psnr_community_weighting(X_train, X_test, y_train):
    if issparse(X_train):
        if np.any(X_train.data < 0):
            raise ValueError("Input X must be non-negative.")
        # the n_classes x n_features contingency matrix never leaves the device
        community_weights = get_engine(0).chi2_psnr_weights(X_train, _label_matrix(y_train))
        X_train, X_test = community_weighting(X_train, X_test, community_weights)
    return X_train, X_test


class ResidentFeatures:
    """The feature matrix kept in HBM across the folds of an experiment.  Per fold,

        X_train, X_test = resident.chi2_psnr_community_weighting(train, test, y_train)

    replaces the reference's (experiments/utility.py:94-104)

        X_train, X_test = feature_matrix[train, :], feature_matrix[test, :]
        contingency_matrix = chi2_contingency_matrix(X_train, y_train)
        community_weights = peak_snr_weight_aggregation(contingency_matrix)
        X_train, X_test = community_weighting(X_train, X_test, community_weights)

    with the row gather, the contingency matrix and the weighting on the device; only the two
    weighted blocks are copied back.  `feature_matrix=None` adopts the matrix the last
    arcte() + Engine.normalize_features() left on the device.
    """

    def __init__(self, feature_matrix=None, device=0):
        self._eng = get_engine(device)
        self._eng.store_features(feature_matrix)
        self.shape = self._eng.stored_shape

    def chi2_psnr_community_weighting(self, train, test, y_train):
        return self._eng.weighted_fold(train, test, _label_matrix(y_train))
