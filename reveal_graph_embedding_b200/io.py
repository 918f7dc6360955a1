"""Edge-list reader and feature writer with the reference's wire formats
(reveal_graph_embedding/datautil/datarw.py:54-143), same names and arguments.

The reference parses the edge list line by line and writes the features entry by entry in
Python; around a one-second GPU extraction that is the whole wall time of the `arcte`
console script.  Both run in the native library (csrc/textio.cu, all host threads) through
the C ABI: arcte_cuda_io_read_edge_list / arcte_cuda_io_write_features.
"""
import ctypes as C

import numpy as np
import scipy.sparse as spsp

from . import _lib
from ._lib import check, ptr


class NodeIds(dict):
    """node_to_id of datarw.py:110-111 (anonymised index -> original node id), kept as one
    int64 array so a million-node mapping costs no Python objects until it is looked at."""

    def __init__(self, ids):
        super().__init__()
        self.array = np.ascontiguousarray(ids, dtype=np.int64)
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            super().update(zip(range(self.array.size), self.array.tolist()))

    def __getitem__(self, k):
        return int(self.array[k])

    def __len__(self):
        return int(self.array.size)

    def __contains__(self, k):
        return isinstance(k, (int, np.integer)) and 0 <= k < self.array.size

    def __iter__(self):
        return iter(range(self.array.size))

    # every other dict read goes through the real dict, filled on first use (dict.get / copy / setdefault / pop
    # bypass __getitem__, so they must see the entries: the reference returns a plain dict)
    def get(self, k, default=None):
        return self[k] if k in self else default

    def copy(self):
        self._fill()
        return dict(super().items())

    def setdefault(self, k, default=None):
        self._fill()
        return super().setdefault(k, default)

    def pop(self, *a):
        self._fill()
        return super().pop(*a)

    def popitem(self):
        self._fill()
        return super().popitem()

    def __setitem__(self, k, v):
        self._fill()
        super().__setitem__(k, v)

    def __delitem__(self, k):
        self._fill()
        super().__delitem__(k)

    def update(self, *a, **kw):
        self._fill()
        super().update(*a, **kw)

    def keys(self):
        self._fill()
        return super().keys()

    def values(self):
        self._fill()
        return super().values()

    def items(self):
        self._fill()
        return super().items()

    def __eq__(self, other):
        self._fill()
        return dict.__eq__(self, other)

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)


def read_adjacency_matrix(file_path, separator, undirected, number_of_threads=0):
    """datarw.py:54-120: `src<sep>dst<sep>weight` rows, '#' comments; node ids are remapped
    to 0..n-1 in first-seen order (source before target).  Returns (COO matrix, node_to_id)."""
    L = _lib.load()
    h, n, m = C.c_void_p(), C.c_int64(), C.c_int64()
    check(L.arcte_cuda_io_read_edge_list(str(file_path).encode(), str(separator).encode(), int(bool(undirected)),
                                         int(number_of_threads), C.byref(h), C.byref(n), C.byref(m)))
    try:
        row = np.empty(max(m.value, 1), dtype=np.int64)
        col = np.empty(max(m.value, 1), dtype=np.int64)
        data = np.empty(max(m.value, 1), dtype=np.float64)
        ids = np.empty(max(n.value, 1), dtype=np.int64)
        check(L.arcte_cuda_io_edge_list_copy(h, ptr(row), ptr(col), ptr(data), ptr(ids)))
    finally:
        L.arcte_cuda_io_edge_list_free(h)
    k, nn = m.value, n.value
    adjacency_matrix = spsp.coo_matrix((data[:k], (row[:k], col[:k])), shape=(nn, nn))
    return adjacency_matrix, NodeIds(ids[:nn])


def write_features(file_path, features, separator, node_to_id, number_of_threads=0):
    """datarw.py:123-143: one `node_id<sep>community_id<sep>int(value)` row per stored entry,
    in the COO order of the CSR (row-major).  Returns the number of bytes written."""
    X = spsp.csr_matrix(features)
    if isinstance(node_to_id, NodeIds):
        ids = node_to_id.array
    else:
        ids = np.array([node_to_id[i] for i in range(X.shape[0])], dtype=np.int64)
    if ids.size != X.shape[0]:
        raise ValueError("node_to_id must map every row of the feature matrix")
    indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(X.indices, dtype=np.int32)
    data = np.ascontiguousarray(X.data, dtype=np.float64)
    nbytes = C.c_int64()
    check(_lib.load().arcte_cuda_io_write_features(str(file_path).encode(), str(separator).encode(), X.shape[0],
                                                   ptr(indptr), ptr(indices) if indices.size else None,
                                                   ptr(data) if data.size else None,
                                                   ptr(np.ascontiguousarray(ids)) if ids.size else None,
                                                   int(number_of_threads), C.byref(nbytes)))
    return nbytes.value
