"""Edge-list reader and feature writer with the reference's wire formats
(reveal_graph_embedding/datautil/datarw.py:54-143), vectorised with numpy."""
import numpy as np
import scipy.sparse as spsp


def read_adjacency_matrix(file_path, separator, undirected):
    """datarw.py:54-120: `src<sep>dst<sep>weight` rows, '#' comments; node ids are remapped
    to 0..n-1 in first-seen order (source before target).  Returns (COO matrix, node_to_id)."""
    src, dst, wgt = [], [], []
    with open(file_path, "r") as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            parts = line.split(separator)
            src.append(int(parts[0]))
            dst.append(int(parts[1]))
            wgt.append(float(parts[2]))
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    wgt = np.asarray(wgt, dtype=np.float64)
    # first-seen order over the interleaved stream s0, t0, s1, t1, ...
    inter = np.empty(2 * src.size, dtype=np.int64)
    inter[0::2] = src
    inter[1::2] = dst
    uniq, first = np.unique(inter, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(uniq.size, dtype=np.int64)
    rank[order] = np.arange(uniq.size)
    row = rank[np.searchsorted(uniq, src)]
    col = rank[np.searchsorted(uniq, dst)]
    node_to_id = {int(i): int(uniq[order[i]]) for i in range(uniq.size)}
    if undirected:
        m = row != col
        row, col, wgt = (np.concatenate([row, col[m]]), np.concatenate([col, row[m]]),
                         np.concatenate([wgt, wgt[m]]))
    n = uniq.size
    return spsp.coo_matrix((wgt, (row, col)), shape=(n, n)), node_to_id


def write_features(file_path, features, separator, node_to_id):
    """datarw.py:123-143: one `node_id<sep>community_id<sep>int(value)` row per stored entry,
    in COO order of the CSR."""
    features = spsp.coo_matrix(features)
    ids = np.array([node_to_id[i] for i in range(features.shape[0])], dtype=np.int64)
    node = ids[features.row]
    with open(file_path, "w") as f:
        for a, b, c in zip(node.tolist(), features.col.tolist(), features.data.astype(np.int64).tolist()):
            f.write("%d%s%d%s%d\n" % (a, separator, b, separator, c))
