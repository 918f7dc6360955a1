"""Time the native edge-list reader / feature writer (csrc/textio.cu) against the reference's
pure-Python loops restated inline (datautil/datarw.py:54-143).  Host only.

    python tools/io_bench.py [youtube|flickr|baNxM] [feature entries to write, millions]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sparse

from bench import make_graph
from reveal_graph_embedding_b200.io import read_adjacency_matrix, write_features

workload = sys.argv[1] if len(sys.argv) > 1 else "youtube"
entries_m = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
A = sparse.triu(make_graph(workload), k=1).tocoo()
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "edges.tsv")
with open(path, "w") as f:
    f.write("\n".join("%d\t%d\t1.0" % (a + 1, b + 1) for a, b in zip(A.row.tolist(), A.col.tolist())) + "\n")
size = os.path.getsize(path)
t = time.perf_counter(); M, ids = read_adjacency_matrix(path, "\t", True); t_native = time.perf_counter() - t


def python_read(path):
    id_to_node, row, col, data = {}, [], [], []
    for line in open(path):
        f = line.strip().split("\t")
        if f[0][0] == "#":
            continue
        s = id_to_node.setdefault(int(f[0]), len(id_to_node))
        t = id_to_node.setdefault(int(f[1]), len(id_to_node))
        w = float(f[2])
        row.append(s); col.append(t); data.append(w)
        if s != t:
            row.append(t); col.append(s); data.append(w)
    return np.array(row), np.array(col), np.array(data)


t = time.perf_counter(); r, c, d = python_read(path); t_py = time.perf_counter() - t
assert np.array_equal(r, M.row) and np.array_equal(c, M.col)
print("read  %s: %d edges, %.1f MB: native %.3f s (%.0f MB/s), python loop %.2f s -> %.0fx"
      % (workload, A.nnz, size / 1e6, t_native, size / 1e6 / t_native, t_py, t_py / t_native))

n = M.shape[0]
k = int(entries_m * 1e6)
rng = np.random.default_rng(0)
rows = np.sort(rng.integers(0, n, size=k))
X = sparse.csr_matrix((np.ones(k), (rows, rng.integers(0, 2 * n, size=k))), shape=(n, 2 * n))
out = os.path.join(tmp, "features.tsv")
t = time.perf_counter(); nbytes = write_features(out, X, "\t", ids); t_native = time.perf_counter() - t
sample = min(X.nnz, 2_000_000)
coo = sparse.coo_matrix(X)
t = time.perf_counter()
with open(out + ".py", "w") as f:
    for e in range(sample):
        f.write(str(ids[coo.row[e]]) + "\t" + str(coo.col[e]) + "\t" + str(int(coo.data[e])) + "\n")
t_py = (time.perf_counter() - t) * X.nnz / sample
print("write %d entries, %.1f MB: native %.3f s (%.0f MB/s), python loop %.1f s (extrapolated from %d entries) -> %.0fx"
      % (X.nnz, nbytes / 1e6, t_native, nbytes / 1e6 / t_native, t_py, sample, t_py / t_native))
