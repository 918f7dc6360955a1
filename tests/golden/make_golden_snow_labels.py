"""Copies the REAL SNOW2014Graph label matrix of the reference (snow2014graph/user_label_matrix.tsv:
533,874 nodes, 90 labels, 27,863 stored entries, 1-based `row<TAB>label<TAB>1` lines after a header) into a
small fixture, so that the config-1 stand-in graph (tools/snow_standin.py) can be planted from it on the GPU box,
where /root/reference does not exist.  The graph file of that dataset (men_ret_graph.tsv) is absent from the
reference tree (.MISSING_LARGE_BLOBS:1).

    python tests/golden/make_golden_snow_labels.py [/root/reference]
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
path = os.path.join(ref, "snow2014graph", "user_label_matrix.tsv")
with open(path) as f:
    header = f.readline().split("\t")
    n_rows, n_cols, nnz = int(header[1]), int(header[3]), int(header[5])
    rows, cols = [], []
    for line in f:
        a = line.split("\t")
        rows.append(int(a[0]) - 1)      # 1-based in the file (snow_read_data.py:140-175)
        cols.append(int(a[1]) - 1)
rows, cols = np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int8)
assert rows.size == nnz and rows.max() < n_rows and cols.max() < n_cols
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "snow_labels.npz")
np.savez_compressed(out, n_rows=n_rows, n_cols=n_cols, rows=rows, cols=cols)
print("wrote %s: %d nodes, %d labels, %d entries, %d labelled nodes" % (out, n_rows, n_cols, nnz, np.unique(rows).size))
