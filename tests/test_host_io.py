"""Host-side text formats (datautil/datarw.py:54-143 of the reference) against files the
reference itself read and wrote (tests/golden/make_golden.py, cli_fixture)."""
import os

import numpy as np
import pytest
import scipy.sparse as sparse

from helpers import GOLDEN_DIR


def test_read_adjacency_matrix_matches_reference():
    from reveal_graph_embedding_b200.io import read_adjacency_matrix
    z = np.load(os.path.join(GOLDEN_DIR, "cli_adjacency.npz"))
    A, node_to_id = read_adjacency_matrix(os.path.join(GOLDEN_DIR, "cli_edges.tsv"), "\t", False)
    n = int(z["n"])
    assert A.shape == (n, n)
    want = sparse.coo_matrix((z["data"], (z["row"], z["col"])), shape=(n, n)).tocsr()
    got = sparse.csr_matrix(A)
    assert (got != want).nnz == 0 and np.array_equal(got.data, want.data)
    assert [node_to_id[i] for i in range(n)] == z["node_ids"].tolist()  # first-seen numbering


def test_read_undirected_adds_reciprocal_edges(tmp_path):
    from reveal_graph_embedding_b200.io import read_adjacency_matrix
    p = tmp_path / "e.csv"
    p.write_text("10,20,1.5\n20,30,2.0\n30,30,4.0\n")
    A, ids = read_adjacency_matrix(str(p), ",", True)
    D = sparse.csr_matrix(A).toarray()
    assert np.array_equal(D, np.array([[0, 1.5, 0], [1.5, 0, 2.0], [0, 2.0, 4.0]]))  # self loop not doubled
    assert ids == {0: 10, 1: 20, 2: 30}


def test_write_features_matches_reference_bytes(tmp_path):
    """Feed the reference's own feature matrix (parsed back from its file) through our writer."""
    from reveal_graph_embedding_b200.io import read_adjacency_matrix, write_features
    z = np.load(os.path.join(GOLDEN_DIR, "cli_adjacency.npz"))
    ids = z["node_ids"]
    inv = {int(v): i for i, v in enumerate(ids)}
    rows, cols, vals = [], [], []
    ref_text = open(os.path.join(GOLDEN_DIR, "cli_features.tsv")).read()
    for ln in ref_text.splitlines():
        a, b, c = ln.split("\t")
        rows.append(inv[int(a)])
        cols.append(int(b))
        vals.append(float(c))
    n = int(z["n"])
    X = sparse.csr_matrix((vals, (rows, cols)), shape=(n, 2 * n))
    out = tmp_path / "f.tsv"
    write_features(str(out), X, "\t", {i: int(v) for i, v in enumerate(ids)})
    assert out.read_text() == ref_text


@pytest.mark.gpu
def test_cli_end_to_end_matches_reference_file(tmp_path):
    from reveal_graph_embedding_b200.entry_points.arcte import main
    out = tmp_path / "features.tsv"
    main(["-i", os.path.join(GOLDEN_DIR, "cli_edges.tsv"), "-o", str(out), "-nt", "1"])
    assert out.read_text() == open(os.path.join(GOLDEN_DIR, "cli_features.tsv")).read()
