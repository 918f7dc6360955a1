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


# ---- native parser / formatter (csrc/textio.cu) against the reference's Python semantics ----
def _python_read(path, separator, undirected):
    """datarw.py:54-120 restated with plain Python (line.strip().split(separator), int, float)."""
    id_to_node, row, col, data = {}, [], [], []
    for line in open(path):
        f = line.strip().split(separator)
        if not f[0] or f[0][0] == "#":
            continue
        s = id_to_node.setdefault(int(f[0]), len(id_to_node))
        t = id_to_node.setdefault(int(f[1]), len(id_to_node))
        w = float(f[2])
        row.append(s); col.append(t); data.append(w)
        if undirected and s != t:
            row.append(t); col.append(s); data.append(w)
    ids = [None] * len(id_to_node)
    for k, v in id_to_node.items():
        ids[v] = k
    return np.array(row), np.array(col), np.array(data), ids


@pytest.mark.parametrize("separator,undirected,threads", [("\t", False, 0), (",", True, 3), ("::", False, 7),
                                                          (" ", True, 1)])
def test_native_reader_equals_python_semantics(tmp_path, separator, undirected, threads):
    from reveal_graph_embedding_b200.io import read_adjacency_matrix
    rng = np.random.default_rng(len(separator) * 7 + threads)
    ids = np.unique(rng.integers(-50, 10 ** 12, size=8000))[:4000]
    lines = ["# header", "#x%sy%sz" % (separator, separator)]
    fmts = ["%d", "%.3f", "%.17g", "%e"]
    for i in range(60000):                       # ~1.5 MB: many 64 KB parser chunks
        a, b = rng.choice(ids, size=2)
        w = fmts[i % 4] % (rng.uniform(0.01, 50.0) if i % 4 else rng.integers(1, 9))
        extra = (separator + "ignored") if i % 97 == 0 else ""
        pad = "  " if i % 31 == 0 and separator != " " else ""
        lines.append("%s%d%s%d%s%s%s%s" % (pad, a, separator, b, separator, w, extra, pad))
        if i % 5000 == 0:
            lines.append("")                     # blank line (skipped)
    p = tmp_path / "edges.txt"
    p.write_text("\n".join(lines))               # no trailing newline on the last row
    A, node_to_id = read_adjacency_matrix(str(p), separator, undirected, number_of_threads=threads)
    row, col, data, want_ids = _python_read(str(p), separator, undirected)
    assert np.array_equal(A.row, row) and np.array_equal(A.col, col) and np.array_equal(A.data, data)
    assert A.shape == (len(want_ids), len(want_ids))
    assert [node_to_id[i] for i in range(len(want_ids))] == want_ids
    assert dict(node_to_id.items()) == dict(enumerate(want_ids))


def test_native_reader_reports_the_bad_line(tmp_path):
    from reveal_graph_embedding_b200._lib import ArcteCudaError
    from reveal_graph_embedding_b200.io import read_adjacency_matrix
    p = tmp_path / "bad.tsv"
    p.write_text("1\t2\t1.0\n# fine\n3\tfour\t1.0\n")
    with pytest.raises(ArcteCudaError, match="line 3"):
        read_adjacency_matrix(str(p), "\t", False)
    with pytest.raises(ArcteCudaError, match="cannot open"):
        read_adjacency_matrix(str(tmp_path / "missing.tsv"), "\t", False)
    q = tmp_path / "short.tsv"
    q.write_text("1\t2\n")
    with pytest.raises(ArcteCudaError, match="line 1"):
        read_adjacency_matrix(str(q), "\t", False)


def test_native_reader_empty_file(tmp_path):
    from reveal_graph_embedding_b200.io import read_adjacency_matrix
    p = tmp_path / "empty.tsv"
    p.write_text("# nothing\n")
    A, ids = read_adjacency_matrix(str(p), "\t", False)
    assert A.shape == (0, 0) and A.nnz == 0 and len(ids) == 0


@pytest.mark.parametrize("threads", [0, 1, 5])
def test_native_writer_equals_python_semantics(tmp_path, threads):
    from reveal_graph_embedding_b200.io import write_features
    rng = np.random.default_rng(threads)
    n = 3000
    X = sparse.random(n, 2 * n, density=0.02, random_state=rng, format="csr",
                      data_rvs=lambda k: rng.choice([1.0, 2.0, 0.0, -3.7, 1e6 + 0.9, 123456789012.0], size=k))
    X = sparse.lil_matrix(X)
    X[17, :] = 0                                  # an empty row in the middle
    X = sparse.csr_matrix(X)
    X.data[::11] = 0.0                            # explicit zeros are written too
    ids = np.unique(rng.integers(-10 ** 6, 10 ** 15, size=2 * n))[:n]
    rng.shuffle(ids)
    node_to_id = {i: int(v) for i, v in enumerate(ids)}
    out = tmp_path / "f.txt"
    nbytes = write_features(str(out), X, ";;", node_to_id, number_of_threads=threads)
    coo = sparse.coo_matrix(X)                    # datarw.py:127-143
    want = "".join(str(node_to_id[r]) + ";;" + str(c) + ";;" + str(int(v)) + "\n"
                   for r, c, v in zip(coo.row.tolist(), coo.col.tolist(), coo.data.tolist()))
    got = out.read_text()
    assert got == want and nbytes == len(want)


def test_native_writer_rejects_non_finite(tmp_path):
    from reveal_graph_embedding_b200._lib import ArcteCudaError
    from reveal_graph_embedding_b200.io import write_features
    X = sparse.csr_matrix(np.array([[1.0, np.nan], [0.0, 2.0]]))
    with pytest.raises(ArcteCudaError, match="not a finite"):
        write_features(str(tmp_path / "f.txt"), X, "\t", {0: 5, 1: 6})


# ---- host-side input handling (engine.canonical_csr; arcte.py:601 csr_matrix(adjacency_matrix)) ----
def test_canonical_csr_host_logic():
    from reveal_graph_embedding_b200.engine import canonical_csr
    rng = np.random.default_rng(3)
    A = sparse.random(50, 50, density=0.1, random_state=rng, format="csr")
    A.sum_duplicates(); A.sort_indices()
    assert canonical_csr(A) is A                                   # canonical float64 CSR: used as is
    for conv in (sparse.coo_matrix, sparse.csc_matrix, sparse.lil_matrix, lambda M: M.astype(np.float32)):
        B = canonical_csr(conv(A))
        assert sparse.isspmatrix_csr(B) and B.dtype == np.float64 and B.has_canonical_format
        assert (B != A).nnz == 0 or conv is not sparse.coo_matrix and np.allclose(B.toarray(), A.toarray())
    # unsorted indices + a duplicate entry: canonicalised on a copy, the caller's arrays untouched
    rows, cols, vals = [0, 0, 0, 1], [3, 1, 3, 2], [1.0, 2.0, 4.0, 5.0]
    D = sparse.csr_matrix((np.array(vals), np.array(cols), np.array([0, 3, 4, 4, 4])), shape=(4, 4))
    before = (D.indices.copy(), D.data.copy(), D.indptr.copy())
    E = canonical_csr(D)
    assert E is not D and E.has_canonical_format
    assert np.array_equal(E.toarray(), np.array([[0, 2.0, 0, 5.0], [0, 0, 5.0, 0], [0] * 4, [0] * 4]))
    assert all(np.array_equal(a, b) for a, b in zip(before, (D.indices, D.data, D.indptr)))
    with pytest.raises(ValueError, match="square"):
        canonical_csr(sparse.csr_matrix((3, 4)))


def test_node_ids_behaves_like_the_reference_dict():
    """datarw.py:110-111 returns a plain dict; NodeIds must answer every dict read the same way."""
    from reveal_graph_embedding_b200.io import NodeIds
    d = NodeIds([20, 30, 40])
    assert d[1] == 30 and d.get(1) == 30 and d.get(9) is None and d.get(9, -1) == -1
    assert len(d) == 3 and 2 in d and 3 not in d and list(d) == [0, 1, 2]
    assert d.copy() == {0: 20, 1: 30, 2: 40} and type(d.copy()) is dict
    assert d.setdefault(0, 99) == 20 and dict(d.items()) == {0: 20, 1: 30, 2: 40}
    assert d == {0: 20, 1: 30, 2: 40}


def test_cli_undirected_flag_forms():
    """`-u` alone, `-u true`, `-u False` (the reference's type=bool reads the last one as True)."""
    import argparse
    from reveal_graph_embedding_b200.entry_points import arcte as cli
    parser = argparse.ArgumentParser()
    for short, long_, dest, typ, default, required, text in cli._FLAGS:
        if typ == "flag":
            parser.add_argument(short, long_, dest=dest, nargs="?", const=True, default=default, type=cli._to_bool)
        else:
            parser.add_argument(short, long_, dest=dest, type=typ, default=default, required=required)
    base = ["-i", "a", "-o", "b"]
    assert parser.parse_args(base).undirected is False
    assert parser.parse_args(base + ["-u"]).undirected is True
    assert parser.parse_args(base + ["-u", "true"]).undirected is True
    assert parser.parse_args(base + ["-u", "False"]).undirected is False


def _engine_shell(A):
    """An Engine with the host-side copies set_graph keeps, without a device (self-loop logic is pure numpy)."""
    from reveal_graph_embedding_b200.engine import Engine
    e = Engine.__new__(Engine)
    e.n, e.nnz, e._loops = int(A.shape[0]), int(A.nnz), None
    e._indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    e._indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    return e


def test_self_loop_rows_and_patch_match_the_reference_base_block():
    """arcte.py:676-679: base = I + pattern(A) stores 2.0 where A stores its diagonal.  The host patches those
    entries into a value array of ones; positions checked against scipy's own I + pattern(A), whole matrix and
    row blocks, from the main thread and from a side thread (as arcte() runs it)."""
    import threading
    from reveal_graph_embedding_b200 import graphs
    A = graphs.barabasi_albert(500, 3, seed=3).tolil()
    loops = [0, 1, 17, 250, 499]
    for i in loops:
        A[i, i] = 0.5
    A = sparse.csr_matrix(A)
    A.sort_indices()
    n = A.shape[0]
    pattern = sparse.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    base = (sparse.identity(n, format="csr") + pattern).tocsr()
    base.sort_indices()
    e = _engine_shell(A)
    th = threading.Thread(target=e.self_loop_rows)
    th.start()
    th.join()
    rows, rank = e._loops
    assert rows.tolist() == loops
    data = np.ones(base.nnz)
    e.patch_self_loops(data, base.indptr.astype(np.int64))
    assert np.array_equal(data, base.data)
    # row blocks, the way the multi-GPU paths patch their slices
    data2 = np.ones(base.nnz)
    for lo, hi in ((0, 100), (100, 260), (260, 500)):
        o0, o1 = base.indptr[lo], base.indptr[hi]
        e.patch_self_loops(data2[o0:o1], (base.indptr[lo:hi + 1] - o0).astype(np.int64), lo, hi)
    assert np.array_equal(data2, base.data)
    # no self loops: nothing is touched
    B = sparse.csr_matrix(graphs.barabasi_albert(200, 2, seed=4))
    eb = _engine_shell(B)
    ones = np.ones(10)
    eb.patch_self_loops(ones, np.zeros(B.shape[0] + 1, dtype=np.int64))
    assert eb._loops[0].size == 0 and np.all(ones == 1.0)


def test_ones_mapping_is_ordinary_writable_memory():
    """hostmem.ones: copy-on-write mappings of one block of ones -- reads 1.0 everywhere, writes stay private to
    the array, a second array is unaffected, small arrays are plain numpy."""
    from reveal_graph_embedding_b200 import hostmem
    count = (5 << 20) + 12345          # > 32 MB of doubles: several mappings of the template, ragged end
    a = hostmem.ones(count)
    assert a.dtype == np.float64 and a.shape == (count,) and a.flags.writeable
    assert a[0] == 1.0 and a[-1] == 1.0 and float(a.sum()) == float(count)
    a[::4096] = 2.0
    a[-1] = 7.0
    b = hostmem.ones(count)
    assert float(b.sum()) == float(count) and b[-1] == 1.0 and b[0] == 1.0
    assert a[4096] == 2.0 and a[-1] == 7.0
    small = hostmem.ones(100)
    assert isinstance(small, np.ndarray) and small.sum() == 100.0
    del a, b


def test_arcte_single_gpu_host_flow_with_a_fake_engine(monkeypatch):
    """Order of the library calls behind arcte() on one GPU, and the side thread that finds the self-loop rows
    while the device walks: started only for large inputs, always joined before features()."""
    import threading
    import reveal_graph_embedding_b200.embedding.arcte.arcte as mod

    class FakeEngine:
        def __init__(self):
            self.calls, self.loop_threads = [], []

        def set_graph(self, A, canonical=False):
            self.calls.append("set_graph")

        def self_loop_rows(self):
            self.loop_threads.append(threading.current_thread() is threading.main_thread())

        def extract(self, rule, rho, eps):
            self.calls.append("extract")

        def assemble(self):
            self.calls.append("assemble")

        def features(self):
            self.calls.append("features")
            return "X"

    A = sparse.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    for min_nnz, threaded in ((1 << 20, False), (1, True)):
        eng = FakeEngine()
        monkeypatch.setattr(mod, "get_engine", lambda d=0, e=eng: e)
        monkeypatch.setattr(mod, "device_count", lambda: 1)
        monkeypatch.setattr(mod, "_SIDE_THREAD_MIN_NNZ", min_nnz)
        assert mod.arcte(A, 0.1, 1e-5) == "X"
        assert eng.calls == ["set_graph", "extract", "assemble", "features"]
        assert eng.loop_threads == ([False] if threaded else [])   # ran off the main thread, or not at all here

    # a failing extract still joins the thread and propagates
    class Failing(FakeEngine):
        def extract(self, rule, rho, eps):
            raise RuntimeError("boom")
    eng = Failing()
    monkeypatch.setattr(mod, "get_engine", lambda d=0, e=eng: e)
    monkeypatch.setattr(mod, "_SIDE_THREAD_MIN_NNZ", 1)
    import pytest
    with pytest.raises(RuntimeError, match="boom"):
        mod.arcte(A, 0.1, 1e-5)
    assert threading.active_count() == 1 or all(not t.name.startswith("Thread-") or not t.is_alive()
                                                 for t in threading.enumerate() if t is not threading.main_thread())
