"""CPU-side checks of the boundary: the shared library loads and exports exactly the
symbols include/arcte_cuda.h declares; the ctypes layer mirrors the header; compute entry
points fail loudly (no CPU fallback) when no GPU is present."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arcte_cuda.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arcte_cuda_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from reveal_graph_embedding_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_library_exports_every_header_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    extra = sorted(s for s in exported if s.startswith("arcte_cuda_") and s not in header_symbols())
    assert not extra, "exported but undeclared: %r" % extra


def test_ctypes_binding_matches_header(lib):
    assert sorted(lib.SYMBOLS) == header_symbols()
    L = lib.load()
    for s in lib.SYMBOLS:
        assert hasattr(L, s)


def test_stats_struct_layout_matches_header(lib):
    src = open(HEADER).read()
    body = re.search(r"typedef struct arcte_cuda_stats \{(.*?)\} arcte_cuda_stats;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int64_t|double)\s+([a-z_0-9]+)\s*;", body)
    got = [(("int64_t" if t is lib.C.c_int64 else "double"), n) for n, t in lib.Stats._fields_]
    assert got == fields


def test_no_cpu_fallback(lib):
    """Without a GPU every compute entry point raises; nothing silently runs on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.engine import ArcteCudaError
    import scipy.sparse as sparse
    A = sparse.csr_matrix(np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float))
    with pytest.raises(ArcteCudaError):
        arcte(A, 0.1, 1e-5, 1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no file of the package may reference it."""
    pkg = os.path.join(ROOT, "reveal_graph_embedding_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "arcte_oracle" not in text and "liboracle" not in text, os.path.join(dirpath, f)
                assert "weighting_oracle" not in text, os.path.join(dirpath, f)
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_last_error_is_reported(lib):
    L = lib.load()
    rc = L.arcte_cuda_get_stats(None, None)
    assert rc == -2
    assert b"null" in L.arcte_cuda_last_error()
