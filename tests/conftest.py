import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        from reveal_graph_embedding_b200.engine import device_count
        return device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a B200 skips the gpu tests instead of erroring (the product itself
    still fails loudly there: tests/test_abi.py::test_no_cpu_fallback)."""
    if not any("gpu" in it.keywords for it in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import arcte_oracle
    arcte_oracle.build()
    return arcte_oracle
