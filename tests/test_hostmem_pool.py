"""Host logic of the page-locked result pool (reveal_graph_embedding_b200/hostmem.py) with the CUDA
allocator replaced by malloc: blocks of ones are handed out only when ready, refilled after a
caller releases them (even if the caller scribbled over them), never more than four alive, and a
miss does not trigger unbounded allocation."""
import ctypes as C
import gc

import numpy as np
import pytest


class FakeLib:
    """arcte_cuda_host_alloc / _free / _fill_f64 on plain heap memory."""

    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]
        self.allocs = 0
        self.live = set()

    def arcte_cuda_host_alloc(self, out, nbytes):
        p = self.libc.malloc(nbytes)
        out._obj.value = p
        self.allocs += 1
        self.live.add(p)
        return 0

    def arcte_cuda_host_free(self, p):
        self.live.discard(p.value)
        self.libc.free(p)
        return 0

    def arcte_cuda_host_fill_f64(self, p, count, value, n_threads):
        np.frombuffer((C.c_char * (count * 8)).from_address(p.value), dtype=np.float64)[:] = value
        return 0


@pytest.fixture
def pool(monkeypatch):
    from reveal_graph_embedding_b200 import _lib, hostmem
    fake = FakeLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setenv("ARCTE_CUDA_PINNED_RESULTS", "1")
    hostmem.wait_idle()
    hostmem._free.clear(); hostmem._ones_free.clear(); hostmem._pending.clear(); hostmem._pending_ones.clear()
    hostmem._ones_alive[0] = 0
    monkeypatch.setattr(hostmem, "_pooled_bytes", 0)
    yield hostmem, fake
    hostmem.wait_idle()
    hostmem.drain()
    hostmem._ones_alive[0] = 0


def test_ones_block_is_prepared_in_the_background_then_reused(pool):
    hostmem, fake = pool
    n = 1 << 18                                   # 2 MB of doubles: above the pinning threshold
    assert hostmem.ones(n) is None                # nothing ready: the caller copies as usual
    hostmem.start_pending()
    hostmem.wait_idle()
    a = hostmem.ones(n)
    assert a is not None and a.size == n and np.all(a == 1.0)
    a[:] = -5.0                                   # a caller scribbles over its result ...
    del a
    gc.collect()
    hostmem.wait_idle()
    b = hostmem.ones(n)                           # ... the block comes back refilled
    assert b is not None and np.all(b == 1.0)
    assert fake.allocs == 1                       # the same block, no new allocation


def test_small_requests_and_disabled_pool_fall_back(pool, monkeypatch):
    hostmem, fake = pool
    assert hostmem.ones(100) is None              # below the pinning threshold: plain copy path
    assert not hostmem._pending_ones
    monkeypatch.setenv("ARCTE_CUDA_ONES_POOL", "0")
    assert hostmem.ones(1 << 18) is None and not hostmem._pending_ones


def test_misses_do_not_allocate_without_bound(pool):
    hostmem, fake = pool
    n = 1 << 18
    held = []
    for _ in range(12):                           # a caller that keeps every result alive
        a = hostmem.ones(n)
        if a is not None:
            held.append(a)
        hostmem.start_pending()
        hostmem.wait_idle()
    assert fake.allocs <= hostmem._ONES_BLOCK_LIMIT
    assert len(held) <= hostmem._ONES_BLOCK_LIMIT
    for a in held:
        assert np.all(a == 1.0)


def test_plain_pinned_blocks_are_recycled(pool):
    hostmem, fake = pool
    n = 1 << 19
    a = hostmem.empty(n, np.int32)                # first request: pageable, a block is pinned for next time
    assert a.size == n
    hostmem.start_pending()
    hostmem.wait_idle()
    b = hostmem.empty(n, np.int32)
    assert b.size == n and fake.allocs == 1
    del b
    gc.collect()
    c = hostmem.empty(n, np.int32)
    assert fake.allocs == 1                       # recycled, not re-pinned
