"""Page-locked host buffers for results, recycled through a small pool.

The feature matrix of a large graph is gigabytes; copying it into ordinary (pageable)
numpy arrays goes through the CUDA driver's staging buffer at a few GB/s, while a
page-locked destination is filled by DMA at PCIe rate.  Page-locking itself is slow, so
blocks are pooled: a numpy array handed to the caller owns its block, and when the last
reference to the array dies the block returns to the pool for the next call (a caller
that keeps its result simply keeps the block).  ARCTE_CUDA_PINNED_RESULTS=0 switches to
plain numpy arrays.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import _lib

_POOL_LIMIT_BYTES = int(os.environ.get("ARCTE_CUDA_PINNED_POOL_GB", "48")) << 30
_MIN_PINNED_BYTES = 1 << 20  # tiny results are not worth a pinned block

_lock = threading.Lock()
_free = []        # [(bytes, address)]
_pooled_bytes = 0
_threads = []
_pending = []


def enabled():
    return os.environ.get("ARCTE_CUDA_PINNED_RESULTS", "1") != "0"


def _release(addr, nbytes):
    global _pooled_bytes
    with _lock:
        if _pooled_bytes + nbytes <= _POOL_LIMIT_BYTES:
            _free.append((nbytes, addr))
            _pooled_bytes += nbytes
            return
    _lib.load().arcte_cuda_host_free(C.c_void_p(addr))


def _acquire(nbytes):
    global _pooled_bytes
    with _lock:
        best = None
        for i, (b, a) in enumerate(_free):
            if b >= nbytes and b <= nbytes + (nbytes >> 2) + 4096 and (best is None or b < _free[best][0]):
                best = i
        if best is not None:
            b, a = _free.pop(best)
            _pooled_bytes -= b
            return a, b
    return None, 0


def _prepin(nbytes):
    """Background: page-lock a block of this size for the NEXT call (pinning runs at a few
    GB/s, slower than one pageable copy, so the call that first needs a size never waits)."""
    global _pooled_bytes
    with _lock:
        if _pooled_bytes + nbytes > _POOL_LIMIT_BYTES:
            return
    p = C.c_void_p()
    if _lib.load().arcte_cuda_host_alloc(C.byref(p), int(nbytes)) != 0:
        return
    _release(p.value, nbytes)


def empty(count, dtype):
    """Uninitialised 1-D array of `count` items: a pooled page-locked block when one of
    the right size is free, else a plain numpy array (and a block is pinned in the
    background for next time)."""
    dtype = np.dtype(dtype)
    nbytes = int(count) * dtype.itemsize
    if not enabled() or nbytes < _MIN_PINNED_BYTES:
        return np.empty(int(count), dtype=dtype)
    addr, block = _acquire(nbytes)
    if os.environ.get("ARCTE_CUDA_DEBUG"):
        import sys
        print("[arcte] hostmem.empty(%d bytes): %s" % (nbytes, "pooled pinned block" if addr else "pageable"),
              file=sys.stderr)
    if addr is None:
        _pending.append(nbytes)  # pinned in the background once the caller's copy is done
        return np.empty(int(count), dtype=dtype)
    buf = (C.c_char * block).from_address(addr)
    weakref.finalize(buf, _release, addr, block).atexit = False   # nothing to recycle at interpreter exit
    return np.frombuffer(buf, dtype=dtype, count=int(count))


# ---- blocks pre-filled with 1.0 -----------------------------------------------------------------
# Every stored value of an ARCTE feature matrix is 1.0 except the diagonal entries of self loops
# (arcte.py:676-679: I + pattern(A); local block np.ones_like, arcte.py:379-381), and the values
# are two thirds of the bytes of the result.  A pooled page-locked block that already holds
# ones lets the caller skip the device-to-host copy of the values altogether (the few 2.0
# entries are patched on the host).  A block that comes back from a caller may have been
# modified, so it is refilled in the background before it is offered again.
_ones_free = []      # [(bytes, address)]
_pending_ones = []
counters = {"ones_hits": 0, "ones_misses": 0}


_ONES_BLOCK_LIMIT = 4   # blocks of ones alive at any time (pooled, handed out or being refilled)
_ones_alive = [0]


def _fill_ones(addr, nbytes):
    # native, multi-threaded, GIL released (ctypes): a 4 GB block takes a fraction of a second
    _lib.load().arcte_cuda_host_fill_f64(C.c_void_p(addr), nbytes // 8, 1.0, 0)


def _release_ones(addr, nbytes):
    t = threading.Thread(target=_refill_and_pool, args=(addr, nbytes), daemon=True)
    t.start()
    _threads.append(t)


def _refill_and_pool(addr, nbytes):
    global _pooled_bytes
    _fill_ones(addr, nbytes)
    with _lock:
        if _pooled_bytes + nbytes <= _POOL_LIMIT_BYTES:
            _ones_free.append((nbytes, addr))
            _pooled_bytes += nbytes
            return
        _ones_alive[0] -= 1
    _lib.load().arcte_cuda_host_free(C.c_void_p(addr))


def _make_ones(nbytes):
    with _lock:
        if _pooled_bytes + nbytes > _POOL_LIMIT_BYTES or _ones_alive[0] >= _ONES_BLOCK_LIMIT:
            return
        _ones_alive[0] += 1
    p = C.c_void_p()
    if _lib.load().arcte_cuda_host_alloc(C.byref(p), int(nbytes)) != 0:
        with _lock:
            _ones_alive[0] -= 1
        return
    _refill_and_pool(p.value, nbytes)


def ones(count):
    """A float64 array of `count` ones in a pooled page-locked block, or None when no such block
    is ready (one is then prepared in the background for the next call)."""
    global _pooled_bytes
    nbytes = int(count) * 8
    if not enabled() or nbytes < _MIN_PINNED_BYTES or os.environ.get("ARCTE_CUDA_ONES_POOL", "1") == "0":
        return None
    with _lock:
        best = None
        for i, (b, a) in enumerate(_ones_free):
            if b >= nbytes and b <= nbytes + (nbytes >> 2) + 4096 and (best is None or b < _ones_free[best][0]):
                best = i
        if best is not None:
            block, addr = _ones_free.pop(best)
            _pooled_bytes -= block
        else:
            addr = None
    if addr is None:
        if _ones_alive[0] + len(_pending_ones) < _ONES_BLOCK_LIMIT:   # else one is being refilled: wait for it
            _pending_ones.append(nbytes + (nbytes >> 4))               # head-room: results of nearby sizes reuse it
        counters["ones_misses"] += 1
        return None
    counters["ones_hits"] += 1
    buf = (C.c_char * block).from_address(addr)
    weakref.finalize(buf, _release_ones, addr, block).atexit = False   # (would start a thread during shutdown)
    return np.frombuffer(buf, dtype=np.float64, count=int(count))


def start_pending():
    """Start page-locking blocks for the sizes that missed the pool (called after the
    device-to-host copy that used the pageable fallback has finished, so the two do not
    compete for the host's memory system)."""
    while _pending:
        t = threading.Thread(target=_prepin, args=(_pending.pop(),), daemon=True)
        t.start()
        _threads.append(t)
    while _pending_ones:
        t = threading.Thread(target=_make_ones, args=(_pending_ones.pop(),), daemon=True)
        t.start()
        _threads.append(t)


def wait_idle():
    """Block until every background pinning thread has finished (benchmarks, tests)."""
    while _threads:
        _threads.pop().join()


def drain():
    """Free every pooled block (tests)."""
    global _pooled_bytes
    with _lock:
        blocks = list(_free) + list(_ones_free)
        _ones_alive[0] -= len(_ones_free)
        _free.clear()
        _ones_free.clear()
        _pooled_bytes = 0
    for _, a in blocks:
        _lib.load().arcte_cuda_host_free(C.c_void_p(a))
