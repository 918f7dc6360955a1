"""Host arrays for results.

Round 1 kept pools of page-locked result buffers here (gigabytes of unswappable memory that grew with
every result a caller kept).  They are gone: results are ordinary memory.  The library streams device memory
into plain numpy arrays through a small fixed ring of pinned slots on several host threads (csrc/hostcopy.cu),
which reaches PCIe rate on the first call of a process as well and pins 128 MB in total; and the value array of
a feature matrix -- all ones but for a few patched entries -- is `ones()`: copy-on-write mappings of one 32 MB
in-memory file of ones, so it is never written at all.
"""
import ctypes as C

import numpy as np

from . import _lib


def empty(count, dtype):
    """An uninitialised 1-D array; its pages are first touched by the library's copy threads."""
    return np.empty(int(count), dtype=dtype)


class _OnesRegion:
    """Owner of one copy-on-write mapping (unmapped when the last array viewing it dies)."""

    def __init__(self, count):
        addr, nbytes = C.c_void_p(), C.c_int64(0)
        _lib.check(_lib.load().arcte_cuda_host_ones_alloc(int(count), C.byref(addr), C.byref(nbytes)))
        self._addr, self._nbytes = addr.value, nbytes.value
        self.__array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (self._addr or 0, False), "version": 3}

    def __del__(self):
        try:
            if self._addr:
                _lib.load().arcte_cuda_host_ones_free(C.c_void_p(self._addr), self._nbytes)
                self._addr = None
        except Exception:
            pass


def ones(count):
    """float64[count] of 1.0, writable, created without touching count * 8 bytes (see the module docstring).
    Small arrays are plain numpy."""
    count = int(count)
    if count < (1 << 20):
        return np.ones(count, dtype=np.float64)
    return np.asarray(_OnesRegion(count))
