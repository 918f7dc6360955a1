"""Host arrays for results.

Round 1 kept pools of page-locked result buffers here (gigabytes of unswappable memory that grew with
every result a caller kept).  They are gone: results are ordinary numpy arrays, and the library streams
device memory into them through a small fixed ring of pinned slots on several host threads
(csrc/hostcopy.cu), which reaches PCIe rate on the first call of a process as well and pins 64 MB in total.
"""
import numpy as np


def empty(count, dtype):
    """An uninitialised 1-D array; its pages are first touched by the library's copy threads."""
    return np.empty(int(count), dtype=dtype)
