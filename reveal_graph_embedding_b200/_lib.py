"""ctypes binding of libarcte_cuda.so (C ABI: include/arcte_cuda.h).

This is the only place the shared library is loaded.  It fails loudly: a missing
library, a missing symbol or a non-zero status raises -- nothing here or above
falls back to a CPU implementation.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARCTE_CUDA_LIB") or os.path.join(_HERE, "libarcte_cuda.so")  # override: kernel experiments

RULE_ABSORBING, RULE_PAGERANK, RULE_LAZY = 0, 1, 2
SCHEDULE_FIFO, SCHEDULE_FRONTIER = 0, 1
ENGINE_AUTO, ENGINE_FIFO_DENSE, ENGINE_BATCHED_DENSE, ENGINE_BATCHED_HASH, ENGINE_FIFO_COMPACT = -1, 0, 1, 2, 3

# every symbol include/arcte_cuda.h declares (tests/test_abi.py checks the two lists agree)
SYMBOLS = [
    "arcte_cuda_device_count", "arcte_cuda_create", "arcte_cuda_destroy", "arcte_cuda_last_error", "arcte_cuda_configure", "arcte_cuda_set_schedule", "arcte_cuda_set_engine",
    "arcte_cuda_set_graph", "arcte_cuda_set_transition", "arcte_cuda_set_seeds", "arcte_cuda_build_transition", "arcte_cuda_get_transition",
    "arcte_cuda_get_seed_count", "arcte_cuda_get_seeds", "arcte_cuda_epsilon_effective",
    "arcte_cuda_push", "arcte_cuda_extract", "arcte_cuda_centrality", "arcte_cuda_get_segments",
    "arcte_cuda_segments_device", "arcte_cuda_export_segments", "arcte_cuda_assemble", "arcte_cuda_assemble_rows", "arcte_cuda_features_device", "arcte_cuda_get_features", "arcte_cuda_fetch_features", "arcte_cuda_fetch_features_to", "arcte_cuda_host_write_to", "arcte_cuda_host_advise_huge", "arcte_cuda_host_ones_alloc", "arcte_cuda_host_ones_free",
    "arcte_cuda_comm_unique_id", "arcte_cuda_comm_init", "arcte_cuda_comm_init_all", "arcte_cuda_comm_info", "arcte_cuda_exchange_assemble",
    "arcte_cuda_normalize_columns", "arcte_cuda_normalize_features", "arcte_cuda_chi2_contingency", "arcte_cuda_peak_snr",
    "arcte_cuda_chi2_psnr_weights", "arcte_cuda_community_weighting",
    "arcte_cuda_store_features", "arcte_cuda_store_assembled", "arcte_cuda_weighted_fold", "arcte_cuda_get_fold",
    "arcte_cuda_io_read_edge_list", "arcte_cuda_io_edge_list_copy", "arcte_cuda_io_edge_list_free", "arcte_cuda_io_write_features",
    "arcte_cuda_host_alloc", "arcte_cuda_host_free", "arcte_cuda_host_fill_f64", "arcte_cuda_timer_start", "arcte_cuda_timer_stop", "arcte_cuda_flush_l2", "arcte_cuda_features_hash", "arcte_cuda_get_stats",
]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_seeds_total", "n_seeds_shard", "pushes", "edge_touches", "enqueues", "max_queue",
        "support", "touched", "seed_degree", "members", "emitted", "retries", "n_slots",
        "launches", "rounds")] + [(k, C.c_double) for k in (
            "ms_transition", "ms_seeds", "ms_push", "ms_assemble", "alg_bytes_push", "slot_utilisation",
            "ms_exchange")] + [("engine", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ArcteCudaError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load():
    """Load the library once; raise if it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ArcteCudaError(
                "libarcte_cuda.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C reveal_graph_embedding_b200/csrc`. There is no CPU fallback."
                % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for s in SYMBOLS:
            getattr(L, s)  # AttributeError if the ABI drifted
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.arcte_cuda_last_error.restype = C.c_char_p
        L.arcte_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
        L.arcte_cuda_create.argtypes = [C.POINTER(vp), i32]
        L.arcte_cuda_destroy.argtypes = [vp]
        L.arcte_cuda_destroy.restype = None
        L.arcte_cuda_configure.argtypes = [vp, i32, i64, i32, i64]
        L.arcte_cuda_set_schedule.argtypes = [vp, i32, i32, i32, i32, i32, i32]
        L.arcte_cuda_set_engine.argtypes = [vp, i32, i64]
        L.arcte_cuda_set_graph.argtypes = [vp, i64, i64, vp, vp, vp]
        L.arcte_cuda_set_transition.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp]
        L.arcte_cuda_set_seeds.argtypes = [vp, i64, vp]
        L.arcte_cuda_build_transition.argtypes = [vp]
        L.arcte_cuda_get_transition.argtypes = [vp, vp, vp, vp]
        L.arcte_cuda_get_seed_count.argtypes = [vp, C.POINTER(i64)]
        L.arcte_cuda_get_seeds.argtypes = [vp, vp]
        L.arcte_cuda_epsilon_effective.argtypes = [vp, dbl, i64, vp, vp]
        L.arcte_cuda_push.argtypes = [vp, i32, i64, dbl, dbl, vp, vp, C.POINTER(i64)]
        L.arcte_cuda_extract.argtypes = [vp, i32, dbl, dbl, i32, i32, vp, C.POINTER(i64), C.POINTER(i64)]
        L.arcte_cuda_centrality.argtypes = [vp, dbl, dbl, vp]
        L.arcte_cuda_get_segments.argtypes = [vp, vp, vp, vp, vp]
        L.arcte_cuda_segments_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.arcte_cuda_export_segments.argtypes = [vp, vp, vp, vp, vp]
        L.arcte_cuda_assemble.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
        L.arcte_cuda_assemble_rows.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i64, i64, C.POINTER(i64)]
        L.arcte_cuda_features_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64),
                                                 C.POINTER(i64)]
        L.arcte_cuda_get_features.argtypes = [vp, vp, vp, vp]
        L.arcte_cuda_fetch_features.argtypes = [vp, vp, vp, vp, i32, i32]
        L.arcte_cuda_fetch_features_to.argtypes = [vp, i64, vp, vp, vp, i32, i32]
        L.arcte_cuda_host_write_to.argtypes = [i64, vp, vp, i64]
        L.arcte_cuda_host_advise_huge.argtypes = [vp, i64]
        L.arcte_cuda_host_ones_alloc.argtypes = [i64, C.POINTER(vp), C.POINTER(i64)]
        L.arcte_cuda_host_ones_free.argtypes = [vp, i64]
        L.arcte_cuda_comm_unique_id.argtypes = [vp]
        L.arcte_cuda_comm_init.argtypes = [vp, i32, i32, vp]
        L.arcte_cuda_comm_init_all.argtypes = [vp, i32]
        L.arcte_cuda_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.arcte_cuda_exchange_assemble.argtypes = [vp, C.POINTER(i64)]
        L.arcte_cuda_normalize_columns.argtypes = [vp, i64, i64, vp, vp, vp, vp]
        L.arcte_cuda_normalize_features.argtypes = [vp]
        L.arcte_cuda_chi2_contingency.argtypes = [vp, i64, i64, vp, vp, i64, vp, vp, vp, vp]
        L.arcte_cuda_peak_snr.argtypes = [vp, i64, i64, vp, vp]
        L.arcte_cuda_chi2_psnr_weights.argtypes = [vp, i64, i64, vp, vp, i64, vp, vp, vp, vp]
        L.arcte_cuda_community_weighting.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
        L.arcte_cuda_store_features.argtypes = [vp, i64, i64, vp, vp, vp]
        L.arcte_cuda_store_assembled.argtypes = [vp]
        L.arcte_cuda_weighted_fold.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]
        L.arcte_cuda_get_fold.argtypes = [vp, i32, vp, vp, vp]
        L.arcte_cuda_io_read_edge_list.argtypes = [C.c_char_p, C.c_char_p, i32, i32, C.POINTER(vp), C.POINTER(i64),
                                                   C.POINTER(i64)]
        L.arcte_cuda_io_edge_list_copy.argtypes = [vp, vp, vp, vp, vp]
        L.arcte_cuda_io_edge_list_free.argtypes = [vp]
        L.arcte_cuda_io_edge_list_free.restype = None
        L.arcte_cuda_io_write_features.argtypes = [C.c_char_p, C.c_char_p, i64, vp, vp, vp, vp, i32, C.POINTER(i64)]
        L.arcte_cuda_host_alloc.argtypes = [C.POINTER(vp), i64]
        L.arcte_cuda_host_free.argtypes = [vp]
        L.arcte_cuda_host_fill_f64.argtypes = [vp, i64, dbl, i32]
        L.arcte_cuda_timer_start.argtypes = [vp]
        L.arcte_cuda_timer_stop.argtypes = [vp, C.POINTER(dbl)]
        L.arcte_cuda_flush_l2.argtypes = [vp]
        L.arcte_cuda_features_hash.argtypes = [vp, i64, i64, C.POINTER(C.c_uint64)]
        L.arcte_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
        for s in SYMBOLS:
            if s not in ("arcte_cuda_last_error", "arcte_cuda_destroy", "arcte_cuda_io_edge_list_free"):
                getattr(L, s).restype = C.c_int
        _lib = L
        return _lib


def check(rc):
    if rc != 0:
        msg = load().arcte_cuda_last_error()
        raise ArcteCudaError("libarcte_cuda status %d: %s" % (rc, (msg or b"").decode("utf-8", "replace")))


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)
