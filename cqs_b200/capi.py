"""ctypes binding of the C ABI declared in include/cqs_b200.h.

Loading is strict: if ``libcqs_b200.so`` is missing the import raises — there
is no fallback implementation behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CQS_B200_LIB") or os.path.join(_HERE, "libcqs_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_POISONED, ERR_OOM, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
METRIC_COSINE, METRIC_DOT = 0, 1
STORAGE_F32, STORAGE_BF16, STORAGE_BF16_F32 = 0, 1, 2
MAX_K = 1024

u64p, u32p, f32p, u8p, i32p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                               C.POINTER(C.c_uint8), C.POINTER(C.c_int32))
vp = C.c_void_p

# name -> (restype, argtypes).  Every symbol include/cqs_b200.h declares.
SIGNATURES = {
    "cqs_b200_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]),
    "cqs_b200_reserve": (C.c_int, [vp, C.c_uint64]),
    "cqs_b200_set_row_base": (C.c_int, [vp, C.c_uint64]),
    "cqs_b200_append_rows_f32": (C.c_int, [vp, vp, C.c_uint64]),
    "cqs_b200_append_rows_f32_device": (C.c_int, [vp, vp, C.c_uint64]),
    "cqs_b200_finalize": (C.c_int, [vp]),
    "cqs_b200_reopen": (C.c_int, [vp]),
    "cqs_b200_destroy": (None, [vp]),
    "cqs_b200_save": (C.c_int, [vp, C.c_char_p]),
    "cqs_b200_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
    "cqs_b200_search": (C.c_int, [vp, vp, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_set_row_meta": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "cqs_b200_set_row_signals": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "cqs_b200_search_filtered": (C.c_int, [vp, vp, C.c_uint32, C.c_float, vp, vp, C.c_int, vp, vp, vp]),
    "cqs_b200_search_typed": (C.c_int, [vp, vp, C.c_uint32, vp, vp, vp, vp, vp]),
    "cqs_b200_rrf_fuse": (C.c_int, [C.c_int, vp, vp, C.c_uint32, C.c_float, C.c_uint32, vp, vp, vp]),
    "cqs_b200_search_batch": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_sparse_attach": (C.c_int, [vp, vp, vp, vp, C.c_uint32]),
    "cqs_b200_sparse_attach_device": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_uint32]),
    "cqs_b200_sparse_save": (C.c_int, [vp, C.c_char_p, C.c_uint64]),
    "cqs_b200_sparse_load": (C.c_int, [vp, C.c_char_p, C.c_uint64]),
    "cqs_b200_search_sparse": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_search_hybrid": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_float, C.c_uint32, vp,
                                         vp, vp, vp, vp, vp, vp]),
    "cqs_b200_fuse_pools": (C.c_int, [C.c_int, vp, vp, C.c_uint32, vp, vp, C.c_uint32, C.c_float,
                                      C.c_uint32, vp, vp, vp, vp, vp, vp]),
    "cqs_b200_route_centroids": (C.c_int, [C.c_int, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32,
                                           C.c_float, vp, vp]),
    "cqs_b200_search_device": (C.c_int, [vp, vp, C.c_uint32, vp, vp, vp, vp, vp]),
    "cqs_b200_merge_topk_device": (C.c_int, [C.c_int, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32,
                                             vp, vp, vp, vp]),
    "cqs_b200_peer_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "cqs_b200_peer_handle": (C.c_int, [vp, vp]),
    "cqs_b200_peer_connect": (C.c_int, [vp, vp]),
    "cqs_b200_peer_connect_local": (C.c_int, [C.POINTER(vp), C.c_uint32]),
    "cqs_b200_peer_set_timeout_ms": (C.c_int, [vp, C.c_uint32]),
    "cqs_b200_peer_status": (C.c_int, [vp]),
    "cqs_b200_peer_destroy": (None, [vp]),
    "cqs_b200_search_sharded": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_search_sharded_device": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp]),
    "cqs_b200_search_many_device": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]),
    "cqs_b200_search_batch_sharded": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_search_hybrid_sharded": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, C.c_float, C.c_uint32, vp,
                                                 vp, vp, vp, vp, vp, vp]),
    "cqs_b200_peer_gather_merge": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp]),
    "cqs_b200_len": (C.c_uint64, [vp]),
    "cqs_b200_dim": (C.c_uint32, [vp]),
    "cqs_b200_max_k": (C.c_uint32, [vp]),
    "cqs_b200_is_poisoned": (C.c_int, [vp]),
    "cqs_b200_scores_are_cosine": (C.c_int, [vp]),
    "cqs_b200_name": (C.c_char_p, []),
    "cqs_b200_last_error": (C.c_char_p, []),
    "cqs_b200_kernel_launches": (C.c_uint64, []),
    "cqs_b200_last_kernel_ms": (C.c_float, [vp]),
    "cqs_b200_set_timing": (C.c_int, [vp, C.c_int]),
}


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cqs_b200 error {code}: {msg}")
        self.code = code


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(cqs_b200 has no CPU fallback)")
    dll = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(dll, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return dll


lib = load_library()


def check(rc: int) -> None:
    if rc != OK:
        raise B200Error(rc, (lib.cqs_b200_last_error() or b"").decode("utf-8", "replace"))


def ptr(a):
    """numpy array -> void* (None passes NULL)."""
    return None if a is None else a.ctypes.data_as(vp)
