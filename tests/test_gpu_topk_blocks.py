"""The CTA-level sort behind every top-k list (csrc/common.cuh: chunk sort + cross rank, and the bitonic
network above 8 keys per thread) against numpy, at every chunk boundary, for both group sizes the kernels
use (256 consumer threads in the scan, 512 in the sparse / batch kernels)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keys(rng, n, ties):
    # keys shaped like candidates: ordered score word << 32 | ~row, unique rows
    score = rng.standard_normal(n).astype(np.float32) * 0.1
    if ties:
        score = np.round(score, 1)                      # many equal scores: order decided by the row half
    bits = score.view(np.uint32).astype(np.uint64)
    ordered = np.where(bits & 0x80000000, ~bits & 0xFFFFFFFF, bits | 0x80000000).astype(np.uint64)
    rows = rng.permutation(1 << 20)[:n].astype(np.uint64)
    return (ordered << np.uint64(32)) | (~rows & np.uint64(0xFFFFFFFF))


@pytest.mark.parametrize("threads", [256, 512])
@pytest.mark.parametrize("op", [1, 2, 4])   # compact, topk_finish (select + compact), forced bitonic network
def test_cta_sort_matches_numpy(threads, op):
    from cqs_b200.capi import lib
    f = lib.cqs_b200_debug_sort
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(threads + op)
    sizes = [1, 2, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 300, 500, 511, 512, 513, 777, 1000, 1023, 1024, 1025,
             1026, 1040, 1056, 1057, 1500, 2047, 2048, 2049, 2050, 2080, 2081, 3000, 4095, 4096]
    for n in sizes:
        for k in sorted({1, min(n, 20), min(n, 500), max(1, n // 2), n, min(4096, n + 7)}):
            if op == 2 and k > 1024:
                continue                                # the kernels never finish with k above max_k
            keys = _keys(rng, n, ties=(n % 2 == 1))
            out = np.zeros(4096, np.uint64)
            cnt = C.c_uint32(0)
            rc = f(0, threads, op, keys.ctypes.data_as(C.c_void_p), n, k, out.ctypes.data_as(C.c_void_p), C.byref(cnt))
            assert rc == 0, (threads, op, n, k)
            want = np.sort(keys)[::-1][:min(n, k)]
            assert cnt.value == want.shape[0], (threads, op, n, k, cnt.value)
            assert np.array_equal(out[:cnt.value], want), (threads, op, n, k)
