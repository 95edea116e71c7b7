"""Development helper: batched tensor-core scan timing (BASELINE configs[2] shape).
Usage: python tools/bench_batch.py [rows] [nq] [k] [storage]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
from cqs_b200.capi import lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
storage = sys.argv[4] if len(sys.argv) > 4 else "bf16"
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage=storage)
ix.reserve(n)
g = torch.Generator(device=dev); g.manual_seed(1)
for b in range(0, n, 100_000):
    m = min(100_000, n - b)
    x = torch.rand((m, 768), device=dev, generator=g) * 2 - 1
    x /= x.norm(dim=1, keepdim=True)
    ix.append_device(x.data_ptr(), m)
ix.finalize()
ix.set_timing(True)
del x; torch.cuda.empty_cache()
rng = np.random.default_rng(0)
q = rng.standard_normal((nq, 768)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
lib.cqs_b200_debug_batch_reruns.restype = C.c_uint32
lib.cqs_b200_debug_batch_reruns.argtypes = [C.c_void_p]
lib.cqs_b200_debug_last_batch_ms.restype = C.c_float
lib.cqs_b200_debug_last_batch_ms.argtypes = [C.c_void_p]
for it in range(4):
    t0 = time.perf_counter()
    r, s, nn = ix.search_batch_rows(q, k)
    dt = time.perf_counter() - t0
    kms = lib.cqs_b200_debug_last_batch_ms(ix._h)
    flops = 2.0 * nq * n * 768
    print(f"iter {it}: e2e {dt*1e3:8.2f} ms ({nq/dt:9.0f} q/s)  device {kms:8.2f} ms ({nq/kms*1e3:9.0f} q/s, "
          f"{flops/kms/1e9:7.1f} TFLOP/s)  reruns so far {lib.cqs_b200_debug_batch_reruns(ix._h)}")
fl = np.zeros(nq, np.uint32)
lib.cqs_b200_debug_batch_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
lib.cqs_b200_debug_batch_flags(ix._h, fl.ctypes.data_as(C.c_void_p), nq)
lib.cqs_b200_debug_max_row_norm.restype = C.c_float
lib.cqs_b200_debug_max_row_norm.argtypes = [C.c_void_p]
print("flag histogram:", np.bincount(fl, minlength=4).tolist(), "max_row_norm", lib.cqs_b200_debug_max_row_norm(ix._h))
# spot-check 4 queries against the single-query path
for i in (0, 1, nq // 2, nq - 1):
    a, b = ix.search_rows(q[i], k)
    assert np.array_equal(r[i, :nn[i]], a), i
print("spot check vs single-query path: identical")
