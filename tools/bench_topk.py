"""Development helper: microbenchmark of the CTA-level top-k building blocks (csrc/debug_topk.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from cqs_b200.capi import lib
f = lib.cqs_b200_debug_topk_ns
f.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
names = {0: "select", 1: "compact", 2: "topk_finish", 3: "prune", 4: "compact/bitonic"}
for op, cases in ((1, [(64, 64), (128, 128), (173, 100), (256, 256), (500, 500), (512, 500), (900, 500), (1024, 1024), (2048, 1024), (4096, 1024)]),
                  (0, [(600, 500), (900, 500), (1024, 100), (1024, 500), (2048, 500), (2048, 1024), (4096, 500), (4096, 1024)]),
                  (2, [(173, 100), (600, 500), (900, 500), (1100, 500), (1600, 1024), (2500, 500), (4096, 1024)]),
                  (4, [(64, 64), (128, 128), (173, 100), (256, 256), (300, 300), (500, 500), (512, 500), (900, 500)]),
                  (3, [(1024, 0), (4096, 0)])):
    for n, k in cases:
        ns, cnt = C.c_uint64(0), C.c_uint32(0)
        rc = f(0, op, n, k, 20, C.byref(ns), C.byref(cnt))
        cold = C.c_uint64(0)
        f(0, op | 16, n, k, 2, C.byref(cold), C.byref(cnt))   # fresh kernel, first repetition: cold instruction cache
        print(f"{names[op]:12s} n={n:5d} k={k:5d}: {ns.value / 1e3:7.2f} us warm, {cold.value / 1e3:7.2f} us cold  (cnt after {cnt.value}) rc={rc}")
