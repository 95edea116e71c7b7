"""Development helper: millions of CTA-level sorts (csrc/debug_topk.cu, debug_sort_stress_kernel), order checked on
the device — a reproducer for rare failures of the sort code itself.  Each case runs in its own process (a device
fault kills the context)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import ctypes as C
    from cqs_b200.capi import lib
    threads, op, n, k, grid, reps = (int(x) for x in sys.argv[1:7])
    f = lib.cqs_b200_debug_sort_stress
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    err = C.c_uint64(0)
    rc = f(0, threads, op, n, k, grid, reps, C.byref(err))
    print(f"threads {threads} op {op} n {n} k {k}: {grid * reps} sorts, rc {rc}, order errors {err.value}", flush=True)
    sys.exit(0 if rc == 0 and err.value == 0 else 1)
cases = [(256, 1, n, k) for n, k in ((100, 100), (256, 64), (500, 500), (1000, 500), (1500, 128), (2048, 1024))]
cases += [(256, 2, 3000, 128), (256, 0, 4096, 128), (512, 1, 1000, 500), (512, 2, 3000, 500)]
for threads, op, n, k in cases:
    subprocess.run([sys.executable, os.path.abspath(__file__), str(threads), str(op), str(n), str(k), "4096", "200"])
