"""Development helper: per-CTA phase timestamps of the scan kernel (CQS_B200_TRACE=1)."""
import os, sys
os.environ["CQS_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
from cqs_b200.capi import lib
n = 1_000_000
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage=sys.argv[1] if len(sys.argv) > 1 else "f32")
ix.reserve(n)
for b in range(0, n, 100_000):
    x = torch.rand((100_000, 768), device=dev) * 2 - 1
    x /= x.norm(dim=1, keepdim=True)
    ix.append_device(x.data_ptr(), 100_000)
ix.finalize()
ix.set_timing(True)
q = np.random.default_rng(0).standard_normal(768).astype(np.float32); q /= np.linalg.norm(q)
lib.cqs_b200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
for k in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 20, 500)):
    for i in range(4):
        ix.search_rows(q, k)
    trall = np.zeros(1024 * 8, np.uint64)
    lib.cqs_b200_debug_trace(ix._h, trall.ctypes.data_as(C.c_void_p), 1024 * 8)
    ms = trall.reshape(1024, 8)[1023].astype(np.int64)       # extra stamps of the merge (last CTA)
    cs = trall.reshape(1024, 8)[1022].astype(np.int64)       # final sort: wall clock and SM cycles
    tr = trall.reshape(1024, 8)[:148].astype(np.int64)
    t0 = tr[:, 0].min()
    rel = (tr - t0) / 1e3
    last = int(np.argmax(tr[:, 4]))
    print(f"k={k}: kernel_ms={ix.last_kernel_ms()*1e3:.1f}us")
    print(f"  start spread   : {rel[:,0].min():.1f} .. {rel[:,0].max():.1f} us")
    print(f"  scan loop end  : min {rel[:,1].min():.1f} med {np.median(rel[:,1]):.1f} max {rel[:,1].max():.1f} us")
    print(f"  final compact  : med {(rel[:,2]-rel[:,1]).mean():.2f} us")
    print(f"  ticket         : max {rel[:,3].max():.1f} us")
    print(f"  merge detail   : start {(ms[0]-t0)/1e3:.1f}, col bounds done {(ms[1]-t0)/1e3:.1f}, rounds done {(ms[2]-t0)/1e3:.1f}, "
          f"select1 done {(ms[4]-t0)/1e3:.1f}, select2 done {(ms[5]-t0)/1e3:.1f}; cnt at finish {ms[7]}, cnt before final sort {ms[6]}, thr score word {ms[3] >> 32:#x}")
    if cs[1] > cs[0] > 0:
        print(f"  final sort     : {(cs[1]-cs[0])/1e3:.2f} us, {cs[2]} SM cycles ({cs[2]/(cs[1]-cs[0]):.2f} GHz)")
    print(f"  merge (CTA {last}) : {rel[last,3]:.1f} -> {rel[last,4]:.1f} us; loads+push done {rel[last,5]:.1f}, compact done {rel[last,6]:.1f}, cnt before compact {tr[last,7]}")
