"""Development helper: why does the batch path re-run queries?  Flag histogram of one 1024-query batch
(1 = sub-pool overflow, 2 = too few exact survivors, 3 = margin not proven) and the inputs of the bound.
Usage: python tools/diag_batch_flags.py [rows] [mode uniform|clustered] [storage]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
import bench as B
from cqs_b200.capi import lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "clustered"
storage = sys.argv[3] if len(sys.argv) > 3 else "bf16+f32"
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage=storage)
ix.reserve(n)
for b in range(n // B.BLK):
    x = B.gen_block(torch, dev, b, mode)
    ix.append_device(x.data_ptr(), x.shape[0])
ix.finalize()
q = B.make_queries(1024, 101, mode)
for f in ("cqs_b200_debug_max_row_norm", "cqs_b200_debug_max_row_delta"):
    getattr(lib, f).restype = C.c_float; getattr(lib, f).argtypes = [C.c_void_p]
lib.cqs_b200_debug_batch_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
print("max_row_norm", lib.cqs_b200_debug_max_row_norm(ix._h), "max_row_delta", lib.cqs_b200_debug_max_row_delta(ix._h))
q16 = torch.from_numpy(q).to(torch.bfloat16).to(torch.float32).numpy()
print("dqnorm median", float(np.median(np.linalg.norm(q - q16, axis=1))))
r, s, nn = ix.search_batch_rows(q, 20)
fl = np.zeros(1024, np.uint32)
lib.cqs_b200_debug_batch_flags(ix._h, fl.ctypes.data_as(C.c_void_p), 1024)
print("flag histogram [ok, overflow, few survivors, margin]:", np.bincount(fl, minlength=4).tolist())
# margins: exact score gaps rank 20 -> 64 -> 128 for a few queries (single-query path, k = 128)
gaps = []
for i in range(0, 1024, 32):
    rr, ss = ix.search_rows(q[i], 128)
    gaps.append((float(ss[19] - ss[63]), float(ss[19] - ss[127]), float(ss[0]), float(ss[19])))
g = np.asarray(gaps)
print("exact score gap rank20-rank64: median %.5f min %.5f | rank20-rank128: median %.5f | top1 %.3f rank20 %.3f" %
      (np.median(g[:, 0]), g[:, 0].min(), np.median(g[:, 1]), np.median(g[:, 2]), np.median(g[:, 3])))
