"""Development helper: do the two legs of a hybrid query overlap?  %globaltimer stamps of the dense scan
(rows 0..147 of the trace buffer) and of the sparse search kernel (rows 512..807) for one query.
Usage: CQS_B200_TRACE is set here; python tools/trace_hybrid.py"""
import os, sys
os.environ["CQS_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
import bench as B
from cqs_b200.capi import lib
n = 1_000_000
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage="f32")
ix.reserve(n)
for b in range(n // B.BLK):
    x = B.gen_block(torch, dev, b, "uniform")
    ix.append_device(x.data_ptr(), x.shape[0])
ix.finalize()
d_indptr, d_tok, d_w, cdf_h = B.gen_sparse_device(torch, dev, n)
ix.sparse_attach_device(d_indptr.data_ptr(), d_tok.data_ptr(), d_w.data_ptr(), int(d_tok.shape[0]), B.VOCAB)
rng = np.random.default_rng(17)
dq = B.make_queries(8, 23)
sq = B.sparse_queries(rng, cdf_h, 8, 64)
lib.cqs_b200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
for i in range(8):
    ix.search_hybrid_rows(dq[i], sq[i][0], sq[i][1], 0.8, 500)
    tr = np.zeros(1024 * 8, np.uint64)
    lib.cqs_b200_debug_trace(ix._h, tr.ctypes.data_as(C.c_void_p), 1024 * 8)
    tr = tr.reshape(1024, 8).astype(np.int64)
    d, s = tr[:148], tr[512:512 + 296]
    s = s[s[:, 3] > d[:, 0].min()]                         # sparse CTAs of THIS query only (the grid can be < 296)
    t0 = d[:, 0].min()
    print(f"query {i}: dense CTAs start {(d[:,0].min()-t0)/1e3:.1f}..{(d[:,0].max()-t0)/1e3:.1f}, stream end {(d[:,1].max()-t0)/1e3:.1f}, "
          f"ticket {(d[:,3].max()-t0)/1e3:.1f}, kernel end {(d[:,4].max()-t0)/1e3:.1f} | sparse search CTAs start "
          f"{(s[:,0].min()-t0)/1e3:.1f}..{(s[:,0].max()-t0)/1e3:.1f}, loop end {(s[:,3].max()-t0)/1e3:.1f}, merge end {(s[:,5].max()-t0)/1e3:.1f} us")
