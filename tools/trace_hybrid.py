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
    d = tr[:148]
    t0 = d[:, 0].min()
    us = lambda x: (x - t0) / 1e3
    live = lambda rows, slot: rows[rows[:, slot] > t0 - 2000000]
    bd = live(tr[812:960], 0)                               # bounds pass (first 148 CTAs), before the scan
    reg = live(tr[662:662 + 296], 6)                        # accumulate CTAs (slots 6, 7)
    sel = live(tr[512:512 + 296], 3)                        # select CTAs (slots 0..5)
    btxt = f"{us(bd[:,0].min()):.1f}..{us(bd[:,1].max()):.1f}" if bd.shape[0] else "n/a"
    print(f"query {i}: bounds {btxt} | dense CTAs start {us(d[:,0].min()):.1f}..{us(d[:,0].max()):.1f}, "
          f"stream end {us(d[:,1].max()):.1f}, ticket {us(d[:,3].max()):.1f}, kernel end {us(d[:,4].max()):.1f}")
    print(f"   accumulate ({reg.shape[0]} CTAs): start {us(reg[:,6].min()):.1f}..{us(reg[:,6].max()):.1f}, end {us(reg[:,7].min()):.1f}..{us(reg[:,7].max()):.1f} | "
          f"select ({sel.shape[0]} CTAs) start {us(sel[:,0].min()):.1f}..{us(sel[:,0].max()):.1f}, loop end {us(sel[:,3].max()):.1f}, merge end {us(sel[:,5].max()):.1f} us")
