"""Development helper: phase stamps of the sparse kernel (CQS_B200_TRACE=1)."""
import os, sys
os.environ["CQS_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
from cqs_b200.capi import lib
n, vocab, mean_nnz, q_nnz = 1_000_000, 30522, 200, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
ix = cqs_b200.B200Index(8, storage="f32")
ix.append(None, np.ones((n, 8), np.float32)); ix.finalize()
ix.set_timing(True)
p = 1.0 / torch.arange(1, vocab + 1, device=dev, dtype=torch.float64) ** 1.1
cdf = torch.cumsum(p / p.sum(), 0).float()
W = 320; indptr = [0]; toks = []; ws = []
for b in range(0, n, 50_000):
    m = 50_000
    want = torch.clamp(torch.poisson(torch.full((m,), float(mean_nnz), device=dev), generator=g), 20, W).long()
    t = torch.searchsorted(cdf, torch.rand((m, W), device=dev, generator=g)).clamp_(0, vocab - 1)
    t, _ = torch.sort(t, dim=1)
    dup = torch.zeros_like(t, dtype=torch.bool); dup[:, 1:] = t[:, 1:] == t[:, :-1]
    rank = torch.cumsum((~dup).long(), 1)
    w = torch.log1p(torch.relu(torch.randn((m, W), device=dev, generator=g) * 0.5 + 0.8))
    keep = (~dup) & (rank <= want[:, None]) & (w > 0.01)
    toks.append(t[keep].to(torch.int32).cpu().numpy().astype(np.uint32)); ws.append(w[keep].cpu().numpy())
    indptr.extend((np.cumsum(keep.sum(1).cpu().numpy()) + indptr[-1]).tolist())
ix.sparse_attach(np.asarray(indptr, np.uint64), np.concatenate(toks), np.concatenate(ws), vocab)
rng = np.random.default_rng(0); cdf_h = cdf.cpu().numpy()
lib.cqs_b200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
for k in (20, 500):
    for it in range(3):
        t = np.unique(np.searchsorted(cdf_h, rng.random(q_nnz * 2)).clip(0, vocab - 1))[:q_nnz].astype(np.uint32)
        qw = np.log1p(np.maximum(rng.normal(0.8, 0.5, t.shape[0]), 0.02)).astype(np.float32)
        ix.search_sparse_rows(t, qw, k)
    tr = np.zeros(1024 * 8, np.uint64)
    lib.cqs_b200_debug_trace(ix._h, tr.ctypes.data_as(C.c_void_p), 1024 * 8)
    tr = tr.reshape(1024, 8).astype(np.int64)
    acc = tr[662:662 + 296]; acc = acc[acc[:, 6] > 0]     # regular accumulate CTAs (rows 150.. of the sparse trace)
    sel = tr[512:512 + 296]; sel = sel[sel[:, 3] > 0]
    t0 = acc[:, 6].min(); us = lambda x: (x - t0) / 1e3
    last = int(np.argmax(sel[:, 5]))
    print(f"k={k}: all kernels {ix.last_kernel_ms()*1e3:.0f} us | accumulate ({acc.shape[0]} CTAs): start 0..{us(acc[:,6].max()):.1f}, "
          f"end {us(acc[:,7].min()):.1f}..{us(acc[:,7].max()):.1f} | select ({sel.shape[0]} CTAs): start {us(sel[:,0].min()):.1f}, first "
          f"half pushed {np.median(us(sel[:,1])):.1f}, loop end med {np.median(us(sel[:,3])):.1f} max {us(sel[:,3].max()):.1f}, "
          f"finish {np.median(sel[:,4]-sel[:,3])/1e3:.1f}, merge (CTA {last}) {us(sel[last,4]):.1f} -> {us(sel[last,5]):.1f}")
