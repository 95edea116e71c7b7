"""Development helper: hybrid (dense + SPLADE + alpha fusion) timing at BASELINE configs[4]
shape: 1M x 768 f32 dense + sparse CSR (vocab 30522, ~200 nnz/doc), pool 500.
Usage: python tools/bench_hybrid.py [docs] [mean_nnz] [q_nnz]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mean_nnz = int(sys.argv[2]) if len(sys.argv) > 2 else 200
q_nnz = int(sys.argv[3]) if len(sys.argv) > 3 else 64
vocab, dim = 30522, 768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
ix = cqs_b200.B200Index(dim, storage="f32")
ix.reserve(n)
for b in range(0, n, 100_000):
    m = min(100_000, n - b)
    x = torch.rand((m, dim), device=dev, generator=g) * 2 - 1
    x /= x.norm(dim=1, keepdim=True)
    ix.append_device(x.data_ptr(), m)
ix.finalize()
ix.set_timing(True)
# sparse: Zipf(1.1) token ids, duplicates inside a doc removed, ascending; weights log1p(relu(N(.8,.5))) > .01
t0 = time.time()
p = 1.0 / torch.arange(1, vocab + 1, device=dev, dtype=torch.float64) ** 1.1
cdf = torch.cumsum(p / p.sum(), 0).float()
W = int(mean_nnz * 1.6)
indptr = [0]; toks = []; ws = []
for b in range(0, n, 50_000):
    m = min(50_000, n - b)
    want = torch.clamp(torch.poisson(torch.full((m,), float(mean_nnz), device=dev), generator=g), 20, min(400, W)).long()
    t = torch.searchsorted(cdf, torch.rand((m, W), device=dev, generator=g)).clamp_(0, vocab - 1)
    t, _ = torch.sort(t, dim=1)
    dup = torch.zeros_like(t, dtype=torch.bool); dup[:, 1:] = t[:, 1:] == t[:, :-1]
    rank = torch.cumsum((~dup).long(), 1)                      # 1-based rank among the unique ones
    w = torch.log1p(torch.relu(torch.randn((m, W), device=dev, generator=g) * 0.5 + 0.8))
    keep = (~dup) & (rank <= want[:, None]) & (w > 0.01)
    cnt = keep.sum(1)
    toks.append(t[keep].to(torch.int32).cpu().numpy().astype(np.uint32)); ws.append(w[keep].cpu().numpy())
    ip = np.cumsum(cnt.cpu().numpy()) + indptr[-1]
    indptr.extend(ip.tolist())
indptr = np.asarray(indptr, np.uint64); tok = np.concatenate(toks); w = np.concatenate(ws)
print(f"sparse corpus: {tok.shape[0]/1e6:.1f}M postings, {tok.shape[0]/n:.0f}/doc, built in {time.time()-t0:.1f}s")
t0 = time.time(); ix.sparse_attach(indptr, tok, w, vocab); print(f"attach (H2D of the CSR + inverted-index build on the device): {time.time()-t0:.2f}s")
d_ip = torch.from_numpy(indptr.view(np.int64)).to(dev); d_tok = torch.from_numpy(tok.view(np.int32)).to(dev); d_w = torch.from_numpy(w).to(dev)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.time(); ix.sparse_attach_device(d_ip.data_ptr(), d_tok.data_ptr(), d_w.data_ptr(), tok.shape[0], vocab)
    dt = time.time() - t0
print(f"attach_device (build only, CSR resident): {dt*1e3:.1f} ms = {tok.shape[0]*24/dt/1e9:.0f} GB/s of CSR-in + postings-out bytes")
del d_ip, d_tok, d_w
rng = np.random.default_rng(0)
cdf_h = cdf.cpu().numpy()
def sparse_query():
    t = np.unique(np.searchsorted(cdf_h, rng.random(q_nnz * 2)).clip(0, vocab - 1))[:q_nnz]
    return t.astype(np.uint32), np.log1p(np.maximum(rng.normal(0.8, 0.5, t.shape[0]), 0.02)).astype(np.float32)
alphas = [0.85, 0.6, 1.0, 0.8, 0.1, 0.8, 0.0, 0.7, 0.8]
lat_h, lat_d, lat_s = [], [], []
for i in range(40):
    q = rng.standard_normal(dim).astype(np.float32); q /= np.linalg.norm(q)
    qt, qw = sparse_query()
    t0 = time.perf_counter(); r = ix.search_hybrid_rows(q, qt, qw, alphas[i % 9], 500); lat_h.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); ix.search_rows(q, 500); lat_d.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); ix.search_sparse_rows(qt, qw, 500); lat_s.append(time.perf_counter() - t0); ks = ix.last_kernel_ms()
touched = sum(int(indptr[0] * 0) for _ in range(1))
post = np.diff(np.asarray(np.searchsorted(np.sort(tok), np.arange(vocab + 1))))  # postings per token
qt, qw = sparse_query()
print(f"hybrid pool-500 e2e p50 {np.median(lat_h[5:])*1e3:.3f} ms | dense-500 alone {np.median(lat_d[5:])*1e3:.3f} ms | "
      f"sparse-500 alone {np.median(lat_s[5:])*1e3:.3f} ms (kernel {ks*1e3:.0f} us; query touches ~{post[qt].sum()*8/1e6:.1f} MB of postings)")
print("rows returned:", r["rows"].shape[0], "in both legs:", int((r["present"] == 3).sum()))
# CPU baseline of the sparse leg + fusion on the host cores (bounded sample: the first 300k docs)
from tools.cpu_sparse_baseline import time_cpu_sparse
ns = min(n, 300_000)
e_end = int(indptr[ns])
cpu = time_cpu_sparse(indptr[:ns + 1], tok[:e_end], w[:e_end], vocab, [sparse_query() for _ in range(8)], 500,
                      [(int(a), float(b)) for a, b in zip(*ix.search_rows(q, 500))])
print(f"CPU port, first {ns} docs ({e_end/1e6:.1f}M postings): index build {cpu['build_s']:.1f} s, sparse leg p50 "
      f"{cpu['sparse_ms_p50']:.2f} ms on 1 core, fusion (numpy restatement) {cpu['fuse_ms']:.2f} ms")
