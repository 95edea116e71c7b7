"""Development helper for ncu captures of the hybrid path (BASELINE configs[4] shape): builds the 1M x 768 f32
corpus + SPLADE postings exactly like bench.py's hybrid_1M record and runs a few hybrid queries.
Usage: python tools/prof_hybrid.py [docs] [queries]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
import bench as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
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
dq = B.make_queries(nq, 23)
sq = B.sparse_queries(rng, cdf_h, nq, 64)
ix.set_timing(True)
lat = []
for i in range(nq):
    t0 = time.perf_counter()
    ix.search_hybrid_rows(dq[i], sq[i][0], sq[i][1], B.ALPHAS[i % 9], 500)
    lat.append(time.perf_counter() - t0)
print("hybrid e2e ms:", [round(x * 1e3, 3) for x in lat])
