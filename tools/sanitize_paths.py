"""Development helper: one small pass over every kernel family (single-query scans at several k, the bf16
shadow scan with re-scoring, the tensor-core batch, the sparse leg, the hybrid call), sized so that
a whole pass takes seconds (a quick "does every path still run" check after a change to shared code)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
import bench as B
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 250_000) // B.BLK * B.BLK
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
for storage in ("f32", "bf16", "bf16+f32"):
    ix = cqs_b200.B200Index(768, storage=storage)
    ix.reserve(n)
    for b in range(n // B.BLK):
        x = B.gen_block(torch, dev, b, "clustered" if storage == "bf16+f32" else "uniform")
        ix.append_device(x.data_ptr(), x.shape[0])
    ix.finalize()
    q = B.make_queries(300, 11)
    for k in (1, 20, 100, 500, 1024):
        ix.search_rows(q[0], k)
    if storage != "f32":
        for nq in (16, 300):
            ix.search_batch_rows(q[:nq], 100)
            ix.search_batch_rows(q[:nq], 20)
    if storage == "f32":
        d_indptr, d_tok, d_w, cdf_h = B.gen_sparse_device(torch, dev, n)
        ix.sparse_attach_device(d_indptr.data_ptr(), d_tok.data_ptr(), d_w.data_ptr(), int(d_tok.shape[0]), B.VOCAB)
        sq = B.sparse_queries(rng, cdf_h, 4, 64)
        for i in range(4):
            ix.search_sparse_rows(sq[i][0], sq[i][1], 500 if i & 1 else 20)
            ix.search_hybrid_rows(q[i], sq[i][0], sq[i][1], 0.8, 500 if i & 1 else 100)
    print(storage, "ok", flush=True)
    del ix
print("done")
