"""Development helper: thousands of single-query scans (every k range, f32 and bf16+f32 storage), sparse and hybrid
searches in one process — the endurance counterpart of tools/stress_batch.py for the kernels whose tail uses the
chunk sort of csrc/common.cuh."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
import bench as B
n = 1_000_000
per = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
only_sparse = len(sys.argv) > 2 and sys.argv[2] == "sparse"   # skip the dense k sweep, f32 storage only
dev = torch.device("cuda", 0)
rng = np.random.default_rng(3)
q = B.make_queries(min(per, 4096), 77)
t0 = time.time()
for storage in (("f32",) if only_sparse else ("f32", "bf16+f32")):
    ix = cqs_b200.B200Index(768, storage=storage)
    ix.reserve(n)
    for b in range(n // B.BLK):
        x = B.gen_block(torch, dev, b, "uniform")
        ix.append_device(x.data_ptr(), x.shape[0])
    ix.finalize()
    for k in (() if only_sparse else (20, 100, 500, 1024)):
        for i in range(per):
            ix.search_rows(q[i % q.shape[0]], k)
        print(f"{storage} k={k}: {per} searches ok ({time.time() - t0:.0f}s)", flush=True)
    if storage == "f32":
        d_indptr, d_tok, d_w, cdf_h = B.gen_sparse_device(torch, dev, n)
        ix.sparse_attach_device(d_indptr.data_ptr(), d_tok.data_ptr(), d_w.data_ptr(), int(d_tok.shape[0]), B.VOCAB)
        sq = B.sparse_queries(rng, cdf_h, 256, 64)
        for k in (20, 500):
            for i in range(per // 2):
                ix.search_sparse_rows(sq[i % 256][0], sq[i % 256][1], k)
            print(f"sparse k={k}: {per // 2} searches ok ({time.time() - t0:.0f}s)", flush=True)
            for i in range(per // 2):
                ix.search_hybrid_rows(q[i % q.shape[0]], sq[i % 256][0], sq[i % 256][1], 0.8, k)
            print(f"hybrid pool={k}: {per // 2} searches ok ({time.time() - t0:.0f}s)", flush=True)
    del ix
print("done")
