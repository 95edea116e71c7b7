"""Development helper: many tensor-core batch searches on one bf16 index with fresh queries every call
(a flaky failure dies with its process: run it several times and compare the iteration counts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
import bench as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage="bf16")
ix.reserve(n)
for b in range(n // B.BLK):
    x = B.gen_block(torch, dev, b, "uniform")
    ix.append_device(x.data_ptr(), x.shape[0])
ix.finalize()
t0 = time.time()
for it in range(iters):
    q = B.make_queries(1024, 1000 + it)
    try:
        ix.search_batch_rows(q, k)
    except Exception as e:
        print(f"FAILED at call {it}: {str(e)[-90:]}", flush=True)
        sys.exit(1)
print(f"{iters} calls ok in {time.time() - t0:.1f}s", flush=True)
