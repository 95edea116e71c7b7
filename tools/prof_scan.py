"""Development helper for ncu captures: a few single-query scans of one shape.
Usage: python tools/prof_scan.py [storage] [k] [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
storage = sys.argv[1] if len(sys.argv) > 1 else "bf16"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage=storage)
ix.reserve(n)
for b in range(0, n, 100_000):
    x = torch.rand((100_000, 768), device=dev) * 2 - 1
    x /= x.norm(dim=1, keepdim=True)
    ix.append_device(x.data_ptr(), 100_000)
ix.finalize()
ix.set_timing(True)
q = np.random.default_rng(0).standard_normal(768).astype(np.float32); q /= np.linalg.norm(q)
for i in range(6):
    ix.search_rows(q, k)
    print(f"{storage} k={k}: {ix.last_kernel_ms()*1e3:.1f} us")
