"""Development helper: kernel-only time (CUDA events around the scan launch inside the
library) for a few shapes.  Usage: python tools/kernel_times.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cqs_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
for storage in ("f32", "bf16", "bf16+f32"):
    ix = cqs_b200.B200Index(768, storage=storage)
    ix.reserve(n)
    for b in range(0, n, 100_000):
        m = min(100_000, n - b)
        x = torch.rand((m, 768), device=dev) * 2 - 1
        x /= x.norm(dim=1, keepdim=True)
        ix.append_device(x.data_ptr(), m)
    ix.finalize()
    ix.set_timing(True)
    q = np.random.default_rng(0).standard_normal(768).astype(np.float32); q /= np.linalg.norm(q)
    for k in (1, 20, 100, 500, 1024):
        ts = []
        for i in range(12):
            ix.search_rows(q, k); ts.append(ix.last_kernel_ms())
        ts = sorted(ts[2:])
        byts = n * 768 * (4 if storage == "f32" else 2)   # bf16+f32: the bf16 shadow is what a (proven) query streams
        print(f"{storage} n={n} k={k:5d} kernel median {ts[len(ts)//2]*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us  -> {byts/ts[len(ts)//2]/1e6:8.1f} GB/s")
    ix.close()
