"""Development helper: can a small CTA run on an SM beside a resident scan CTA?  Launches one dense scan
(1M x 768 f32, k = 500) asynchronously and, on a second stream, 148 probe CTAs of T threads; prints when the
probes started relative to the scan (from the scan's own %globaltimer trace).  Findings on B200
(profiles/r02_probe_coresidency.log): with the scan's shared-memory carveout at its maximum (CQS_B200_MAX_CARVEOUT,
set here; the default carveout leaves no room for a CTA that uses shared memory) ONE CTA of at most 3 warps per SM
starts beside the scan — and only from the first kernel launched after it: a second kernel, on the same stream
or not, is not placed until the scan's CTAs retire."""
import os, sys
os.environ["CQS_B200_TRACE"] = "1"
os.environ.setdefault("CQS_B200_MAX_CARVEOUT", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
import bench as B
from cqs_b200.capi import lib, check
n = 1_000_000
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage="f32")
ix.reserve(n)
for b in range(n // B.BLK):
    x = B.gen_block(torch, dev, b, "uniform")
    ix.append_device(x.data_ptr(), x.shape[0])
ix.finalize()
q = torch.from_numpy(B.make_queries(4, 3)).to(dev)
k = 500
o_s = torch.empty((k,), dtype=torch.float32, device=dev); o_r = torch.empty((k,), dtype=torch.int64, device=dev)
o_n = torch.empty((1,), dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
lib.cqs_b200_debug_probe.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
lib.cqs_b200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
for threads, smem, grid in ((96, 0, 148), (96, 24 * 1024, 148), (96, 24 * 1024, 296), (96, 40 * 1024, 148), (128, 0, 148)):
    stamps = torch.zeros((grid,), dtype=torch.int64, device=dev); smid = torch.zeros((grid,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        check(lib.cqs_b200_search_device(ix._h, C.c_void_p(q.data_ptr()), k, None, C.c_void_p(o_s.data_ptr()),
                                         C.c_void_p(o_r.data_ptr()), C.c_void_p(o_n.data_ptr()), C.c_void_p(s1.cuda_stream)))
        rc = lib.cqs_b200_debug_probe(threads, grid, C.c_void_p(stamps.data_ptr()), C.c_void_p(smid.data_ptr()), 20000,
                                      C.c_void_p(s2.cuda_stream), smem)
        assert rc == 0
        torch.cuda.synchronize()
    tr = np.zeros(1024 * 8, np.uint64)
    lib.cqs_b200_debug_trace(ix._h, tr.ctypes.data_as(C.c_void_p), 1024 * 8)
    tr = tr.reshape(1024, 8)[:148].astype(np.int64)
    t0, t_stream_end, t_end = tr[:, 0].min(), tr[:, 1].max(), tr[:, 4].max()
    ps = stamps.cpu().numpy() - t0
    print(f"{grid} probe CTAs of {threads} threads, {smem // 1024} KB shared each: scan streams until {(t_stream_end - t0) / 1e3:.0f} us, ends {(t_end - t0) / 1e3:.0f} us; "
          f"probes start min {ps.min() / 1e3:.1f} / median {np.median(ps) / 1e3:.1f} / max {ps.max() / 1e3:.1f} us; "
          f"{int((ps < (t_stream_end - t0) * 0.5).sum())} of {grid} started in the first half of the scan")

# Second question: does a kernel that FOLLOWS another one on the probe's stream still start beside the scan?
for smem in (0, 24 * 1024):
    grid = 148
    st_a = torch.zeros((grid,), dtype=torch.int64, device=dev); st_b = torch.zeros((grid,), dtype=torch.int64, device=dev)
    smid = torch.zeros((grid,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        check(lib.cqs_b200_search_device(ix._h, C.c_void_p(q.data_ptr()), k, None, C.c_void_p(o_s.data_ptr()),
                                         C.c_void_p(o_r.data_ptr()), C.c_void_p(o_n.data_ptr()), C.c_void_p(s1.cuda_stream)))
        for st in (st_a, st_b):
            assert lib.cqs_b200_debug_probe(96, grid, C.c_void_p(st.data_ptr()), C.c_void_p(smid.data_ptr()), 20000,
                                            C.c_void_p(s2.cuda_stream), smem) == 0
        torch.cuda.synchronize()
    tr = np.zeros(1024 * 8, np.uint64)
    lib.cqs_b200_debug_trace(ix._h, tr.ctypes.data_as(C.c_void_p), 1024 * 8)
    t0 = tr.reshape(1024, 8)[:148].astype(np.int64)[:, 0].min()
    a, b = (st_a.cpu().numpy() - t0) / 1e3, (st_b.cpu().numpy() - t0) / 1e3
    print(f"two 96-thread probe kernels back to back on one stream, {smem // 1024} KB shared: first starts {a.min():.1f}..{a.max():.1f} us, "
          f"second starts {b.min():.1f}..{b.max():.1f} us after the scan's start")
