"""Small instances of every kernel family, meant to be run under compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
from cqs_b200.capi import lib, check
from cqs_b200.sharded import PeerGroup, search_sharded
from oracle import cqs_oracle as O

f32 = np.float32
n, dim = 6000, 768
rows = O.fast_unit_rows(n, dim, seed=1)
qs = O.fast_unit_rows(6, dim, seed=2)
for storage in ("f32", "bf16"):
    ix = cqs_b200.B200Index(dim, storage=storage)
    ix.append(None, rows); ix.finalize()
    for k in (1, 20, 100, 500, 1024):
        a, b = ix.search_rows(qs[0], k)
        assert a.shape[0] == k
    r, s, nn = ix.search_batch_rows(qs, 20)                     # exact pipelined (nq < 8) path
    pg = PeerGroup(0, 1, 0)
    c, d = search_sharded(ix, pg, qs[1], 20)
    assert np.array_equal(c, ix.search_rows(qs[1], 20)[0])
    pg.close()
    if storage == "bf16":
        big = O.fast_unit_rows(16, dim, seed=3)
        r, s, nn = ix.search_batch_rows(big, 20)                # tensor-core path
        assert np.array_equal(r[3], ix.search_rows(big[3], 20)[0])
    ix.close()
print("dense ok")
# emulated 2-rank gather+merge on one device
dev = torch.device("cuda", 0)
G, Q, k = 2, 5, 20
groups = [PeerGroup(0, G, g) for g in range(G)]
PeerGroup.connect_local(groups)
streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
ins, outs = [], []
for g in range(G):
    sc = torch.sort(torch.rand((Q, k), device=dev), dim=1, descending=True)[0]
    rw = (torch.arange(Q * k, device=dev, dtype=torch.int64).reshape(Q, k) * G + g)
    nn_ = torch.full((Q,), k, dtype=torch.int32, device=dev)
    ins.append((sc, rw, nn_))
    outs.append((torch.empty_like(sc), torch.empty_like(rw), torch.empty_like(nn_)))
torch.cuda.synchronize()
for rep in range(3):
    for g in range(G):
        check(lib.cqs_b200_peer_gather_merge(groups[g]._h, C.c_void_p(ins[g][0].data_ptr()), C.c_void_p(ins[g][1].data_ptr()),
                                             C.c_void_p(ins[g][2].data_ptr()), Q, k, C.c_void_p(outs[g][0].data_ptr()),
                                             C.c_void_p(outs[g][1].data_ptr()), C.c_void_p(outs[g][2].data_ptr()),
                                             C.c_void_p(streams[g].cuda_stream)))
torch.cuda.synchronize()
assert all(g.status() == 0 for g in groups) and torch.equal(outs[0][1], outs[1][1])
for g in groups:
    g.close()
print("peer ok")
# sparse build (+ static block index), search, hybrid
rng = np.random.default_rng(0)
nd, vocab = 5000, 300
nnz_d = rng.integers(0, 40, nd)
indptr = np.zeros(nd + 1, np.uint64); indptr[1:] = np.cumsum(nnz_d)
tok = np.concatenate([np.sort(rng.choice(vocab, size=int(c), replace=False)) for c in nnz_d]).astype(np.uint32)
w = (rng.random(tok.shape[0]) + 0.01).astype(f32)
ix = cqs_b200.B200Index(dim)
ix.append(None, rows[:nd]); ix.finalize()
ix.sparse_attach(indptr, tok, w, vocab)
qt = np.asarray([0, 5, 7, 100, 299], np.uint32); qw = np.asarray([1, .5, .2, .9, .3], f32)
o_rows, o_sc = O.sparse_search_csr(indptr, tok, w, qt, qw, nd, 500)
g_rows, g_sc = ix.search_sparse_rows(qt, qw, 500)
assert g_rows.astype(np.int64).tolist() == o_rows.tolist()
got = ix.search_hybrid_rows(qs[0], qt, qw, 0.8, 500)
assert got["rows"].shape[0] == 500
ix.close()
print("sparse ok")
