"""Row-sharded search over NVLink peer memory (csrc/peer.cuh): the
exchange + merge that rides in the kernels must give, on EVERY rank, exactly the
unsharded answer (same rows, bit-identical scores, reference order
(score desc, id asc) — candidate.rs:321-329).

One GPU is enough for: the fused tail with world = 1, the stand-alone gather+merge kernel
with the ranks emulated on one device (one launch per rank, separate streams), a missing
rank (timeout instead of a hang) and the poison hook.

The tests that run SEVERAL mutually waiting launches per rank (fused scan with launch lanes,
sharded hybrid, sharded batch with re-runs, shadow-scan flag propagation, the two-process CUDA
IPC group) need every rank on its own GPU and skip below 2 GPUs.  Emulating them on one device
was tried in round 2 and is not reliable: kernels of different ranks that wait on each other
are separate launches, and on one GPU nothing guarantees they run at the same time — streams
share hardware queues, so rank B's launch can sit behind a launch of rank A that is itself
waiting (through a stream event) for the kernel that spins for B: a deadlock until the exchange
times out; with processes the driver's context-switch timeout fires (Xid 109,
B200_PROFILING.md).  Their multi-GPU run is committed under profiles/ (r02_pytest_multi_gpu.log);
the driver sees the multi-rank path through bench.py's in-run parity keys at N = 2, 4, 8, and the
host-side logic runs on CPU with gloo (tests/test_sharded.py)."""
import ctypes as C
import os
import socket
import threading

import numpy as np
import pytest

from oracle import cqs_oracle as O

pytestmark = pytest.mark.gpu
f32 = np.float32
EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def _vp(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_world1_fused_tail_equals_plain_search(storage):
    import cqs_b200
    from cqs_b200.sharded import PeerGroup, search_sharded
    n, dim = 40_007, 768
    rows = O.fast_unit_rows(n, dim, seed=11)
    rows[3] = rows[n - 9]
    ix = cqs_b200.B200Index(dim, storage=storage, row_base=1000)
    ix.append(None, rows); ix.finalize()
    pg = PeerGroup(0, 1, 0)
    mask = np.random.default_rng(1).random(n) < 0.3
    for qi in (3, 77, n - 1):
        for k in (1, 20, 33, 500, 1024):
            a, b = ix.search_rows(rows[qi], k)
            c, d = search_sharded(ix, pg, rows[qi], k)
            assert np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32))
        a, b = ix.search_rows(rows[qi], 20, O.mask_to_bitset(mask))
        c, d = search_sharded(ix, pg, rows[qi], 20, O.mask_to_bitset(mask))
        assert np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32))
    bad = rows[0].copy(); bad[5] = np.nan
    c, d = search_sharded(ix, pg, bad, 20)
    assert c.shape[0] == 0                                  # non-finite query -> empty (src/cagra.rs:458-470)
    assert pg.status() == 0
    pg.close(); ix.close()


def _random_sorted_lists(rng, G, Q, k, tie_every=7):
    """Per emulated rank: Q sorted lists of <= k (score, global row) pairs with cross-list score ties."""
    sc = np.full((G, Q, k), -np.inf, f32)
    rw = np.full((G, Q, k), EMPTY, np.uint64)
    nn = np.zeros((G, Q), np.uint32)
    pool = (rng.standard_normal(64) * 0.3).astype(f32)     # few distinct scores -> many exact ties
    for g in range(G):
        for q in range(Q):
            n = int(rng.integers(0, k + 1)) if q % 5 else k
            s = np.where(rng.random(n) < 1.0 / tie_every, rng.choice(pool, n), rng.standard_normal(n).astype(f32) * 0.3)
            r = rng.choice(1_000_000, n, replace=False).astype(np.uint64) * np.uint64(G) + np.uint64(g)  # disjoint across ranks
            order = np.lexsort((r, -s.astype(np.float64)))
            sc[g, q, :n] = s[order].astype(f32)
            rw[g, q, :n] = r[order]
            nn[g, q] = n
    return sc, rw, nn


@pytest.mark.parametrize("G,Q,k", [(2, 8, 20), (4, 140, 20), (8, 64, 64), (3, 5, 1024), (2, 1024, 20)])
def test_gather_merge_emulated_ranks_on_one_device(G, Q, k):
    import torch
    from cqs_b200.capi import lib, check
    from cqs_b200.sharded import PeerGroup, merge_topk_host
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(G * 1000 + Q + k)
    sc, rw, nn = _random_sorted_lists(rng, G, Q, k)
    groups = [PeerGroup(0, G, g, max_elems=max(Q * k, 1024)) for g in range(G)]
    PeerGroup.connect_local(groups)
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    d_in = [(torch.from_numpy(sc[g]).to(dev), torch.from_numpy(rw[g].view(np.int64)).to(dev),
             torch.from_numpy(nn[g].view(np.int32)).to(dev)) for g in range(G)]
    d_out = [(torch.empty((Q, k), dtype=torch.float32, device=dev), torch.empty((Q, k), dtype=torch.int64, device=dev),
              torch.empty((Q,), dtype=torch.int32, device=dev)) for g in range(G)]
    torch.cuda.synchronize()
    for rep in range(3):                                    # three exchanges: both mailbox slots get reused
        for g in range(G):
            check(lib.cqs_b200_peer_gather_merge(groups[g]._h, _vp(d_in[g][0]), _vp(d_in[g][1]), _vp(d_in[g][2]), Q, k,
                                                 _vp(d_out[g][0]), _vp(d_out[g][1]), _vp(d_out[g][2]),
                                                 C.c_void_p(streams[g].cuda_stream)))
    torch.cuda.synchronize()
    for g in range(G):
        assert groups[g].status() == 0
        o_s = d_out[g][0].cpu().numpy(); o_r = d_out[g][1].cpu().numpy().view(np.uint64); o_n = d_out[g][2].cpu().numpy()
        for q in range(Q):
            es, er = merge_topk_host(sc[:, q], rw[:, q], k)
            assert int(o_n[q]) == er.shape[0]
            assert o_r[q, :er.shape[0]].tolist() == er.tolist()
            assert np.array_equal(o_s[q, :er.shape[0]].view(np.uint32), es.view(np.uint32))
            assert (o_r[q, er.shape[0]:] == EMPTY).all()
    for g in groups:
        g.close()


def test_missing_rank_times_out_instead_of_hanging():
    import torch
    from cqs_b200.capi import lib, check, B200Error
    from cqs_b200.sharded import PeerGroup
    dev = torch.device("cuda", 0)
    groups = [PeerGroup(0, 2, g) for g in range(2)]
    PeerGroup.connect_local(groups)
    groups[0].set_timeout_ms(200)
    Q, k = 4, 20
    d_s = torch.zeros((Q, k), dtype=torch.float32, device=dev)
    d_r = torch.arange(Q * k, dtype=torch.int64, device=dev).reshape(Q, k)
    d_n = torch.full((Q,), k, dtype=torch.int32, device=dev)
    o_s = torch.empty_like(d_s); o_r = torch.empty_like(d_r); o_n = torch.empty_like(d_n)
    st = torch.cuda.Stream(device=dev)
    check(lib.cqs_b200_peer_gather_merge(groups[0]._h, _vp(d_s), _vp(d_r), _vp(d_n), Q, k, _vp(o_s), _vp(o_r), _vp(o_n),
                                         C.c_void_p(st.cuda_stream)))           # rank 1 never shows up
    st.synchronize()
    assert groups[0].status() == 1
    assert o_n.cpu().tolist() == [0] * Q
    with pytest.raises(B200Error):                           # the group is failed from now on
        check(lib.cqs_b200_peer_gather_merge(groups[0]._h, _vp(d_s), _vp(d_r), _vp(d_n), Q, k, _vp(o_s), _vp(o_r),
                                             _vp(o_n), C.c_void_p(st.cuda_stream)))
    for g in groups:
        g.close()


def test_peer_failure_reaches_is_poisoned():
    """A rank that never shows up: the sharded search times out, returns an error, and the failure is
    visible through VectorIndex::is_poisoned() — the reference's only recovery hook
    (src/index.rs:203-205 -> daemon rebuild, src/cli/batch/view.rs:737-767)."""
    import cqs_b200
    from cqs_b200.capi import B200Error
    from cqs_b200.sharded import PeerGroup, search_sharded
    rows = O.fast_unit_rows(5000, 768, seed=3)
    ix = cqs_b200.B200Index(768, row_base=0)
    ix.append(None, rows); ix.finalize()
    groups = [PeerGroup(0, 2, g) for g in range(2)]
    PeerGroup.connect_local(groups)
    groups[0].set_timeout_ms(200)
    assert not ix.is_poisoned()
    with pytest.raises(B200Error):
        search_sharded(ix, groups[0], rows[1], 20)           # rank 1 never searches
    assert ix.is_poisoned()                                   # -> the shim's is_poisoned() -> rebuild
    assert ix.search(rows[1], 5) == []                        # every later call: error -> empty Vec
    with pytest.raises(B200Error):
        search_sharded(ix, groups[0], rows[1], 20)
    for g in groups:
        g.close()
    ix.close()


def _warm(ix, dim, batch=False, sparse_nnz=0):
    """Touch the library's lazily allocated scratch (tensor-core batch scratch, sparse bounds) once per
    shard before the rank threads start, so that no rank sits in a cudaMalloc while its peers' exchange
    kernels already wait for it (with several ranks in ONE process a device allocation can serialise
    against running kernels)."""
    if batch:
        ix.search_batch_rows(O.fast_unit_rows(8, dim, seed=999), 20)
    if sparse_nnz:
        ix.search_sparse_rows(np.arange(sparse_nnz, dtype=np.uint32), np.ones(sparse_nnz, f32), 500)


def _rank_devices(max_ranks=4):
    """Device of every rank, one GPU each; skips on a single-GPU box (see the module docstring)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs one GPU per rank (>= 2 GPUs)")
    return list(range(min(n, max_ranks)))


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_ranks_in_process_fused_scan_exchange_merge(storage):
    import torch
    import cqs_b200
    from cqs_b200.capi import lib, check
    from cqs_b200.sharded import PeerGroup, shard_range, search_sharded, search_batch_sharded
    devs = _rank_devices()
    G = len(devs)
    n, dim, k, Q = 120_011, 768, 20, 12
    rows = O.fast_unit_rows(n, dim, seed=31)
    rows[17] = rows[n - 3]                                   # cross-shard exact tie
    queries = O.fast_unit_rows(Q, dim, seed=32)
    queries[0] = rows[17]
    whole = cqs_b200.B200Index(dim, storage=storage, devices=[0])
    whole.append(None, rows); whole.finalize()
    shards, groups = [], []
    for g in range(G):
        row0, nl = shard_range(n, G, g)
        ix = cqs_b200.B200Index(dim, storage=storage, devices=[devs[g]], row_base=row0)
        ix.append(None, rows[row0:row0 + nl]); ix.finalize()
        _warm(ix, dim, batch=True)
        shards.append(ix)
        groups.append(PeerGroup(devs[g], G, g))
    PeerGroup.connect_local(groups)
    # (a) asynchronous device entry: one launch per (rank, query), nothing else
    outs = []
    for g in range(G):
        dev = torch.device("cuda", devs[g])
        outs.append((torch.from_numpy(queries).to(dev), torch.empty((Q, k), dtype=torch.float32, device=dev),
                     torch.empty((Q, k), dtype=torch.int64, device=dev), torch.empty((Q,), dtype=torch.int32, device=dev)))
    for g in range(G):
        torch.cuda.synchronize(devs[g])
    # two launch lanes per rank: the exchange of query i overlaps the scan of query i+1
    lanes = [[torch.cuda.Stream(device=torch.device("cuda", devs[g])) for _ in range(2)] for g in range(G)]
    for qi in range(Q):
        for g in range(G):
            d_q, d_s, d_r, d_n = outs[g]
            check(lib.cqs_b200_search_sharded_device(shards[g]._h, groups[g]._h, C.c_void_p(d_q.data_ptr() + qi * dim * 4), k, None,
                                                     C.c_void_p(d_s[qi].data_ptr()), C.c_void_p(d_r[qi].data_ptr()),
                                                     C.c_void_p(d_n[qi].data_ptr()), C.c_void_p(lanes[g][qi % 2].cuda_stream)))
    for g in range(G):
        torch.cuda.synchronize(devs[g])
    for qi in range(Q):
        a, b = whole.search_rows(queries[qi], k)
        for g in range(G):
            _, d_s, d_r, d_n = outs[g]
            assert int(d_n[qi]) == k
            assert d_r[qi].cpu().numpy().view(np.uint64).tolist() == a.tolist()
            assert np.array_equal(d_s[qi].cpu().numpy().view(np.uint32), b.view(np.uint32))
    assert outs[0][2][0, :2].cpu().tolist() == [17, n - 3]
    # (b) blocking host entry from one thread per rank (the daemon's client threads), k up to max_k
    res = [None] * G

    def run(g):
        out = []
        for qi in range(Q):
            for kk in (1, 20, 500):
                out.append(search_sharded(shards[g], groups[g], queries[qi], kk))
        out.append(search_batch_sharded(shards[g], groups[g], queries, k))   # bf16: tensor cores; f32: pipelined exact scans
        res[g] = out

    th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
        assert not t.is_alive()
    i = 0
    for qi in range(Q):
        for kk in (1, 20, 500):
            a, b = whole.search_rows(queries[qi], kk)
            for g in range(G):
                c, d = res[g][i]
                assert np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32))
            i += 1
    for g in range(G):
        r, s, nn = res[g][i]
        for qi in range(Q):
            a, b = whole.search_rows(queries[qi], k)
            assert int(nn[qi]) == k and np.array_equal(r[qi], a) and np.array_equal(s[qi].view(np.uint32), b.view(np.uint32))
    for g in range(G):
        assert groups[g].status() == 0
        groups[g].close(); shards[g].close()
    whole.close()


def test_ranks_shadow_scan_flag_travels_with_the_exchange():
    """STORAGE_BF16_F32 shards: the shadow scan + f32 rescoring runs per shard inside the fused
    kernel; a shard that cannot prove its list raises bit 31 of its list length, the bit reaches every
    rank with the exchange, and every rank repeats the search on its f32 rows.  Result == unsharded f32."""
    import cqs_b200
    from cqs_b200.sharded import PeerGroup, shard_range, search_sharded, search_batch_sharded
    devs = _rank_devices(2)
    G = len(devs)
    rng = np.random.default_rng(19)
    n, dim = 60_000, 768
    rows = O.fast_unit_rows(n, dim, seed=61)
    base = rows[11].copy()
    for j in range(150):                                    # near-duplicates, all inside shard 0
        v = base + rng.standard_normal(dim).astype(f32) * f32(2e-5)
        rows[50 + 13 * j] = v / np.linalg.norm(v)
    qs = O.fast_unit_rows(6, dim, seed=62)
    qs[0] = base
    whole = cqs_b200.B200Index(dim, storage="f32", devices=[0])
    whole.append(None, rows); whole.finalize()
    shards, groups = [], []
    for g in range(G):
        row0, nl = shard_range(n, G, g)
        ix = cqs_b200.B200Index(dim, storage="bf16+f32", devices=[devs[g]], row_base=row0)
        ix.append(None, rows[row0:row0 + nl]); ix.finalize()
        shards.append(ix); groups.append(PeerGroup(devs[g], G, g))
    PeerGroup.connect_local(groups)
    res = [None] * G

    def run(g):
        out = [search_sharded(shards[g], groups[g], qs[i], kk) for i in range(6) for kk in (20, 100)]
        out.append(search_batch_sharded(shards[g], groups[g], qs[:4], 20))     # < 8 queries: exact lanes
        res[g] = out

    th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
        assert not t.is_alive()
    i = 0
    for qi in range(6):
        for kk in (20, 100):
            a, b = whole.search_rows(qs[qi], kk)
            for g in range(G):
                c, d = res[g][i]
                assert np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32)), (qi, kk, g)
            i += 1
    for g in range(G):
        r, s, nn = res[g][i]
        for qi in range(4):
            a, b = whole.search_rows(qs[qi], 20)
            assert int(nn[qi]) == 20 and np.array_equal(r[qi], a) and np.array_equal(s[qi].view(np.uint32), b.view(np.uint32))
    for g in range(G):
        assert groups[g].status() == 0
        groups[g].close(); shards[g].close()
    whole.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _ipc_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import cqs_b200
    from cqs_b200.sharded import PeerGroup, shard_range, search_sharded, search_batch_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank
    torch.cuda.set_device(dev)
    n, dim, k, Q = 90_001, 768, 20, 40
    rows = O.fast_unit_rows(n, dim, seed=51)
    rows[5] = rows[n - 2]
    queries = O.fast_unit_rows(Q, dim, seed=52)
    queries[0] = rows[5]
    ok = True
    for storage in ("f32", "bf16"):
        row0, nl = shard_range(n, world, rank)
        ix = cqs_b200.B200Index(dim, storage=storage, devices=[dev], row_base=row0)
        ix.append(None, rows[row0:row0 + nl]); ix.finalize()
        whole = cqs_b200.B200Index(dim, storage=storage, devices=[dev])
        whole.append(None, rows); whole.finalize()
        pg = PeerGroup.from_dist(dist, dev)                  # CUDA IPC handles over all_gather_object
        for qi in range(Q):
            a, b = whole.search_rows(queries[qi], k)
            c, d = search_sharded(ix, pg, queries[qi], k)
            ok &= bool(np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32)))
        r, s, nn = search_batch_sharded(ix, pg, queries, k)
        for qi in range(Q):
            a, b = whole.search_rows(queries[qi], k)
            ok &= bool(int(nn[qi]) == k and np.array_equal(r[qi], a) and np.array_equal(s[qi].view(np.uint32), b.view(np.uint32)))
        ok &= pg.status() == 0
        dist.barrier()
        pg.close(); ix.close(); whole.close()
    out[rank] = ok
    dist.destroy_process_group()


def test_two_processes_cuda_ipc_mailboxes():
    _rank_devices(2)
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert out[0] and out[1]


def test_ranks_sharded_hybrid_equals_unsharded():
    """Dense pool from the scan's own exchange + sparse pool through the gather+merge kernel + the same
    fusion on every rank == cqs_b200_search_hybrid on the whole corpus (src/search/query.rs:914-1005)."""
    import cqs_b200
    from cqs_b200.sharded import PeerGroup, shard_range
    devs = _rank_devices()
    G = len(devs)
    rng = np.random.default_rng(77)
    n, dim, vocab, pool = 24_000, 768, 3000, 500
    rows = O.fast_unit_rows(n, dim, seed=77, clustered=True)
    nnz_d = rng.integers(0, 40, n)
    indptr = np.zeros(n + 1, np.uint64); indptr[1:] = np.cumsum(nnz_d)
    tok = np.concatenate([np.sort(rng.choice(vocab, size=int(c), replace=False)) for c in nnz_d]).astype(np.uint32)
    w = (rng.random(tok.shape[0]) + 0.01).astype(f32)
    whole = cqs_b200.B200Index(dim, devices=[0])
    whole.append(None, rows); whole.finalize()
    whole.sparse_attach(indptr, tok, w, vocab)
    shards, groups = [], []
    for g in range(G):
        row0, nl = shard_range(n, G, g)
        ix = cqs_b200.B200Index(dim, devices=[devs[g]], row_base=row0)
        ix.append(None, rows[row0:row0 + nl]); ix.finalize()
        lo, hi = int(indptr[row0]), int(indptr[row0 + nl])
        ix.sparse_attach(indptr[row0:row0 + nl + 1] - indptr[row0], tok[lo:hi], w[lo:hi], vocab)
        _warm(ix, dim, sparse_nnz=48)
        shards.append(ix); groups.append(PeerGroup(devs[g], G, g))
    PeerGroup.connect_local(groups)
    cases = []
    for trial in range(6):
        q = rows[int(rng.integers(0, n))] + rng.standard_normal(dim).astype(f32) * f32(0.02)
        q = (q / np.linalg.norm(q)).astype(f32)
        qn = int(rng.integers(1, 48))
        qt = np.sort(rng.choice(vocab, size=qn, replace=False)).astype(np.uint32)
        cases.append((q, qt, rng.random(qn).astype(f32), [0.85, 0.0, 1.0, 0.6, 0.1, 0.8][trial]))
    bad = cases[0][0].copy(); bad[0] = np.nan
    cases.append((bad, cases[0][1], cases[0][2], 0.8))        # empty dense pool, sparse leg still runs
    cases.append((cases[1][0], np.zeros(0, np.uint32), np.zeros(0, f32), 0.8))   # no sparse query
    res = [None] * G

    def run(g):
        res[g] = [shards[g].search_hybrid_rows(q, qt, qw, a, pool, peer=groups[g]) for q, qt, qw, a in cases]

    th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
        assert not t.is_alive()
    for i, (q, qt, qw, a) in enumerate(cases):
        want = whole.search_hybrid_rows(q, qt, qw, a, pool)
        for g in range(G):
            got = res[g][i]
            assert np.array_equal(got["rows"], want["rows"]), (i, g)
            for key in ("fused", "dense", "sparse_raw"):
                assert np.array_equal(got[key].view(np.uint32), want[key].view(np.uint32)), (i, g, key)
            assert np.array_equal(got["present"], want["present"])
    for g in range(G):
        assert groups[g].status() == 0
        groups[g].close(); shards[g].close()
    whole.close()


def test_ranks_sharded_batch_with_exact_fallback():
    """Near-duplicate rows inside one shard: that shard's tensor-core candidate pool cannot be proven
    complete for the matching queries, they are re-run through the exact scan and patched into the
    shard's lists BEFORE the exchange; the merged result must still equal the unsharded exact answer."""
    import ctypes as C
    import cqs_b200
    from cqs_b200.capi import lib
    from cqs_b200.sharded import PeerGroup, shard_range, search_batch_sharded
    G = 2
    devs = _rank_devices(2)
    rng = np.random.default_rng(5)
    n, dim, k = 80_000, 768, 20
    rows = O.fast_unit_rows(n, dim, seed=41)
    base = rows[7].copy()
    for j in range(300):                                   # all inside shard 0 (rows < 40,000)
        v = base + rng.standard_normal(dim).astype(f32) * f32(2e-5)
        rows[100 + j * 100] = v / np.linalg.norm(v)
    q = O.fast_unit_rows(32, dim, seed=42)
    q[0] = base
    q[1] = rows[100]
    whole = cqs_b200.B200Index(dim, storage="bf16", devices=[0])
    whole.append(None, rows); whole.finalize()
    shards, groups = [], []
    for g in range(G):
        row0, nl = shard_range(n, G, g)
        ix = cqs_b200.B200Index(dim, storage="bf16", devices=[devs[g]], row_base=row0)
        ix.append(None, rows[row0:row0 + nl]); ix.finalize()
        _warm(ix, dim, batch=True)
        shards.append(ix); groups.append(PeerGroup(devs[g], G, g))
    PeerGroup.connect_local(groups)
    res = [None] * G

    def run(g):
        res[g] = search_batch_sharded(shards[g], groups[g], q, k)

    th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
        assert not t.is_alive()
    lib.cqs_b200_debug_batch_reruns.restype = C.c_uint32
    lib.cqs_b200_debug_batch_reruns.argtypes = [C.c_void_p]
    assert lib.cqs_b200_debug_batch_reruns(shards[0]._h) >= 1      # the fallback really ran on shard 0
    for qi in range(32):
        a, b = whole.search_rows(q[qi], k)
        for g in range(G):
            r, s, nn = res[g]
            assert int(nn[qi]) == k and np.array_equal(r[qi], a) and np.array_equal(s[qi].view(np.uint32), b.view(np.uint32)), (qi, g)
    for g in range(G):
        assert groups[g].status() == 0
        groups[g].close(); shards[g].close()
    whole.close()
