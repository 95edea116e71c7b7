"""N > 1 path.  CPU: world_size-2 gloo test of the host-side logic (partition,
all-gather plumbing, merge rule) with the oracle standing in for the per-shard
scan.  GPU: the same corpus split into shards inside ONE process (emulating the
ranks), per-shard device top-k, concatenation = the all-gather result, merge
kernel — must equal the unsharded answer."""
import os
import socket

import numpy as np
import pytest

from oracle import cqs_oracle as O

f32 = np.float32


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, k, out):
    import torch.distributed as dist
    import torch
    from cqs_b200.sharded import shard_range, merge_topk_host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = O.fast_unit_rows(n, 64, seed=3)
    rows[10] = rows[n - 5]                                   # a cross-shard exact tie
    queries = O.fast_unit_rows(6, 64, seed=4)
    queries[0] = rows[10]
    row0, nl = shard_range(n, world, rank)
    sc = np.full((6, k), -np.inf, f32)
    rw = np.full((6, k), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    for qi in range(6):
        r, s = O.brute_force_search(rows[row0:row0 + nl], queries[qi], k)   # stand-in for the shard scan
        sc[qi, :r.shape[0]] = s
        rw[qi, :r.shape[0]] = (r + row0).astype(np.uint64)
    g_sc = [torch.empty((6, k), dtype=torch.float32) for _ in range(world)]
    g_rw = [torch.empty((6, k), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(g_sc, torch.from_numpy(sc))
    dist.all_gather(g_rw, torch.from_numpy(rw.view(np.int64)))
    ok = True
    for qi in range(6):
        s, r = merge_topk_host(np.stack([t[qi].numpy() for t in g_sc]),
                               np.stack([t[qi].numpy().view(np.uint64) for t in g_rw]), k)
        er, es = O.brute_force_search(rows, queries[qi], k)
        ok &= r.astype(np.int64).tolist() == er.tolist()
        ok &= np.array_equal(s.view(np.uint32), es.view(np.uint32))
    out[rank] = ok
    dist.destroy_process_group()


def test_gloo_world2_partition_gather_merge():
    import torch.multiprocessing as mp
    from cqs_b200.sharded import shard_range
    assert shard_range(10, 3, 0) == (0, 4) and shard_range(10, 3, 2) == (8, 2) and shard_range(3, 8, 7) == (3, 0)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5003, 20, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out[0] and out[1]


def test_merge_topk_host_rule():
    from cqs_b200.sharded import merge_topk_host
    sc = np.array([[0.5, 0.4, -np.inf], [0.5, 0.45, 0.1]], f32)
    rw = np.array([[9, 2, 2**64 - 1], [3, 7, 8]], np.uint64)
    s, r = merge_topk_host(sc, rw, 4)
    assert r.tolist() == [3, 9, 7, 2] and s.tolist() == [0.5, 0.5, f32(0.45), f32(0.4)]


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_emulated_shards_device_merge_equals_unsharded(storage):
    import ctypes as C
    import torch
    import cqs_b200
    from cqs_b200.capi import lib, check
    from cqs_b200.sharded import shard_range
    n, dim, k, G, Q = 50_003, 768, 20, 4, 8
    rows = O.fast_unit_rows(n, dim, seed=71)
    rows[17] = rows[n - 3]                                   # cross-shard tie
    queries = O.fast_unit_rows(Q, dim, seed=72)
    queries[0] = rows[17]
    whole = cqs_b200.B200Index(dim, storage=storage)
    whole.append(None, rows); whole.finalize()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    d_q = torch.from_numpy(queries).to(dev)
    g_sc = torch.empty((G, Q, k), dtype=torch.float32, device=dev)
    g_rw = torch.empty((G, Q, k), dtype=torch.int64, device=dev)
    g_n = torch.empty((G, Q), dtype=torch.int32, device=dev)
    shards = []
    with torch.cuda.stream(stream):
        sp = C.c_void_p(stream.cuda_stream)
        for g in range(G):
            row0, nl = shard_range(n, G, g)
            ix = cqs_b200.B200Index(dim, storage=storage, row_base=row0)
            ix.append(None, rows[row0:row0 + nl]); ix.finalize()
            shards.append(ix)
            for qi in range(Q):
                check(lib.cqs_b200_search_device(ix._h, C.c_void_p(d_q.data_ptr() + qi * dim * 4), k, None,
                                                 C.c_void_p(g_sc[g, qi].data_ptr()), C.c_void_p(g_rw[g, qi].data_ptr()),
                                                 C.c_void_p(g_n[g, qi].data_ptr()), sp))
        m_sc = torch.empty((Q, k), dtype=torch.float32, device=dev)
        m_rw = torch.empty((Q, k), dtype=torch.int64, device=dev)
        m_n = torch.empty((Q,), dtype=torch.int32, device=dev)
        check(lib.cqs_b200_merge_topk_device(0, C.c_void_p(g_sc.data_ptr()), C.c_void_p(g_rw.data_ptr()), G, Q, k,
                                             C.c_void_p(m_sc.data_ptr()), C.c_void_p(m_rw.data_ptr()),
                                             C.c_void_p(m_n.data_ptr()), sp))
    stream.synchronize()
    for qi in range(Q):
        a, b = whole.search_rows(queries[qi], k)
        assert int(m_n[qi]) == k
        assert m_rw[qi].cpu().numpy().view(np.uint64).tolist() == a.tolist()
        assert np.array_equal(m_sc[qi].cpu().numpy().view(np.uint32), b.view(np.uint32))
    assert m_rw[0, :2].cpu().tolist() == [17, n - 3]
    for ix in shards + [whole]:
        ix.close()


@pytest.mark.gpu
def test_in_process_multi_device_index_equals_single():
    """cqs_b200_create(device_ids, n_dev > 1): the form the single-process daemon would use —
    contiguous row blocks per device, host merge.  On a 1-GPU box the shards are two blocks on
    cuda:0 (same code path: per-shard launch, host-mapped results, host merge)."""
    import torch
    import cqs_b200
    n, dim = 30_011, 768
    rows = O.fast_unit_rows(n, dim, seed=91)
    rows[5] = rows[n - 2]
    one = cqs_b200.B200Index(dim, devices=[0])
    one.append(None, rows); one.finalize()
    nd = min(torch.cuda.device_count(), 4)
    devices = list(range(nd)) if nd >= 2 else [0, 0, 0]
    multi = cqs_b200.B200Index(dim, devices=devices)
    multi.reserve(n)
    multi.append(None, rows[:10_000]); multi.append(None, rows[10_000:])
    multi.finalize()
    assert len(multi) == n
    rng = np.random.default_rng(0)
    mask = rng.random(n) < 0.5
    for qi in (5, 100, 29_999):
        for k in (1, 20, 500):
            a, b = one.search_rows(rows[qi], k)
            c, d = multi.search_rows(rows[qi], k)
            assert np.array_equal(a, c) and np.array_equal(b.view(np.uint32), d.view(np.uint32))
        a, b = one.search_rows(rows[qi], 20, O.mask_to_bitset(mask))
        c, d = multi.search_rows(rows[qi], 20, O.mask_to_bitset(mask))
        assert np.array_equal(a, c)
    one.close(); multi.close()


@pytest.mark.gpu
def test_search_device_and_host_search_interleave_safely():
    """Launches on a caller stream (search_device) and on the index's own stream share the
    per-index scratch; the library must order them."""
    import ctypes as C
    import torch
    import cqs_b200
    from cqs_b200.capi import lib, check
    n, dim, k = 200_000, 768, 20
    rows = O.fast_unit_rows(n, dim, seed=93)
    ix = cqs_b200.B200Index(dim)
    ix.append(None, rows); ix.finalize()
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    qs = O.fast_unit_rows(16, dim, seed=94)
    d_q = torch.from_numpy(qs).to(dev)
    d_sc = torch.empty((16, k), dtype=torch.float32, device=dev)
    d_rw = torch.empty((16, k), dtype=torch.int64, device=dev)
    d_n = torch.empty((16,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    host = []
    for i in range(16):
        check(lib.cqs_b200_search_device(ix._h, C.c_void_p(d_q.data_ptr() + i * dim * 4), k, None,
                                         C.c_void_p(d_sc[i].data_ptr()), C.c_void_p(d_rw[i].data_ptr()),
                                         C.c_void_p(d_n[i].data_ptr()), C.c_void_p(st.cuda_stream)))
        host.append(ix.search_rows(qs[15 - i], k))          # own stream, interleaved
    st.synchronize()
    for i in range(16):
        a, b = ix.search_rows(qs[i], k)
        assert d_rw[i].cpu().numpy().view(np.uint64).tolist() == a.tolist()
        assert np.array_equal(host[15 - i][0], a)
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [20, 500])
def test_search_device_on_two_alternating_streams(k):
    """Launches issued alternately on two streams use the two scratch sets of the index: the tail
    of one launch may overlap the next launch's streaming phase, results must not change."""
    import ctypes as C
    import torch
    import cqs_b200
    from cqs_b200.capi import lib, check
    n, dim, Q = 300_000, 768, 48
    rows = O.fast_unit_rows(n, dim, seed=95)
    ix = cqs_b200.B200Index(dim)
    ix.append(None, rows); ix.finalize()
    dev = torch.device("cuda", 0)
    lanes = [torch.cuda.Stream(device=dev) for _ in range(2)]
    qs = O.fast_unit_rows(Q, dim, seed=96)
    d_q = torch.from_numpy(qs).to(dev)
    d_sc = torch.empty((Q, k), dtype=torch.float32, device=dev)
    d_rw = torch.empty((Q, k), dtype=torch.int64, device=dev)
    d_n = torch.empty((Q,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        for i in range(Q):
            lane = lanes[i % 2] if rep == 0 else lanes[(i // 3) % 2]      # in and out of phase with the scratch sets
            check(lib.cqs_b200_search_device(ix._h, C.c_void_p(d_q.data_ptr() + i * dim * 4), k, None,
                                             C.c_void_p(d_sc[i].data_ptr()), C.c_void_p(d_rw[i].data_ptr()),
                                             C.c_void_p(d_n[i].data_ptr()), C.c_void_p(lane.cuda_stream)))
        torch.cuda.synchronize()
        for i in range(Q):
            a, b = ix.search_rows(qs[i], k)
            assert d_rw[i].cpu().numpy().view(np.uint64).tolist() == a.tolist()
            assert np.array_equal(d_sc[i].cpu().numpy().view(np.uint32), b.view(np.uint32))
    ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,storage", [(768, "f32"), (100, "f32"), (768, "bf16")])
def test_exact_batch_is_pipelined_single_queries(dim, storage):
    """cqs_b200_search_batch on f32 storage (and small bf16 batches) = nq exact scans issued on two
    launch lanes with one H2D / one D2H; cqs_b200_search_many_device is the device-resident form.
    Both must equal nq cqs_b200_search calls, including bitset, padded dims and non-finite queries."""
    import ctypes as C
    import torch
    import cqs_b200
    from cqs_b200.capi import lib, check
    n, k = 60_001, 20
    nq = 37 if storage == "f32" else 7                       # bf16 with nq < 8 stays on the exact path
    rows = O.fast_unit_rows(n, dim, seed=97)
    rows[4] = rows[n - 1]
    ix = cqs_b200.B200Index(dim, storage=storage)
    ix.append(None, rows); ix.finalize()
    qs = O.fast_unit_rows(nq, dim, seed=98)
    qs[0] = rows[4]
    qs[3, 1] = np.inf                                        # -> empty result, the others unaffected
    mask = np.random.default_rng(5).random(n) < 0.4
    for bits in (None, O.mask_to_bitset(mask)):
        r, s, nn = ix.search_batch_rows(qs, k, bits)
        for i in range(nq):
            a, b = ix.search_rows(qs[i], k, bits)
            assert int(nn[i]) == a.shape[0]
            assert np.array_equal(r[i, :nn[i]], a) and np.array_equal(s[i, :nn[i]].view(np.uint32), b.view(np.uint32))
        assert int(nn[3]) == 0
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    good = np.delete(qs, 3, axis=0)
    d_q = torch.from_numpy(good).to(dev)
    m = good.shape[0]
    d_sc = torch.empty((m, k), dtype=torch.float32, device=dev)
    d_rw = torch.empty((m, k), dtype=torch.int64, device=dev)
    d_n = torch.empty((m,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        check(lib.cqs_b200_search_many_device(ix._h, None, C.c_void_p(d_q.data_ptr()), m, k, None,
                                              C.c_void_p(d_sc.data_ptr()), C.c_void_p(d_rw.data_ptr()),
                                              C.c_void_p(d_n.data_ptr()), C.c_void_p(st.cuda_stream)))
    st.synchronize()
    for i in range(m):
        a, b = ix.search_rows(good[i], k)
        assert d_rw[i].cpu().numpy().view(np.uint64).tolist() == a.tolist()
        assert np.array_equal(d_sc[i].cpu().numpy().view(np.uint32), b.view(np.uint32))
    ix.close()


def _peer_setup_worker(rank, world, port, out):
    import torch.distributed as dist
    from cqs_b200.sharded import PeerGroup
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # no CUDA device here: creating the mailbox fails on every rank; the ranks must AGREE on the failure
    # (no rank may be left waiting in the handle exchange) and report it the way the caller asked
    res = PeerGroup.from_dist(dist, 0, strict=False)
    raised = False
    try:
        PeerGroup.from_dist(dist, 0, strict=True)
    except RuntimeError as e:
        raised = "peer group setup failed" in str(e) and "rank 0" in str(e) and "rank 1" in str(e)
    out[rank] = (res is None) and raised
    dist.destroy_process_group()


def test_gloo_world2_peer_setup_failure_is_agreed_not_hung():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA (the failure path)")
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_peer_setup_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out[0] and out[1]
