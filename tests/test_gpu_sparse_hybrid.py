"""GPU parity: SPLADE sparse leg, alpha fusion and centroid routing (kernel 4)
vs the oracle.  These are bit-exact: the kernels reproduce the reference's f32
operation order (src/splade/index.rs:251-259, src/search/query.rs:914-1005,
src/search/router.rs:1415-1444)."""
import numpy as np
import pytest

from oracle import cqs_oracle as O
from tests.parity import bits

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def cqs():
    import cqs_b200
    return cqs_b200


def _mk_index(cqs, ids, dim=8, seed=0):
    emb = O.fast_unit_rows(len(ids), dim, seed=seed)
    return cqs.B200Index.build(ids, emb), emb


# ---- reference unit tests, src/splade/index.rs:1114-1241 ----------------------------

def test_splade_reference_fixture(cqs):
    ix, _ = _mk_index(cqs, ["chunk_a", "chunk_b", "chunk_c"])
    sp = cqs.SpladeIndex(ix, [("chunk_a", [(1, .5), (2, .3), (3, .8)]),
                              ("chunk_b", [(1, .7), (4, .6)]),
                              ("chunk_c", [(2, .9), (3, .1), (5, .4)])])
    assert len(sp) == 3
    r = sp.search([(1, 1.0)], 10)                                   # test_build_and_search
    assert [x.id for x in r] == ["chunk_b", "chunk_a"]
    r = sp.search([(1, 1.0), (2, 1.0)], 10)                         # test_dot_product_correct
    assert [x.id for x in r] == ["chunk_c", "chunk_a", "chunk_b"]
    for x, want in zip(r, (0.9, 0.8, 0.7)):
        assert abs(x.score - want) < 1e-5
    r = sp.search_with_filter([(1, 1.0)], 10, lambda i: i == "chunk_a")   # test_search_filter
    assert [x.id for x in r] == ["chunk_a"]
    assert sp.search([(999, 1.0)], 10) == []                        # test_search_no_match
    assert sp.search([(40000, 1.0)], 10) == []                      # token beyond vocab
    assert sp.search([], 10) == []                                  # test_search_empty_query
    assert len(sp.search([(1, 1.0), (2, 1.0), (3, 1.0)], 2)) == 2   # test_search_respects_k
    assert sp.search([(1, 1.0), (2, 1.0), (3, 1.0)], 0) == []       # test_search_k_zero_returns_empty
    for w in (np.nan, np.inf):                                      # NaN / Inf weights: no crash, known ids
        for x in sp.search([(1, w)], 10):
            assert x.id in ("chunk_a", "chunk_b")
    assert not ix.is_poisoned()
    ix.close()


# ---- reference integration test, tests/search_test.rs:549-791 (legs_corpus) ----------

def test_fuse_pools_legs_corpus(cqs):
    from cqs_b200.index import fuse_pools
    # ids A=0, B=1, C=2; dense pool A .92, B .61; sparse pool B .8, C .5; alpha .5
    r = fuse_pools([0, 1], [0.92, 0.61], [1, 2], [0.8, 0.5], 0.5, 500)
    assert r["rows"].tolist() == [1, 0, 2]                          # B, A, C
    want = O.fuse_hybrid([(0, 0.92), (1, 0.61)], [(1, f32(0.8)), (2, f32(0.5))], 0.5, 500)
    assert [w["id"] for w in want] == [1, 0, 2]
    assert np.array_equal(bits(r["fused"]), bits([w["fused"] for w in want]))
    for got, exp in zip(r["fused"], (0.805, 0.46, 0.3125)):
        assert abs(float(got) - exp) < 1e-6
    assert r["dense"].tolist() == [f32(0.61), f32(0.92), 0.0]       # raw cosine, absent leg injected 0.0
    assert np.allclose(r["sparse_raw"], [0.8, 0.0, 0.5], atol=1e-7)
    assert r["present"].tolist() == [3, 1, 2]


def test_fuse_pools_modes_and_edges(cqs):
    from cqs_b200.index import fuse_pools

    def run(dense, sparse, alpha, pool_k):
        r = fuse_pools([d[0] for d in dense], [d[1] for d in dense], [s[0] for s in sparse],
                       [s[1] for s in sparse], alpha, pool_k)
        want = O.fuse_hybrid(dense, sparse, alpha, pool_k)
        assert r["rows"].tolist() == [w["id"] for w in want]
        assert np.array_equal(bits(r["fused"]), bits([w["fused"] for w in want]))
        assert np.array_equal(bits(r["dense"]), bits([w["dense"] for w in want]))
        assert np.array_equal(bits(r["sparse_raw"]), bits([w["sparse_raw"] for w in want]))
        assert r["present"].tolist() == [int(w["in_dense"]) | 2 * int(w["in_sparse"]) for w in want]

    run([(0, 0.5), (1, 0.4)], [(1, 2.0), (2, 1.0)], 0.0, 10)         # alpha <= 0: d + 0.1*s'
    run([(0, 0.5), (1, 0.4)], [(2, 0.0), (3, -1.0)], 0.5, 10)        # max_sparse <= 0 -> zeros
    run([(5, 0.3), (9, 0.3)], [], 1.0, 1)                            # tie -> row asc, truncate
    run([], [(4, 0.7), (2, 0.7)], 0.3, 10)                           # empty dense leg
    run([(0, -0.2), (1, 0.9)], [(0, 3.0)], 0.85, 10)                 # negative cosine stays raw
    rng = np.random.default_rng(5)
    for trial in range(6):
        nd, ns = int(rng.integers(0, 600)), int(rng.integers(0, 600))
        dr = rng.choice(2000, size=nd, replace=False)
        sr = rng.choice(2000, size=ns, replace=False)
        ds = np.sort(rng.uniform(-0.2, 1.0, nd).astype(f32))[::-1]
        ss = np.sort(rng.uniform(0, 30, ns).astype(f32))[::-1]
        alpha = float(rng.choice([0.0, 0.1, 0.6, 0.7, 0.8, 0.85, 1.0]))
        run(list(zip(dr.tolist(), ds.tolist())), list(zip(sr.tolist(), ss.tolist())), alpha, 500)


# ---- random CSR vs oracle (bit-exact), with filters, arbitrary query order -------------

def _rand_csr(rng, n_docs, vocab, max_nnz):
    indptr, tok, w = [0], [], []
    for d in range(n_docs):
        nnz = int(rng.integers(0, max_nnz))
        t = np.sort(rng.choice(vocab, size=min(nnz, vocab), replace=False))
        tok += t.tolist(); w += rng.random(t.shape[0]).astype(f32).tolist(); indptr.append(len(tok))
    return np.asarray(indptr, np.uint64), np.asarray(tok, np.uint32), np.asarray(w, f32)


@pytest.mark.parametrize("n_docs,vocab,max_nnz", [(300, 50, 12), (5000, 200, 30), (9000, 30522, 60)])
def test_sparse_random_vs_oracle(cqs, n_docs, vocab, max_nnz):
    rng = np.random.default_rng(n_docs)
    indptr, tok, w = _rand_csr(rng, n_docs, vocab, max_nnz)
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(n_docs, 8, seed=1)); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    for trial in range(8):
        qn = int(rng.integers(1, 40))
        qt = rng.choice(vocab + 5, size=qn, replace=False).astype(np.uint32)   # unsorted, some out of vocab
        qw = rng.random(qn).astype(f32)
        mask = rng.random(n_docs) > 0.3 if trial % 2 else None
        bs = None if mask is None else O.mask_to_bitset(mask)
        for k in (1, 20, 500, 1024):
            o_rows, o_sc = O.sparse_search_csr(indptr, tok, w, qt, qw, n_docs, k, mask)
            g_rows, g_sc = ix.search_sparse_rows(qt, qw, k, bs)
            assert g_rows.astype(np.int64).tolist() == o_rows.tolist()
            assert np.array_equal(bits(g_sc), bits(o_sc))
    ix.close()


def test_sparse_duplicate_token_in_doc_is_rejected(cqs):
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(2, 8, seed=1)); ix.finalize()
    with pytest.raises(cqs.B200Error):
        ix.sparse_attach(np.asarray([0, 2, 3], np.uint64), np.asarray([4, 4, 1], np.uint32),
                         np.asarray([.1, .2, .3], f32), 10)
    ix.close()


# ---- hybrid end to end: dense leg + sparse leg + fusion in one call ---------------------

@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_hybrid_matches_oracle_pipeline(cqs, storage):
    rng = np.random.default_rng(21)
    n, dim, vocab = 6000, 768, 3000
    rows = O.fast_unit_rows(n, dim, seed=21, clustered=True)
    indptr, tok, w = _rand_csr(rng, n, vocab, 50)
    ix = cqs.B200Index(dim, storage=storage)
    ix.append(None, rows); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    corpus = O.bf16_to_f32(O.f32_to_bf16_rne(rows)) if storage == "bf16" else rows
    for trial, cat in enumerate(O.CATEGORIES):
        alpha = O.resolve_splade_alpha(cat, env={})
        q = rows[int(rng.integers(0, n))] + rng.standard_normal(dim).astype(f32) * f32(0.02)
        q = (q / np.linalg.norm(q)).astype(f32)
        qn = int(rng.integers(1, 64))
        qt = np.sort(rng.choice(vocab, size=qn, replace=False)).astype(np.uint32)
        qw = rng.random(qn).astype(f32)
        mask = rng.random(n) > 0.4 if trial % 3 == 0 else None
        bs = None if mask is None else O.mask_to_bitset(mask)
        got = ix.search_hybrid_rows(q, qt, qw, float(alpha), 500, bs)
        # oracle: the same three steps, dense pool taken from the GPU's own dense
        # leg scores (dense parity is tested separately to 1e-5; fusion is bit-exact
        # GIVEN the pools)
        d_rows, d_sc = ix.search_rows(q, 500, bs)
        s_rows, s_sc = O.sparse_search_csr(indptr, tok, w, qt, qw, n, 500, mask)
        want = O.fuse_hybrid(list(zip(d_rows.tolist(), d_sc.tolist())),
                             list(zip(s_rows.tolist(), s_sc.tolist())), alpha, 500)
        assert got["rows"].tolist() == [x["id"] for x in want]
        assert np.array_equal(bits(got["fused"]), bits([x["fused"] for x in want]))
        assert np.array_equal(bits(got["sparse_raw"]), bits([x["sparse_raw"] for x in want]))
        # and the dense pool itself against the f64 oracle
        full = O.dense_scores(corpus, q)
        o_rows, o_sc = O.topk_rows(full, 500, mask)
        from tests.parity import assert_topk_parity
        assert_topk_parity(d_rows, d_sc, o_rows, o_sc, full)
    ix.close()


def test_hybrid_id_level_mirror_and_nonfinite_query(cqs):
    ids = [f"c{i:03d}" for i in range(64)]
    ix, emb = _mk_index(cqs, ids, dim=768, seed=3)
    sp = cqs.SpladeIndex(ix, [(ids[i], [(i % 7, 0.5 + i / 100), (10 + i % 3, 0.25)]) for i in range(0, 64, 2)],
                         vocab=64)
    out = cqs.search_hybrid(ix, emb[5], [(5, 1.0), (11, 0.5)], 0.6, 5)
    assert out and out[0]["id"] == "c005" or out[0]["in_dense"]
    assert all(out[i]["fused"] >= out[i + 1]["fused"] for i in range(len(out) - 1))
    only_even = lambda cid: int(cid[1:]) % 2 == 0
    out = cqs.search_hybrid(ix, emb[5], [(5, 1.0)], 0.6, 5, only_even)
    assert out and all(only_even(o["id"]) for o in out)
    # malformed dense query: dense leg empty, sparse leg still contributes (d = 0.0)
    q = emb[5].copy(); q[0] = np.nan
    out = cqs.search_hybrid(ix, q, [(5, 1.0)], 0.6, 5)
    assert out and all((not o["in_dense"]) and o["in_sparse"] and o["dense"] == 0.0 for o in out)
    ix.close()


# ---- centroid routing --------------------------------------------------------------------

def test_centroid_routing_bit_exact(cqs):
    rng = np.random.default_rng(9)
    dim = 768
    cents = {c: (v / np.linalg.norm(v)).astype(f32)
             for c, v in zip(cqs.CATEGORIES, rng.standard_normal((9, dim)))}
    clf = cqs.CentroidClassifier(cents, threshold=0.01)
    qs = rng.standard_normal((300, dim)).astype(f32)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    qs[:50] = np.stack([cents[cqs.CATEGORIES[i % 9]] for i in range(50)]) + qs[:50] * f32(0.05)
    cats, margin = clf.classify_batch(qs)
    o_cat, o_margin = O.centroid_classify(clf.matrix, qs, 0.01)
    assert [(-1 if c is None else clf.names.index(c)) for c in cats] == o_cat.tolist()
    assert np.array_equal(bits(margin), bits(o_margin))
    assert any(c is None for c in cats) and any(c is not None for c in cats)
    assert clf.classify(np.ones(5, f32)) is None                      # dim mismatch -> None
    # alpha with the centroid floor (src/cli/commands/search/query.rs:648-657)
    from cqs_b200.router import reclassify_with_centroid
    cat, applied = reclassify_with_centroid("unknown", qs[0], clf, env={})
    assert applied and cqs.apply_centroid_floor(cqs.resolve_splade_alpha(cat, env={}), applied) >= f32(0.7)
    cat, applied = reclassify_with_centroid("structural", qs[0], clf, env={})
    assert cat == "structural" and not applied


# ---- a14 rrf_fuse_n (src/search/scoring/fusion.rs:208-331) ---------------------------------

def test_rrf_fuse_reference_vectors_and_random(cqs):
    from cqs_b200.index import rrf_fuse_n
    ids, sc = rrf_fuse_n([[0, 1, 2, 3, 0, 4]], 10)                        # per-list dedup
    assert abs(float(sc[ids.tolist().index(0)]) - 1.0 / 61.0) < 1e-6
    ids, sc = rrf_fuse_n([[9, 1, 2], [9, 3], [9, 4]], 10)                  # cumulative overlap
    by = dict(zip(ids.tolist(), sc.tolist()))
    assert abs(by[9] - 3.0 / 61.0) < 1e-6 and abs(by[1] - 1.0 / 62.0) < 1e-6 and ids[0] == 9
    ids, sc = rrf_fuse_n([[0, 1], [2, 3], [4, 5], [6, 7]], 3)              # limit, ties -> id asc
    assert ids.tolist() == [0, 2, 4] and sc[0] >= sc[1] >= sc[2]
    assert rrf_fuse_n([], 10)[0].shape[0] == 0
    rng = np.random.default_rng(8)
    for trial in range(5):
        lists = [rng.choice(900, size=int(rng.integers(1, 500)), replace=True).tolist()
                 for _ in range(int(rng.integers(1, 4)))]
        limit = int(rng.integers(1, 300))
        want = O.rrf_fuse_n(lists, limit)
        ids, sc = rrf_fuse_n(lists, limit)
        assert ids.tolist() == [i for i, _ in want]
        assert np.array_equal(bits(sc), bits([s for _, s in want]))


# ---- inverted-index build on the device (sparse_build.cu) and its persistence ----------------

def _postings(ix, vocab, nnz):
    import ctypes as C
    from cqs_b200.capi import lib
    lib.cqs_b200_debug_sparse_postings.restype = C.c_int
    lib.cqs_b200_debug_sparse_postings.argtypes = [C.c_void_p] * 4
    tptr = np.zeros(vocab + 1, np.uint64); doc = np.zeros(max(nnz, 1), np.uint32); w = np.zeros(max(nnz, 1), f32)
    assert lib.cqs_b200_debug_sparse_postings(ix._h, tptr.ctypes.data_as(C.c_void_p), doc.ctypes.data_as(C.c_void_p),
                                              w.ctypes.data_as(C.c_void_p)) == 0
    return tptr, doc[:nnz], w[:nnz]


def _zipf_csr(rng, n_docs, vocab, mean_nnz):
    """Zipf-distributed tokens (a few very long posting lists, many same-token neighbours), some empty docs."""
    p = 1.0 / np.arange(1, vocab + 1) ** 1.1
    p /= p.sum()
    nnz_d = np.clip(rng.poisson(mean_nnz, n_docs), 0, min(vocab, 4 * mean_nnz))
    nnz_d[rng.random(n_docs) < 0.02] = 0
    indptr = np.zeros(n_docs + 1, np.uint64); indptr[1:] = np.cumsum(nnz_d)
    tok = np.empty(int(indptr[-1]), np.uint32)
    for d in range(n_docs):
        if nnz_d[d]:
            tok[int(indptr[d]):int(indptr[d + 1])] = np.sort(rng.choice(vocab, size=int(nnz_d[d]), replace=False, p=p))
    w = rng.random(tok.shape[0]).astype(f32) + f32(0.01)
    return indptr, tok, w


@pytest.mark.parametrize("n_docs,vocab,mean_nnz", [(1, 7, 3), (700, 97, 20), (40_000, 30522, 60), (3000, 56000, 40)])
def test_device_build_equals_stable_transposition(cqs, n_docs, vocab, mean_nnz):
    """SpladeIndex::build (src/splade/index.rs:177-221): token-major lists in ascending chunk order ==
    a stable sort of the doc-major entries by token id."""
    rng = np.random.default_rng(n_docs + vocab)
    indptr, tok, w = _zipf_csr(rng, n_docs, vocab, mean_nnz)
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(n_docs, 8, seed=1)); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    nnz = tok.shape[0]
    tptr, pdoc, pw = _postings(ix, vocab, nnz)
    order = np.argsort(tok, kind="stable")
    doc_of = np.repeat(np.arange(n_docs, dtype=np.uint32), np.diff(indptr.astype(np.int64)))
    exp_tptr = np.zeros(vocab + 1, np.uint64); exp_tptr[1:] = np.cumsum(np.bincount(tok, minlength=vocab))
    assert np.array_equal(tptr, exp_tptr)
    assert np.array_equal(pdoc, doc_of[order])
    assert np.array_equal(bits(pw), bits(w[order]))
    ix.close()


def test_device_build_rejects_bad_tokens_and_accepts_device_input(cqs):
    import torch
    rng = np.random.default_rng(3)
    n_docs, vocab = 2000, 500
    indptr, tok, w = _zipf_csr(rng, n_docs, vocab, 25)
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(n_docs, 8, seed=1)); ix.finalize()
    bad = tok.copy(); bad[len(bad) // 2] = vocab
    with pytest.raises(cqs.B200Error):
        ix.sparse_attach(indptr, bad, w, vocab)
    dev = torch.device("cuda", 0)
    d_ip = torch.from_numpy(indptr.view(np.int64)).to(dev)
    d_tok = torch.from_numpy(tok.view(np.int32)).to(dev)
    d_w = torch.from_numpy(w).to(dev)
    ix.sparse_attach_device(d_ip.data_ptr(), d_tok.data_ptr(), d_w.data_ptr(), tok.shape[0], vocab)
    a = _postings(ix, vocab, tok.shape[0])
    ix.sparse_attach(indptr, tok, w, vocab)
    b = _postings(ix, vocab, tok.shape[0])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ix.close()


def test_sparse_persistence_generation_and_checksum(cqs, tmp_path):
    """SpladeIndex::save/load contract (src/splade/index.rs:13-16, :308-560): stale generation,
    damaged body or a different chunk count -> load refuses and the caller rebuilds."""
    rng = np.random.default_rng(5)
    n_docs, vocab = 3000, 800
    indptr, tok, w = _zipf_csr(rng, n_docs, vocab, 30)
    rows = O.fast_unit_rows(n_docs, 8, seed=1)
    ix = cqs.B200Index(8)
    ix.append(None, rows); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    path = str(tmp_path / "splade.b200.bin")
    ix.sparse_save(path, generation=41)
    qt = np.asarray([0, 1, 5, 17, 300], np.uint32); qw = np.asarray([1.0, .5, .25, 2.0, .7], f32)
    want = ix.search_sparse_rows(qt, qw, 50)
    ix.close()
    ix2 = cqs.B200Index(8)
    ix2.append(None, rows); ix2.finalize()
    assert not ix2.sparse_load(path, expected_generation=42)            # sparse_vectors changed since the save
    assert not ix2.sparse_load(str(tmp_path / "missing.bin"), 41)
    assert ix2.sparse_load(path, expected_generation=41)
    got = ix2.search_sparse_rows(qt, qw, 50)
    assert np.array_equal(got[0], want[0]) and np.array_equal(bits(got[1]), bits(want[1]))
    ix3 = cqs.B200Index(8)
    ix3.append(None, rows[:-1]); ix3.finalize()
    assert not ix3.sparse_load(path, expected_generation=41)            # chunk count differs
    ix3.close()
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x40
    open(path, "wb").write(bytes(raw))
    assert not ix2.sparse_load(path, expected_generation=41)            # checksum
    got = ix2.search_sparse_rows(qt, qw, 50)                            # a refused load leaves the old index in place
    assert np.array_equal(got[0], want[0])
    ix2.close()


def test_sparse_load_survives_damaged_headers(cqs, tmp_path):
    """A damaged header must make the load fail cleanly (-> rebuild, src/splade/index.rs:308-560),
    never size a buffer from it: vocab / nnz fields are bounded and matched against the file size
    before anything is allocated."""
    import struct
    rng = np.random.default_rng(6)
    n_docs, vocab = 500, 300
    indptr, tok, w = _zipf_csr(rng, n_docs, vocab, 12)
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(n_docs, 8, seed=1)); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    path = str(tmp_path / "splade.bin")
    ix.sparse_save(path, generation=7)
    good = open(path, "rb").read()
    # header: magic[8] version u32 vocab u32 generation u64 n_docs u64 nnz u64 checksum u64 pad[16]
    for off, fmt, val in ((12, "<I", 0xFFFFFFFF), (12, "<I", 56321), (12, "<I", 0), (32, "<Q", 1 << 36),
                          (32, "<Q", (1 << 36) + 1), (32, "<Q", 0)):
        raw = bytearray(good)
        struct.pack_into(fmt, raw, off, val)
        open(path, "wb").write(bytes(raw))
        assert not ix.sparse_load(path, expected_generation=7)
    open(path, "wb").write(good + b"\0")
    assert not ix.sparse_load(path, expected_generation=7)              # trailing byte
    open(path, "wb").write(good[:40])
    assert not ix.sparse_load(path, expected_generation=7)              # short header
    open(path, "wb").write(good)
    assert ix.sparse_load(path, expected_generation=7)
    assert not ix.is_poisoned()
    ix.close()


@pytest.mark.parametrize("n_docs,vocab,mean_nnz", [(40_000, 30522, 60), (70_001, 600, 25)])
def test_sparse_search_with_static_block_index_vs_oracle(cqs, n_docs, vocab, mean_nnz):
    """Long posting lists take their block boundaries from the static index built at attach time,
    short ones from the per-query bounds pass; the scores stay bit-exact and the order identical."""
    rng = np.random.default_rng(n_docs)
    indptr, tok, w = _zipf_csr(rng, n_docs, vocab, mean_nnz)
    ix = cqs.B200Index(8)
    ix.append(None, O.fast_unit_rows(n_docs, 8, seed=1)); ix.finalize()
    ix.sparse_attach(indptr, tok, w, vocab)
    for trial in range(6):
        qn = int(rng.integers(1, 64))
        head = rng.choice(min(vocab, 40), size=min(qn, 8), replace=False)           # the heaviest tokens
        tail = rng.choice(vocab, size=qn, replace=False)
        qt = np.unique(np.concatenate([head, tail]))[:64].astype(np.uint32)
        rng.shuffle(qt)                                                               # query order matters for f32 sums
        qw = (rng.random(qt.shape[0]) + 0.05).astype(f32)
        mask = rng.random(n_docs) > 0.5 if trial % 2 else None
        bs = None if mask is None else O.mask_to_bitset(mask)
        for k in (20, 500):
            o_rows, o_sc = O.sparse_search_csr(indptr, tok, w, qt, qw, n_docs, k, mask)
            g_rows, g_sc = ix.search_sparse_rows(qt, qw, k, bs)
            assert g_rows.astype(np.int64).tolist() == o_rows.tolist()
            assert np.array_equal(bits(g_sc), bits(o_sc))
    ix.close()
