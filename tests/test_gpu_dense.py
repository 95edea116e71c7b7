"""GPU parity: dense exact scan + fused top-k (kernels 1 and 3) vs the oracle.

All calls go through the C ABI (ctypes) — cqs_b200.B200Index is the mirror of
the reference's VectorIndex (src/index.rs:139-239)."""
import numpy as np
import pytest

from oracle import cqs_oracle as O, c_oracle as CO
from tests.parity import assert_topk_parity, bits

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def cqs():
    import cqs_b200
    return cqs_b200


@pytest.fixture(scope="module")
def config1():
    """BASELINE configs[0]: 17,523 x 768 f32 from the reference's own generator
    (examples/exp_level_scale.rs:200-224), 218 queries: the generator stream
    continued, plus 32 'row + noise' self-match queries; 32 exact duplicate rows."""
    rows, st = CO.synth_vectors(17523, 768)
    rows[100:132] = rows[5000:5032]            # exact duplicates -> ties broken by row asc
    q, _ = CO.synth_vectors(218, 768, st)
    rng = np.random.default_rng(1)
    pick = rng.choice(17523, size=32, replace=False)
    noisy = rows[pick] + rng.standard_normal((32, 768)).astype(f32) * f32(0.002)
    noisy /= np.linalg.norm(noisy, axis=1, keepdims=True)
    q[:32] = noisy.astype(f32)
    q[32] = rows[5003]                          # exact self-match, hits the duplicate pair
    return rows, q


@pytest.fixture(scope="module")
def index1(cqs, config1):
    rows, _ = config1
    ix = cqs.B200Index(768, storage="f32")
    ix.append(None, rows)
    ix.finalize()
    yield ix
    ix.close()


def test_config1_top20_all_218_queries(index1, config1):
    rows, queries = config1
    assert len(index1) == 17523 and index1.dim() == 768 and index1.name() == "B200"
    assert index1.index_scores_are_cosine() and index1.max_k() == 1024 and not index1.is_poisoned()
    for qi in range(queries.shape[0]):
        full = O.dense_scores(rows, queries[qi])
        o_rows, o_sc = O.topk_rows(full, 20)
        g_rows, g_sc = index1.search_rows(queries[qi], 20)
        assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)


def test_config1_duplicates_tie_break_is_row_ascending(index1, config1):
    rows, queries = config1
    g_rows, g_sc = index1.search_rows(queries[32], 4)   # query == rows[5003] == rows[103]
    assert g_rows[0] == 103 and g_rows[1] == 5003
    assert bits(g_sc[0]) == bits(g_sc[1])


def test_config1_production_pool_k500_and_k1024(index1, config1):
    rows, queries = config1
    for qi in (0, 33, 100):
        full = O.dense_scores(rows, queries[qi])
        for k in (500, 1024):
            o_rows, o_sc = O.topk_rows(full, k)
            g_rows, g_sc = index1.search_rows(queries[qi], k)
            assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)


@pytest.mark.parametrize("k", [1, 2, 7, 32, 33, 100])
def test_config1_various_k(index1, config1, k):
    rows, queries = config1
    full = O.dense_scores(rows, queries[40])
    o_rows, o_sc = O.topk_rows(full, k)
    g_rows, g_sc = index1.search_rows(queries[40], k)
    assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)


def test_filter_bitset_semantics(index1, config1):
    rows, queries = config1
    n = rows.shape[0]
    rng = np.random.default_rng(2)
    full = O.dense_scores(rows, queries[50])
    for frac in (0.5, 0.01):
        mask = rng.random(n) < frac
        bs = O.mask_to_bitset(mask)
        o_rows, o_sc = O.topk_rows(full, 20, mask)
        g_rows, g_sc = index1.search_rows(queries[50], 20, bs)
        assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
        assert mask[g_rows.astype(np.int64)].all()          # only passing ids
    one = np.zeros(n, bool); one[n - 1] = True               # last row only, k > included
    g_rows, g_sc = index1.search_rows(queries[50], 20, O.mask_to_bitset(one))
    assert g_rows.tolist() == [n - 1]
    none = np.zeros(n, bool)
    g_rows, _ = index1.search_rows(queries[50], 20, O.mask_to_bitset(none))
    assert g_rows.shape[0] == 0


def test_query_guards_and_k_edges(cqs, index1, config1):
    rows, queries = config1
    q = queries[0]
    assert index1.search(q, 0) == []                                      # k = 0 (src/cagra.rs:445)
    assert index1.search_rows(q[:100], 5)[0].shape[0] == 0                # wrong dim -> empty
    for bad in (np.nan, np.inf, -np.inf):                                 # tests/hnsw_test.rs:460-552
        qq = q.copy(); qq[7] = bad
        assert index1.search_rows(qq, 5)[0].shape[0] == 0
    with pytest.raises(cqs.B200Error):                                    # k > max_k is refused loudly...
        index1.search_rows(q, 1025)
    assert index1.search(q, 1025) == []                                   # ...and the trait method maps it to empty
    assert not index1.is_poisoned()
    # idempotent
    a = index1.search_rows(q, 20); b = index1.search_rows(q, 20)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))


@pytest.mark.parametrize("n", [1, 2, 31, 33, 127, 1000])
def test_small_corpora_k_larger_than_n(cqs, n):
    rows = O.fast_unit_rows(n, 768, seed=n)
    ix = cqs.B200Index(768)
    ix.append(None, rows); ix.finalize()
    q = O.fast_unit_rows(1, 768, seed=99)[0]
    g_rows, g_sc = ix.search_rows(q, 50)
    full = O.dense_scores(rows, q)
    o_rows, o_sc = O.topk_rows(full, 50)
    assert g_rows.shape[0] == min(n, 50)                                  # <= n unique real ids
    assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
    ix.close()


def test_empty_index_returns_empty(cqs):
    ix = cqs.B200Index(768)
    ix.finalize()
    assert len(ix) == 0 and ix.is_empty()
    assert ix.search(np.ones(768, f32), 5) == []
    ix.close()


def test_zero_and_nonfinite_rows_in_corpus(cqs):
    rows = O.fast_unit_rows(500, 768, seed=4)
    rows[10] = 0.0                       # zero vector: finite score 0.0 (src/cagra.rs:2606-2640)
    rows[20, 5] = np.nan                 # NaN row: cosine_similarity -> None -> never a candidate
    rows[30, 6] = np.inf
    ix = cqs.B200Index(768)
    ix.append(None, rows); ix.finalize()
    q = O.fast_unit_rows(1, 768, seed=5)[0]
    g_rows, g_sc = ix.search_rows(q, 500)
    assert g_rows.shape[0] == 498 and np.all(np.isfinite(g_sc))
    assert 20 not in g_rows.tolist() and 30 not in g_rows.tolist() and 10 in g_rows.tolist()
    full = O.dense_scores(rows, q)
    o_rows, o_sc = O.topk_rows(full, 500)
    assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
    ix.close()


@pytest.mark.parametrize("dim", [768, 1024, 384, 100, 1536, 2048, 8])
@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_dims_and_storage(cqs, dim, storage):
    n = 3000
    rows = O.fast_unit_rows(n, dim, seed=dim)
    ix = cqs.B200Index(dim, storage=storage)
    ix.append(None, rows[:1700]); ix.append(None, rows[1700:])            # two appends (regrowth path)
    ix.finalize()
    corpus = O.bf16_to_f32(O.f32_to_bf16_rne(rows)) if storage == "bf16" else rows
    assert ix.index_scores_are_cosine() == (storage == "f32")
    for s in range(3):
        q = O.fast_unit_rows(1, dim, seed=1000 + s)[0]
        full = O.dense_scores(corpus, q)
        o_rows, o_sc = O.topk_rows(full, 20)
        g_rows, g_sc = ix.search_rows(q, 20)
        assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
    ix.close()


def test_clustered_corpus_near_ties_bf16_and_f32(cqs):
    rows = O.fast_unit_rows(20000, 768, seed=8, clustered=True)
    q = rows[123] + np.random.default_rng(0).standard_normal(768).astype(f32) * f32(0.01)
    q /= np.linalg.norm(q)
    for storage in ("f32", "bf16"):
        ix = cqs.B200Index(768, storage=storage)
        ix.append(None, rows); ix.finalize()
        corpus = O.bf16_to_f32(O.f32_to_bf16_rne(rows)) if storage == "bf16" else rows
        full = O.dense_scores(corpus, q)
        for k in (20, 500):
            o_rows, o_sc = O.topk_rows(full, k)
            g_rows, g_sc = ix.search_rows(q, k)
            assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
        ix.close()


def test_vector_index_trait_mirror_ids_filter_and_order(cqs):
    """B200Index::{build, search, search_with_filter} at the id level: the feed
    arrives in rowid order, ids tie-break ascending (candidate.rs:321-329)."""
    rng = np.random.default_rng(12)
    n = 400
    emb = O.fast_unit_rows(n, 768, seed=12)
    ids = [f"f{int(x):05d}.rs:1:{int(x) * 7919 % 100000:08x}" for x in rng.permutation(n)]
    emb[7] = emb[3]                                  # tie between ids[3] and ids[7]
    emb[50] = 0.0                                    # dropped at build like prepare_index_data
    ix = cqs.B200Index.build(ids, emb)
    assert len(ix) == n - 1 and ix.id_map == sorted(i for j, i in enumerate(ids) if j != 50)
    res = ix.search(emb[3], 5)
    assert [r.id for r in res[:2]] == sorted([ids[3], ids[7]])
    assert all(res[i].score >= res[i + 1].score for i in range(len(res) - 1))
    # literal oracle over (id, score): BoundedScoreHeap semantics
    h = O.BoundedScoreHeap(5)
    for j, cid in enumerate(ids):
        if j == 50:
            continue
        h.push(cid, float(O.cosine_similarity(emb[3], emb[j])))
    assert [r.id for r in res] == [i for i, _ in h.into_sorted_vec()]
    only_rs = lambda cid: cid.endswith("0") or cid.endswith("a")
    res = ix.search_with_filter(emb[3], 10, only_rs)
    assert res and all(only_rs(r.id) for r in res)
    assert ix.search_with_filter(emb[3], 10, lambda cid: False) == []     # all-reject -> empty
    assert [r.id for r in ix.search_with_filter(emb[3], 5, lambda cid: True)] == [r.id for r in ix.search(emb[3], 5)]
    ix.close()


def test_search_batch_equals_single_calls(index1, config1):
    rows, queries = config1
    r, s, n = index1.search_batch_rows(queries[:16], 20)
    for i in range(16):
        a, b = index1.search_rows(queries[i], 20)
        assert n[i] == 20 and np.array_equal(r[i], a) and np.array_equal(bits(s[i]), bits(b))


def test_config2_full_size_1m_f32(cqs):
    """BASELINE configs[1] at full size: 1,000,000 x 768 f32, top-20 and top-500;
    oracle on 3 queries plus size-independent properties (self-match, idempotence,
    sortedness, filter subset)."""
    n, dim = 1_000_000, 768
    ix = cqs.B200Index(dim)
    ix.reserve(n)
    blocks = []
    for b in range(10):
        blk = O.fast_unit_rows(n // 10, dim, seed=100 + b)
        ix.append(None, blk)
        blocks.append(blk)
    ix.finalize()
    ix.set_timing(True)
    rows = np.concatenate(blocks)
    del blocks
    assert len(ix) == n
    for s in range(3):
        q = O.fast_unit_rows(1, dim, seed=2000 + s)[0]
        full = O.dense_scores(rows, q)
        for k in (20, 500):
            o_rows, o_sc = O.topk_rows(full, k)
            g_rows, g_sc = ix.search_rows(q, k)
            assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
    for i in (0, 123_456, n - 1):                        # self-match
        g_rows, g_sc = ix.search_rows(rows[i], 1)
        assert g_rows[0] == i and abs(g_sc[0] - 1.0) < 1e-5
    mask = np.zeros(n, bool); mask[::7] = True
    g_rows, _ = ix.search_rows(rows[77], 20, O.mask_to_bitset(mask))
    assert (g_rows % 7 == 0).all() and g_rows[0] == 77
    assert ix.last_kernel_ms() > 0
    ix.close()


def test_search_filtered_on_device_matches_reference_semantics(cqs):
    """Store::search_filtered (src/search/query.rs:348-510) with default signals: type/language
    filter, clamp -> note boost -> demotion -> threshold, all before the heap."""
    rng = np.random.default_rng(17)
    n, dim = 30_000, 768
    rows = O.fast_unit_rows(n, dim, seed=17, clustered=True)
    ctype = rng.integers(0, 12, n).astype(np.uint8)
    lang = rng.integers(0, 70, n).astype(np.uint8)          # > 64 codes: second mask word
    note = np.ones(n, f32); note[rng.random(n) < 0.05] = f32(1.15); note[rng.random(n) < 0.05] = f32(0.85)
    imp = np.ones(n, f32); imp[rng.random(n) < 0.2] = f32(0.70); imp[rng.random(n) < 0.1] = f32(0.80)
    ix = cqs.B200Index(dim)
    ix.append(None, rows); ix.finalize()
    ix.set_row_meta(ctype, lang)
    ix.set_row_signals(note, imp)
    for trial in range(6):
        q = rows[int(rng.integers(0, n))] + rng.standard_normal(dim).astype(f32) * f32(0.03)
        q = (q / np.linalg.norm(q)).astype(f32)
        types = None if trial % 3 == 0 else [1, 4, 7]
        langs = None if trial % 2 == 0 else [0, 3, 65, 69]
        thr = [0.0, 0.3, 0.5][trial % 3]
        demote = trial != 4
        for limit in (5, 20, 100):
            g_rows, g_sc = ix.search_filtered_rows(q, limit, thr, types, langs, demote)
            o_rows, o_sc = O.search_filtered(rows, q, limit, thr, ctype, lang, types, langs, note, imp, demote)
            assert g_rows.shape[0] == o_rows.shape[0]
            assert np.allclose(g_sc, o_sc, rtol=2e-5, atol=1e-7)
            if not np.array_equal(g_rows.astype(np.int64), o_rows):      # only near-ties may differ
                diff = np.nonzero(g_rows.astype(np.int64) != o_rows)[0]
                for i in diff:
                    assert abs(float(g_sc[i]) - float(o_sc[i])) <= 2e-5 * abs(float(o_sc[i])) + 1e-7
            if types is not None:
                assert np.isin(ctype[g_rows.astype(np.int64)], types).all()
            if langs is not None:
                assert np.isin(lang[g_rows.astype(np.int64)], langs).all()
            assert (g_sc >= thr).all()
    # plain search is unaffected by the uploaded signals
    a, b = ix.search_rows(q, 10)
    full = O.dense_scores(rows, q)
    o_rows, o_sc = O.topk_rows(full, 10)
    assert_topk_parity(a, b, o_rows, o_sc, full)
    ix.close()


def test_search_typed_equals_bitset_filter(cqs):
    rng = np.random.default_rng(23)
    n, dim = 20_000, 768
    rows = O.fast_unit_rows(n, dim, seed=23)
    ctype = rng.integers(0, 9, n).astype(np.uint8)
    lang = rng.integers(0, 5, n).astype(np.uint8)
    ix = cqs.B200Index(dim)
    ix.append(None, rows); ix.finalize()
    with pytest.raises(cqs.B200Error):
        ix.search_typed_rows(rows[0], 5, include_types=[1])              # meta not uploaded yet
    ix.set_row_meta(ctype, lang)
    q = O.fast_unit_rows(1, dim, seed=24)[0]
    for types, langs in (([1, 2], None), (None, [3]), ([0, 8], [1, 4]), (None, None)):
        mask = np.ones(n, bool)
        if types is not None: mask &= np.isin(ctype, types)
        if langs is not None: mask &= np.isin(lang, langs)
        a = ix.search_typed_rows(q, 50, types, langs)
        b = ix.search_rows(q, 50, O.mask_to_bitset(mask))
        assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    ix.close()


def test_extend_after_reopen(cqs):
    """TieredIndex::extend analogue (src/tiered.rs:317-360): reopen, append, finalize."""
    from cqs_b200.capi import lib, check
    rows = O.fast_unit_rows(5000, 768, seed=33)
    ix = cqs.B200Index(768)
    ix.append(None, rows[:3000]); ix.finalize()
    with pytest.raises(cqs.B200Error):
        ix.append(None, rows[3000:])                                      # sealed
    check(lib.cqs_b200_reopen(ix._h))
    ix.append(None, rows[3000:]); ix.finalize()
    q = rows[4500]
    g_rows, g_sc = ix.search_rows(q, 10)
    full = O.dense_scores(rows, q)
    o_rows, o_sc = O.topk_rows(full, 10)
    assert g_rows[0] == 4500
    assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
    ix.close()


def test_bf16_shadow_scan_with_f32_rescoring_equals_f32_storage(cqs, config1):
    """STORAGE_BF16_F32: single queries stream the bf16 shadow (2 B/elem), over-fetch k' candidates,
    re-score them on the f32 master rows inside the kernel tail and PROVE the pool complete (measured
    rounding distance) — or are re-run on the f32 rows.  Either way the answer must be bit-identical
    to STORAGE_F32 (same rows, same scores), for every k, with a filter, on self-matches, exact
    duplicates (ties by row) and on near-duplicate clusters that defeat the proof."""
    import ctypes as C
    from cqs_b200.capi import lib
    lib.cqs_b200_debug_shadow_reruns.restype = C.c_uint64
    lib.cqs_b200_debug_shadow_reruns.argtypes = [C.c_void_p]
    rows, queries = config1
    rows = rows.copy()
    rng = np.random.default_rng(9)
    base = rows[9000].copy()
    for j in range(6000):                                  # 6000 rows within ~2e-5 of each other: more than the
        v = base + rng.standard_normal(768).astype(f32) * f32(2e-5)   # 148 CTAs x 32 candidates a k <= 24 scan re-scores,
        rows[9001 + j] = v / np.linalg.norm(v)             # so rows it did NOT re-score tie with the answer: no proof
    a = cqs.B200Index(768, storage="f32")
    a.append(None, rows); a.finalize()
    b = cqs.B200Index(768, storage="bf16+f32")
    b.append(None, rows); b.finalize()
    mask = rng.random(rows.shape[0]) < 0.4
    bs = O.mask_to_bitset(mask)
    qs = [queries[i] for i in (0, 5, 32, 40, 100, 217)] + [base, rows[9001]]
    for qi, q in enumerate(qs):
        for k in (1, 20, 24, 25, 100, 500, 700, 1024):
            ra, sa = a.search_rows(q, k)
            rb, sb = b.search_rows(q, k)
            assert np.array_equal(ra, rb), (qi, k)
            assert np.array_equal(bits(sa), bits(sb)), (qi, k)
        ra, sa = a.search_rows(q, 20, bs)
        rb, sb = b.search_rows(q, 20, bs)
        assert np.array_equal(ra, rb) and np.array_equal(bits(sa), bits(sb))
    reruns = lib.cqs_b200_debug_shadow_reruns(b._h)
    assert reruns >= 2, reruns                              # the near-duplicate queries fell back to the f32 rows
    assert reruns < 50, reruns                              # ... and most of the 72 searches were proven on the shadow
    # batch entry (exact lanes for < 8 queries, tensor cores otherwise): same answers
    q8 = np.stack(qs[:8])
    for nq in (3, 8):
        r, sc, nn = b.search_batch_rows(q8[:nq], 20)
        for i in range(nq):
            ra, sa = a.search_rows(q8[i], 20)
            assert int(nn[i]) == 20 and np.array_equal(r[i], ra) and np.array_equal(bits(sc[i]), bits(sa)), (nq, i)
    assert b.index_scores_are_cosine()
    a.close(); b.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_extend_after_build_and_after_load(cqs, tmp_path, storage):
    """INTEGRATION.md §6 (watch loop): an index that was built (reserve + append) or loaded from
    disk is extended in place — reopen, append, finalize — and then answers over all rows.
    `cqs_b200_reserve` is a capacity hint on one device, not a cap."""
    from cqs_b200.capi import lib, check
    n0, n1, dim = 3000, 1777, 768
    emb = O.fast_unit_rows(n0 + n1, dim, seed=71)
    ids = [f"c{i:06d}" for i in range(n0 + n1)]
    corpus = emb if storage == "f32" else O.bf16_to_f32(O.f32_to_bf16_rne(emb))
    built = cqs.B200Index.build(ids[:n0], emb[:n0], storage=storage)          # calls reserve(n0)
    path = str(tmp_path / "ix.b200")
    built.save(path)
    loaded = cqs.B200Index.load(path)                                         # load reserves n0 too
    assert loaded is not None
    for ix in (built, loaded):
        check(lib.cqs_b200_reopen(ix._h))
        ix.append(ids[n0:], emb[n0:])
        ix.finalize()
        assert len(ix) == n0 + n1
        for qi in (n0 + 5, 17):
            q = emb[qi]
            g_rows, g_sc = ix.search_rows(q, 10)
            full = O.dense_scores(corpus, q)
            o_rows, o_sc = O.topk_rows(full, 10)
            assert g_rows[0] == qi
            assert_topk_parity(g_rows, g_sc, o_rows, o_sc, full)
        ix.close()


def test_index_score_reuse_matches_recompute(cqs, config1, index1):
    """src/search/query.rs:2061-2170 (test_index_score_reuse_matches_recompute): a backend may only
    answer index_scores_are_cosine() == true if the scores it returns equal a recomputed
    cosine_similarity(query, stored embedding) to < 1e-6 — that is what lets
    search_filtered_with_index skip the BLOB re-fetch (query.rs:1152-1172).  Checked for f32
    storage on self-matches (score ~ 1, the worst case for an absolute bound), near-duplicates
    and ordinary queries, k = 500 (the production pool)."""
    rows, queries = config1
    assert index1.index_scores_are_cosine()
    worst = 0.0
    for qi in list(range(0, 40)) + [100, 150, 217]:
        q = queries[qi]
        g_rows, g_sc = index1.search_rows(q, 500)
        recomputed = np.asarray([O.cosine_similarity(q, rows[int(r)]) for r in g_rows], f32)
        worst = max(worst, float(np.max(np.abs(g_sc.astype(np.float64) - recomputed.astype(np.float64)))))
    assert worst < 1e-6, worst
    # bf16-only storage scores the ROUNDED rows: not the f32 cosine, so it must say so
    ixb = cqs.B200Index(768, storage="bf16")
    ixb.append(None, rows[:2000]); ixb.finalize()
    assert not ixb.index_scores_are_cosine()
    ixb.close()
    # bf16 + f32 master: results come from the f32 rows
    ixm = cqs.B200Index(768, storage="bf16+f32")
    ixm.append(None, rows[:2000]); ixm.finalize()
    assert ixm.index_scores_are_cosine()
    g_rows, g_sc = ixm.search_rows(queries[3], 50)
    rec = np.asarray([O.cosine_similarity(queries[3], rows[int(r)]) for r in g_rows], f32)
    assert np.max(np.abs(g_sc.astype(np.float64) - rec.astype(np.float64))) < 1e-6
    ixm.close()


@pytest.mark.parametrize("storage", ["f32", "bf16", "bf16+f32"])
def test_save_load_roundtrip_and_corruption(cqs, tmp_path, storage):
    """Persistence conventions of src/cagra.rs:963-1652: checksummed blob + id sidecar,
    atomic write; any mismatch -> None -> rebuild."""
    n, dim = 4000, 768
    emb = O.fast_unit_rows(n, dim, seed=44)
    ids = [f"chunk_{i:05d}" for i in range(n)]
    ix = cqs.B200Index.build(ids, emb, storage=storage)
    path = str(tmp_path / "index.b200")
    ix.save(path)
    ix2 = cqs.B200Index.load(path)
    assert ix2 is not None and len(ix2) == n and ix2.id_map == ix.id_map
    for s in range(3):
        q = O.fast_unit_rows(1, dim, seed=500 + s)[0]
        a, b = ix.search_rows(q, 20)
        c, d = ix2.search_rows(q, 20)
        assert np.array_equal(a, c) and np.array_equal(bits(b), bits(d))
    assert [r.id for r in ix2.search(emb[7], 3)][0] == "chunk_00007"
    ix2.close()
    raw = bytearray(open(path, "rb").read())
    raw[64 + 12345] ^= 0x40                                     # flip one payload bit
    open(path, "wb").write(bytes(raw))
    assert cqs.B200Index.load(path) is None                     # checksum mismatch
    open(path, "wb").write(bytes(raw[: len(raw) // 2]))
    assert cqs.B200Index.load(path) is None                     # truncated
    assert cqs.B200Index.load(str(tmp_path / "missing.b200")) is None
    ix.close()
