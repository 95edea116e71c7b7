"""Pin the CPU oracle against the reference's own unit-test vectors.

Each test names the reference test (file:line) whose values it reproduces.
"""
import math

import numpy as np
import pytest

from oracle import cqs_oracle as O

f32 = np.float32


# --- a1: src/math.rs:95-215 --------------------------------------------------

def test_cosine_identical_orthogonal_symmetry():
    v = np.zeros(768, f32); v[0] = 1.0
    w = np.zeros(768, f32); w[1] = 1.0
    assert O.cosine_similarity(v, v) > 0.99          # math.rs:95-104
    assert abs(O.cosine_similarity(v, w)) < 0.01      # math.rs:106-116
    rng = np.random.default_rng(0)
    a = rng.standard_normal(768).astype(f32); a /= np.linalg.norm(a)
    b = rng.standard_normal(768).astype(f32); b /= np.linalg.norm(b)
    assert abs(O.cosine_similarity(a, b) - O.cosine_similarity(b, a)) < 1e-6  # symmetry


def test_cosine_len_mismatch_empty_nonfinite_zero():
    assert O.cosine_similarity(np.ones(3, f32), np.ones(4, f32)) is None   # math.rs:118-126
    assert O.cosine_similarity(np.zeros(0, f32), np.zeros(0, f32)) is None
    a = np.ones(768, f32); b = np.ones(768, f32)
    b[5] = np.nan
    assert O.cosine_similarity(a, b) is None                                # cosine_nan_embedding
    b[5] = np.inf
    assert O.cosine_similarity(a, b) is None                                # cosine_inf_embedding
    z = np.zeros(768, f32)
    s = O.cosine_similarity(z, a)                                           # cosine_zero_norm_vector
    assert s is not None and s == 0.0


# --- a4: src/search/scoring/candidate.rs:588-707, :1426-1490 -------------------

def _heap(cap, pushes):
    h = O.BoundedScoreHeap(cap)
    for i, s in pushes:
        h.push(i, s)
    return h.into_sorted_vec()


def test_bounded_heap_equal_scores_and_push_order():
    assert [i for i, _ in _heap(2, [("a", .5), ("b", .5), ("c", .5)])] == ["a", "b"]
    assert [i for i, _ in _heap(2, [("c", .5), ("b", .5), ("a", .5)])] == ["a", "b"]


def test_bounded_heap_evicts_lowest():
    assert [i for i, _ in _heap(2, [("low", .1), ("mid", .5), ("high", .9)])] == ["high", "mid"]


def test_bounded_heap_non_finite_and_empty_and_zero_capacity():
    r = _heap(5, [("nan", math.nan), ("inf", math.inf), ("neginf", -math.inf), ("ok", .5)])
    assert [i for i, _ in r] == ["ok"]
    assert _heap(5, []) == []
    assert _heap(5, [("a", math.nan), ("b", math.nan), ("c", math.nan)]) == []
    assert _heap(0, [("a", .9), ("b", .8)]) == []
    r = _heap(10, [("nan1", math.nan), ("ok1", .7), ("inf", math.inf), ("ok2", .9),
                   ("nan2", math.nan), ("neginf", -math.inf), ("ok3", .5)])
    assert [i for i, _ in r] == ["ok2", "ok1", "ok3"]


def test_bounded_heap_negative_scores():
    assert [i for i, _ in _heap(5, [("a", -.1), ("b", -.5), ("c", -.3)])] == ["a", "c", "b"]


def test_would_accept_below_and_at_capacity():
    h = O.BoundedScoreHeap(2)
    assert h.would_accept(0.1)
    h.push("a", 0.5)
    assert h.would_accept(-1.0)
    assert not h.would_accept(math.nan)
    h.push("b", 0.9)
    assert h.would_accept(0.7)
    assert not h.would_accept(0.1)
    assert h.would_accept(0.5)
    assert not O.BoundedScoreHeap(0).would_accept(1.0)


def test_topk_rows_matches_literal_heap_on_random_ties():
    rng = np.random.default_rng(3)
    for trial in range(30):
        n = int(rng.integers(1, 200))
        k = int(rng.integers(0, 40))
        sc = rng.choice(np.array([0.1, 0.2, 0.2, 0.5, -0.0, 0.0, np.nan, np.inf, -0.3], f32), size=n)
        mask = rng.random(n) > 0.2
        h = O.BoundedScoreHeap(k)
        for i in rng.permutation(n):
            if mask[i]:
                h.push(int(i), float(sc[i]))
        want = h.into_sorted_vec()
        rows, scores = O.topk_rows(sc, k, mask)
        assert [int(r) for r in rows] == [i for i, _ in want]
        assert [float(s) for s in scores] == [float(f32(s)) for _, s in want]


# --- a5: candidate.rs:1674-1802 -------------------------------------------------

def test_apply_scoring_pipeline_pinned_exact_scores():
    # A: ((.7*.62 + .3*1.0).max(0) * 1.15) * 1.0
    nb = f32(0.3)
    base = f32(f32(f32(1.0) - nb) * f32(0.62)) + f32(nb * f32(1.0))
    exp_a = f32(f32(max(f32(base), f32(0))) * f32(f32(1.0) + f32(f32(1.0) * f32(0.15)))) * f32(1.0)
    got_a = O.apply_scoring_pipeline(0.62, name_boost=0.3, name_score=1.0,
                                     note_boost=f32(f32(1.0) + f32(0.15)), importance=1.0)
    assert got_a.view(np.uint32) == f32(exp_a).view(np.uint32)
    # B: .81 * .70
    got_b = O.apply_scoring_pipeline(0.81, importance=0.70)
    assert got_b.view(np.uint32) == f32(f32(f32(0.81) * f32(1.0)) * f32(0.70)).view(np.uint32)
    # C: glob mismatch
    assert O.apply_scoring_pipeline(0.9, glob_ok=False) is None
    # D: threshold inclusive
    assert O.apply_scoring_pipeline(0.75, threshold=0.75) == f32(0.75)
    assert O.apply_scoring_pipeline(0.7499999, threshold=0.75) is None
    # E: negative base clamps to 0
    got_e = O.apply_scoring_pipeline(-0.4, note_boost=f32(f32(1.0) + f32(-0.15)))
    assert got_e == 0.0
    # F: > 1 clamps to 1
    assert O.apply_scoring_pipeline(1.05) == f32(1.0)


# --- a10: src/splade/index.rs:1114-1241 -----------------------------------------

def _splade_fixture():
    return O.SpladeIndex([
        ("chunk_a", [(1, .5), (2, .3), (3, .8)]),
        ("chunk_b", [(1, .7), (4, .6)]),
        ("chunk_c", [(2, .9), (3, .1), (5, .4)]),
    ])


def test_splade_build_and_search():
    ix = _splade_fixture()
    assert len(ix) == 3
    r = ix.search([(1, 1.0)], 10)
    assert [i for i, _ in r] == ["chunk_b", "chunk_a"]


def test_splade_dot_product_correct():
    r = _splade_fixture().search([(1, 1.0), (2, 1.0)], 10)
    assert [i for i, _ in r] == ["chunk_c", "chunk_a", "chunk_b"]
    for (_, s), want in zip(r, (0.9, 0.8, 0.7)):
        assert abs(float(s) - want) < 1e-5


def test_splade_filter_nomatch_empty_k():
    ix = _splade_fixture()
    r = ix.search_with_filter([(1, 1.0)], 10, lambda i: i == "chunk_a")
    assert [i for i, _ in r] == ["chunk_a"]
    assert ix.search([(999, 1.0)], 10) == []
    assert ix.search([], 10) == []
    assert len(ix.search([(1, 1.0), (2, 1.0), (3, 1.0)], 2)) == 2
    assert ix.search([(1, 1.0), (2, 1.0), (3, 1.0)], 0) == []
    assert O.SpladeIndex([]).search([(1, 1.0)], 10) == []


def test_splade_nan_inf_query_weights_do_not_crash():
    ix = _splade_fixture()
    for w in (math.nan, math.inf):
        for i, _ in ix.search([(1, w)], 10):
            assert i in ("chunk_a", "chunk_b")


def test_sparse_csr_matches_literal_index():
    rng = np.random.default_rng(11)
    n_docs, vocab = 300, 50
    chunks, indptr, tok, w = [], [0], [], []
    for d in range(n_docs):
        nnz = int(rng.integers(0, 12))
        t = np.sort(rng.choice(vocab, size=nnz, replace=False))
        ww = rng.random(nnz).astype(f32)
        chunks.append((d, list(zip(t.tolist(), ww.tolist()))))
        tok += t.tolist(); w += ww.tolist(); indptr.append(len(tok))
    ix = O.SpladeIndex(chunks)
    for trial in range(10):
        qn = int(rng.integers(1, 10))
        qt = rng.choice(vocab + 5, size=qn, replace=False)  # arbitrary (unsorted) order
        qw = rng.random(qn).astype(f32)
        mask = rng.random(n_docs) > 0.3
        want = ix.search_with_filter(list(zip(qt.tolist(), qw.tolist())), 25, lambda i: bool(mask[i]))
        rows, sc = O.sparse_search_csr(indptr, tok, w, qt, qw, n_docs, 25, mask)
        assert [int(r) for r in rows] == [i for i, _ in want]
        assert [s.view(np.uint32) for s in sc] == [f32(s).view(np.uint32) for _, s in want]


# --- a11: tests/search_test.rs:549-791 (legs_corpus) ------------------------------

def test_alpha_fusion_legs_corpus():
    dense = [("A", 0.92), ("B", 0.61)]
    sparse = O.SpladeIndex([("B", [(1, 0.8)]), ("C", [(1, 0.5)])]).search([(1, 1.0)], 500)
    assert [i for i, _ in sparse] == ["B", "C"]
    assert abs(float(sparse[0][1]) - 0.8) < 1e-6 and abs(float(sparse[1][1]) - 0.5) < 1e-6
    fused = O.fuse_hybrid(dense, sparse, 0.5, 500)
    by = {r["id"]: r for r in fused}
    # raw cosines exact, min-max B 1.0 / C .625, absent legs 0.0
    assert by["A"]["dense"] == f32(0.92) and by["B"]["dense"] == f32(0.61)
    assert by["B"]["sparse_norm"] == f32(1.0)
    assert abs(float(by["C"]["sparse_norm"]) - 0.625) < 1e-6
    assert by["A"]["sparse_norm"] == 0.0 and not by["A"]["in_sparse"]
    assert by["C"]["dense"] == 0.0 and not by["C"]["in_dense"]
    # derived: A .46, B .805, C .3125 -> B, A, C
    assert [r["id"] for r in fused] == ["B", "A", "C"]
    for i, want in (("A", 0.46), ("B", 0.805), ("C", 0.3125)):
        assert abs(float(by[i]["fused"]) - want) < 1e-6
    # non-increasing
    assert all(fused[i]["fused"] >= fused[i + 1]["fused"] for i in range(len(fused) - 1))


def test_alpha_fusion_rerank_mode_nonpositive_max_and_truncate():
    dense = [("a", 0.5), ("b", 0.4)]
    sparse = [("b", 2.0), ("c", 1.0)]
    r = O.fuse_hybrid(dense, sparse, 0.0, 10)                 # alpha<=0: d + 0.1*s'
    by = {x["id"]: x["fused"] for x in r}
    assert by["a"] == f32(0.5)
    assert by["b"] == f32(f32(0.4) + f32(f32(1.0) * f32(0.1)))
    assert by["c"] == f32(f32(0.0) + f32(f32(0.5) * f32(0.1)))
    r = O.fuse_hybrid(dense, [("c", 0.0), ("d", -1.0)], 0.5, 10)   # max_sparse <= 0 -> zeros
    assert all(x["sparse_norm"] == 0.0 for x in r)
    r = O.fuse_hybrid([("x", 0.3), ("y", 0.3)], [], 1.0, 1)   # tie -> id asc, truncate
    assert [x["id"] for x in r] == ["x"]


def test_candidate_count_and_cap_k():
    assert O.candidate_count_for(5, 500) == 500      # src/search/query.rs:1934
    assert O.candidate_count_for(20, 500) == 500
    assert O.candidate_count_for(200, 500) == 1000
    assert O.candidate_count_for((1 << 64) - 1, 500) == (1 << 64) - 1  # saturating
    assert O.cap_k_to_backend(441, 500) == 441       # src/search/query.rs:1905
    assert O.cap_k_to_backend(None, 500) == 500
    assert O.cap_k_to_backend(1024, 500) == 500


# --- a12: tests/router_test.rs:92-226, :257-342 -----------------------------------

def test_alpha_defaults_table():
    want = dict(identifier_lookup=.85, structural=.60, behavioral=1.0, conceptual=.80,
                multi_step=.10, negation=.80, type_filtered=.0, cross_language=.70, unknown=.80)
    for cat, a in want.items():
        assert O.resolve_splade_alpha(cat, env={}) == f32(a)


def test_alpha_env_precedence_and_clamping():
    env = {"CQS_SPLADE_ALPHA": "0.3", "CQS_SPLADE_ALPHA_STRUCTURAL": "0.9"}
    assert O.resolve_splade_alpha("structural", env=env) == f32(0.9)       # per-cat > global
    assert O.resolve_splade_alpha("conceptual", env=env) == f32(0.3)       # global > default
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA_STRUCTURAL": "NaN"}) == f32(0.6)
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA_STRUCTURAL": "inf"}) == f32(0.6)
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA": "garbage"}) == f32(0.6)
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA": "1.5"}) == f32(1.0)
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA": "-2"}) == f32(0.0)
    assert O.resolve_splade_alpha("structural", env={}, slot_table={"structural": 0.42}) == f32(0.42)
    assert O.resolve_splade_alpha("structural", env={"CQS_SPLADE_ALPHA": "0.3"},
                                  slot_table={"structural": 0.42}) == f32(0.3)
    assert O.apply_centroid_floor(0.1, True) == f32(0.7)
    assert O.apply_centroid_floor(0.1, False) == f32(0.1)
    assert O.apply_centroid_floor(0.85, True) == f32(0.85)


# --- a13 (unpinned in the reference; literal restatement) --------------------------

def test_centroid_classify_margin_gate():
    c = np.eye(3, 8, dtype=f32)
    q = np.zeros((3, 8), f32)
    q[0, 0] = 1.0                       # clear winner 0
    q[1, 0] = 0.5; q[1, 1] = 0.495      # margin .005 < .01 -> -1
    q[2, 2] = 0.3; q[2, 1] = 0.29       # margin .01 (>=) -> 2 (f32 rounding aside)
    cat, margin = O.centroid_classify(c, q, 0.01)
    assert cat[0] == 0 and cat[1] == -1
    assert abs(float(margin[1]) - 0.005) < 1e-6
    assert cat[2] in (2, -1)


# --- a14: src/search/scoring/fusion.rs:208-331 --------------------------------------

def test_rrf_fuse_n():
    assert O.rrf_fuse_n([], 10) == []
    r = dict(O.rrf_fuse_n([["a", "b", "c", "d", "a", "e"]], 10))
    assert abs(float(r["a"]) - 1.0 / 61.0) < 1e-6
    r = dict(O.rrf_fuse_n([["common", "x", "y"], ["common", "z"], ["common", "w"]], 10))
    assert abs(float(r["common"]) - 3.0 / 61.0) < 1e-6 and abs(float(r["x"]) - 1.0 / 62.0) < 1e-6
    r = O.rrf_fuse_n([["a", "b"], ["c", "d"], ["e", "f"], ["g", "h"]], 3)
    assert len(r) == 3 and r[0][1] >= r[1][1] >= r[2][1]
    r = dict(O.rrf_fuse_n([["common", "sem_only"], ["common", "fts_only"]], 10))
    assert abs(float(r["common"]) - 2.0 / 61.0) < 1e-6
    assert abs(float(r["sem_only"]) - 1.0 / 62.0) < 1e-6


# --- synthetic generator: examples/exp_level_scale.rs:200-224, :466-477 ---------------

def test_xorshift_known_answer():
    # first xorshift64 step of seed 0x9E3779B97F4A7C15 computed by hand below
    s = 0x9E3779B97F4A7C15
    m = (1 << 64) - 1
    s ^= (s << 13) & m; s ^= s >> 7; s ^= (s << 17) & m
    vals, st = O.xorshift_stream(1)
    assert st == s
    assert vals[0] == f32(f32(s >> 11) / f32(2.0 ** 53))
    assert 0.0 <= vals[0] <= 1.0


def test_synth_vectors_unit_norm_and_gt():
    v, _ = O.synth_vectors(64, 768)
    assert v.dtype == f32 and v.shape == (64, 768)
    assert np.allclose(np.linalg.norm(v.astype(np.float64), axis=1), 1.0, atol=1e-5)
    gt = O.brute_force_topk_f32(v[3], v, 5)
    assert gt[0] == 3
    rows, sc = O.brute_force_search(v, v[3], 5)
    assert rows[0] == 3 and sc[0] > 0.999
    assert set(rows.tolist()) == set(gt.tolist())


def test_brute_force_guards():
    v, _ = O.synth_vectors(8, 16)
    assert O.brute_force_search(v, v[0], 0)[0].size == 0
    assert O.brute_force_search(v[:0], v[0], 5)[0].size == 0
    assert O.brute_force_search(v, np.ones(15, f32), 5)[0].size == 0      # wrong dim
    q = v[0].copy(); q[3] = np.nan
    assert O.brute_force_search(v, q, 5)[0].size == 0                     # NaN query
    rows, _ = O.brute_force_search(v, v[0], 100)                         # k > n -> n unique
    assert sorted(rows.tolist()) == list(range(8))
    vz = v.copy(); vz[2] = 0.0
    rows, sc = O.brute_force_search(vz, v[0], 8)
    assert np.all(np.isfinite(sc)) and 2 in rows.tolist()                 # zero row finite (0.0)


def test_bitset_roundtrip_and_bf16_rne():
    rng = np.random.default_rng(5)
    m = rng.random(77) > 0.5
    bs = O.mask_to_bitset(m)
    assert bs.shape[0] == 3
    assert np.array_equal(O.bitset_to_mask(bs, 77), m)
    for i in range(77):
        assert bool((int(bs[i // 32]) >> (i % 32)) & 1) == bool(m[i])     # cagra.rs:747-757
    x = np.array([1.0, 1.00390625, 1.01171875, -2.5, 0.0, 3.1415927], f32)
    b = O.f32_to_bf16_rne(x)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(b, want)
    assert np.array_equal(O.bf16_to_f32(b), torch.from_numpy(x).to(torch.bfloat16).float().numpy())
