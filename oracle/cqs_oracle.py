"""CPU oracle for the cqs retrieval hot path (TEST INFRASTRUCTURE ONLY).

This module restates, in numpy / plain Python, the algorithms the reference
(jamie8johnson/cqs v1.51.0, Rust) runs on the path named by BASELINE.json's
north_star. It is the *checker* the CUDA library is compared against. Nothing
under ``cqs_b200/`` may import it; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do.

Pinning status:

* The reference cannot be compiled or run in this image (no cargo/rustc), and
  its inner f32 dot lives in the un-vendored crate ``simsimd 6.5.16``
  (Cargo.lock:4049-4052) whose SIMD lane order is not recoverable here.
  **Bit-level dense-score parity is therefore unpinned.**  The oracle value
  is the reference's own documented fallback: the f64-accumulated dot rounded
  to f32 (src/math.rs:17-22); the contract is <=1e-5 relative.
* Everything else (heap ties, sparse accumulate order, alpha fusion, alpha
  table, RRF, scoring fold, synthetic generator) is pinned against the
  reference's own unit-test vectors, reproduced in ``tests/test_oracle_golden.py``.
* Centroid classifier: the reference has no numeric unit test -> unpinned;
  the oracle follows src/search/router.rs:1415-1444 literally.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Iterable, Sequence

import numpy as np

# ---------------------------------------------------------------------------
# a1  math::cosine_similarity                                 src/math.rs:11-28
# ---------------------------------------------------------------------------


def cosine_similarity(a: np.ndarray, b: np.ndarray):
    """Dot of two (unit-norm) f32 vectors; ``None`` on length mismatch, empty
    input or non-finite result.  f64 accumulation rounded to f32 is the
    reference's scalar fallback (src/math.rs:17-22) and is *the* oracle value.
    """
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if a.shape[0] != b.shape[0] or a.shape[0] == 0:
        return None
    with np.errstate(all="ignore"):
        s = np.float32(np.dot(a.astype(np.float64), b.astype(np.float64)))
    if not np.isfinite(s):
        return None
    return s


def dense_scores(rows: np.ndarray, query: np.ndarray, block: int = 65536) -> np.ndarray:
    """a1 applied to every row: f64-accumulated dot, rounded to f32.

    Rows may be f32 or (bf16 already up-cast to) f32.  Non-finite results are
    returned as NaN so the caller's heap drops them, matching ``None`` ->
    skipped candidate at src/search/scoring/candidate.rs:570-578.
    """
    rows = np.asarray(rows)
    q64 = np.asarray(query, dtype=np.float32).astype(np.float64)
    n = rows.shape[0]
    out = np.empty(n, dtype=np.float32)
    with np.errstate(all="ignore"):
        for s in range(0, n, block):
            e = min(n, s + block)
            out[s:e] = (rows[s:e].astype(np.float64) @ q64).astype(np.float32)
    out[~np.isfinite(out)] = np.nan
    return out


# ---------------------------------------------------------------------------
# a4  BoundedScoreHeap                 src/search/scoring/candidate.rs:162-330
# ---------------------------------------------------------------------------


def _total_key(x: float) -> int:
    """Sort key equal to f32::total_cmp ordering (candidate.rs:202-204)."""
    u = int(np.float32(x).view(np.uint32))
    return (u ^ 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)


class BoundedScoreHeap:
    """Literal restatement (small inputs): push order independent result.

    * non-finite scores ignored (candidate.rs:275)
    * below capacity: insert (:281)
    * at capacity: evict worst iff ``score > worst`` or (``==`` and
      ``id < worst_id``), comparisons by total_cmp (:303-307)
    * ``into_sorted_vec``: (score desc by total_cmp, id asc) (:321-329)
    """

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.items: list[tuple[float, object]] = []

    def _worst_index(self) -> int:
        # worst under final order = smallest score, largest id among ties
        best_i = 0
        for i, (s, ident) in enumerate(self.items):
            ws, wid = self.items[best_i]
            ks, kw = _total_key(s), _total_key(ws)
            if ks < kw or (ks == kw and ident > wid):
                best_i = i
        return best_i

    def would_accept(self, score: float) -> bool:
        if not math.isfinite(score):
            return False
        if self.capacity == 0:
            return False
        if len(self.items) < self.capacity:
            return True
        ws, _ = self.items[self._worst_index()]
        return not (_total_key(score) < _total_key(ws))

    def push(self, ident, score: float) -> None:
        score = float(np.float32(score))
        if not math.isfinite(score):
            return
        if len(self.items) < self.capacity:
            self.items.append((score, ident))
            return
        if not self.items:
            return  # capacity 0
        wi = self._worst_index()
        ws, wid = self.items[wi]
        ks, kw = _total_key(score), _total_key(ws)
        if ks > kw or (ks == kw and ident < wid):
            self.items[wi] = (score, ident)

    def into_sorted_vec(self) -> list[tuple[object, float]]:
        out = sorted(self.items, key=lambda t: (-_total_key(t[0]), t[1]))
        return [(ident, s) for (s, ident) in out]


def ordered_u32(scores: np.ndarray) -> np.ndarray:
    """Monotone map f32 -> u32 reproducing total_cmp (vectorised _total_key)."""
    u = np.asarray(scores, dtype=np.float32).view(np.uint32)
    neg = (u & np.uint32(0x80000000)) != 0
    return np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def topk_rows(scores: np.ndarray, k: int, mask: np.ndarray | None = None):
    """Vectorised a4 for ids that are *row indices in ascending-id order*:
    the final content of a BoundedScoreHeap(k) fed every (row, score) is the
    k best under (score desc by total_cmp, row asc), non-finite dropped —
    independent of push order (candidate.rs:303-329).  Returns (rows, scores).
    """
    scores = np.asarray(scores, dtype=np.float32)
    ok = np.isfinite(scores)
    if mask is not None:
        ok &= np.asarray(mask, dtype=bool)
    idx = np.nonzero(ok)[0]
    if k <= 0 or idx.size == 0:
        return np.empty(0, np.int64), np.empty(0, np.float32)
    key = ordered_u32(scores[idx]).astype(np.int64)
    order = np.lexsort((idx, -key))  # primary: key desc, secondary: row asc
    sel = idx[order[:k]]
    return sel.astype(np.int64), scores[sel]


# ---------------------------------------------------------------------------
# a3  Store::search_filtered_with_notes (brute force)  src/search/query.rs:453-484
# ---------------------------------------------------------------------------


def brute_force_search(rows: np.ndarray, query: np.ndarray, k: int, mask=None):
    """Exact scan: a1 per row -> a4 heap of size k -> sorted (score desc, id asc).

    ``mask`` plays the role of the SQL type/language filter (rows that do not
    pass are never scored).  The multiplicative signals of a5 stay on the host
    side of the boundary and are applied by the caller.
    Guards mirror the index backends (src/cagra.rs:445-470): k==0, empty
    index, wrong dim or non-finite query -> empty.
    """
    rows = np.asarray(rows)
    query = np.asarray(query, dtype=np.float32)
    if k == 0 or rows.shape[0] == 0 or query.shape[0] != rows.shape[1]:
        return np.empty(0, np.int64), np.empty(0, np.float32)
    if not np.all(np.isfinite(query)):
        return np.empty(0, np.int64), np.empty(0, np.float32)
    return topk_rows(dense_scores(rows, query), k, mask)


def search_filtered(rows, query, limit, threshold=0.0, chunk_type=None, lang=None, include_types=None,
                    languages=None, note_boost=None, importance=None, enable_demotion=True):
    """Store::search_filtered_with_notes (src/search/query.rs:348-510) with
    SearchFilter::default()-style signals (no name matcher, no glob): SQL type/language
    filter, then per row score_candidate = cosine -> apply_scoring_pipeline
    (candidate.rs:538-578), then the bounded heap — the pipeline runs BEFORE the heap
    (query.rs:469-481).  Returns (rows, folded scores)."""
    rows = np.asarray(rows)
    query = np.asarray(query, dtype=np.float32)
    n = rows.shape[0]
    if limit == 0 or n == 0 or query.shape[0] != rows.shape[1] or not np.all(np.isfinite(query)):
        return np.empty(0, np.int64), np.empty(0, np.float32)
    ok = np.ones(n, bool)
    if include_types is not None:
        ok &= np.isin(np.asarray(chunk_type), list(include_types))
    if languages is not None:
        ok &= np.isin(np.asarray(lang), list(languages))
    cos = dense_scores(rows, query)
    folded = np.full(n, np.nan, np.float32)
    for r in np.nonzero(ok & np.isfinite(cos))[0]:
        s = apply_scoring_pipeline(cos[r], note_boost=1.0 if note_boost is None else note_boost[r],
                                   importance=(None if (importance is None or not enable_demotion) else importance[r]),
                                   threshold=threshold)
        if s is not None:
            folded[r] = s
    return topk_rows(folded, limit)


def bitset_to_mask(bitset: np.ndarray, n: int) -> np.ndarray:
    """Host filter bitset convention, bit i%32 of word i/32 (src/cagra.rs:747-757)."""
    bits = np.unpackbits(np.asarray(bitset, dtype="<u4").view(np.uint8), bitorder="little")
    return bits[:n].astype(bool)


def mask_to_bitset(mask: np.ndarray) -> np.ndarray:
    mask = np.asarray(mask, dtype=bool)
    n = mask.shape[0]
    pad = (-n) % 32
    bits = np.concatenate([mask, np.zeros(pad, bool)]).astype(np.uint8)
    return np.packbits(bits, bitorder="little").view("<u4").copy()


# ---------------------------------------------------------------------------
# bf16 storage (configs 3-4; SURVEY.md §8d "bf16 oracle note")
# ---------------------------------------------------------------------------


def f32_to_bf16_rne(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16, returned as uint16 bit patterns."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = np.isnan(np.asarray(x, dtype=np.float32))
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FC0
    return r


def bf16_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


# ---------------------------------------------------------------------------
# a10  SpladeIndex                                   src/splade/index.rs:177-291
# ---------------------------------------------------------------------------


class SpladeIndex:
    """postings: token -> [(chunk_index, weight)] in build (chunk) order."""

    def __init__(self, chunks: Sequence[tuple[object, Sequence[tuple[int, float]]]]):
        # build: src/splade/index.rs:191-211
        self.postings: dict[int, list[tuple[int, np.float32]]] = {}
        self.id_map: list[object] = []
        for idx, (chunk_id, sparse) in enumerate(chunks):
            for tok, w in sparse:
                self.postings.setdefault(int(tok), []).append((idx, np.float32(w)))
            self.id_map.append(chunk_id)

    def __len__(self):
        return len(self.id_map)

    def raw_scores(self, query, filt: Callable[[object], bool] | None = None) -> dict[int, np.float32]:
        """search_with_filter's accumulate loop (:251-261): per chunk, in
        *query-token order*, ``score += qw * dw`` with separate f32 mul/add;
        only chunks touched by >=1 posting exist in the map."""
        scores: dict[int, np.float32] = {}
        with np.errstate(all="ignore"):
            for tok, qw in query:
                qw = np.float32(qw)
                for idx, dw in self.postings.get(int(tok), ()):
                    if filt is not None and not filt(self.id_map[idx]):
                        continue
                    scores[idx] = np.float32(scores.get(idx, np.float32(0.0)) + np.float32(qw * dw))
        return scores

    def search_with_filter(self, query, k: int, filt=None):
        """-> [(id, score)] sorted (score desc, id asc)   (:223-291)"""
        if len(query) == 0 or len(self.id_map) == 0:
            return []
        heap = BoundedScoreHeap(k)
        for idx, s in self.raw_scores(query, filt).items():
            if not heap.would_accept(float(s)):
                continue
            heap.push(self.id_map[idx], float(s))
        return [(i, np.float32(s)) for i, s in heap.into_sorted_vec()]

    def search(self, query, k: int):
        return self.search_with_filter(query, k, None)


def sparse_scores_csr(indptr, tok, w, q_tok, q_w, n_docs, mask=None):
    """Vectorised a10 over a doc-major CSR whose rows are in ascending-id order.

    Returns (scores f32[n_docs], touched bool[n_docs]).  Accumulation order per
    doc is the *query-token order* (src/splade/index.rs:251-259): we loop over
    query tokens sequentially and add each token's contribution to all docs at
    once (a doc appears at most once per token when its token ids are unique;
    duplicates inside one doc are accumulated in doc order by np.add.at's
    sequential semantics).
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    tok = np.asarray(tok, dtype=np.uint32)
    w = np.asarray(w, dtype=np.float32)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), np.diff(indptr))
    scores = np.zeros(n_docs, dtype=np.float32)
    touched = np.zeros(n_docs, dtype=bool)
    order = np.argsort(tok, kind="stable")
    tok_sorted = tok[order]
    with np.errstate(all="ignore"):
        for t, qw in zip(np.asarray(q_tok, dtype=np.uint32), np.asarray(q_w, dtype=np.float32)):
            lo = np.searchsorted(tok_sorted, t, "left")
            hi = np.searchsorted(tok_sorted, t, "right")
            ent = order[lo:hi]  # ascending entry index == ascending doc (stable)
            d = doc_of[ent]
            if mask is not None:
                keep = mask[d]
                ent, d = ent[keep], d[keep]
            contrib = (np.float32(qw) * w[ent]).astype(np.float32)
            if d.size and np.unique(d).size != d.size:
                np.add.at(scores, d, contrib)  # sequential f32 adds, doc order
            else:
                scores[d] = scores[d] + contrib
            touched[d] = True
    return scores, touched


def sparse_search_csr(indptr, tok, w, q_tok, q_w, n_docs, k, mask=None):
    if len(q_tok) == 0 or n_docs == 0:
        return np.empty(0, np.int64), np.empty(0, np.float32)
    s, touched = sparse_scores_csr(indptr, tok, w, q_tok, q_w, n_docs, mask)
    return topk_rows(s, k, touched)


# ---------------------------------------------------------------------------
# a11  alpha fusion in search_hybrid_inner       src/search/query.rs:914-1005
# ---------------------------------------------------------------------------


def fuse_hybrid(dense: Sequence[tuple[object, float]], sparse: Sequence[tuple[object, float]],
                alpha: float, candidate_count: int):
    """dense / sparse: [(id, score)] pools, each sorted as its leg returned it.

    Returns list of dicts {id, fused, dense, sparse_raw, sparse_norm,
    in_dense, in_sparse} sorted (fused desc total_cmp, id asc), truncated.
    Every arithmetic step is f32 exactly as written in the reference:
      max_sparse = reduce(f32::max) or 0.0                      (:914-919)
      s' = s / max_sparse if max_sparse > 0 else 0               (:945-950)
      union in insertion order, dense first                      (:958-971)
      alpha <= 0 -> d + s'*0.1 ; else alpha*d + (1-alpha)*s'     (:984-997)
      sort (score desc, id asc), truncate                        (:1004-1005)
    """
    f = np.float32
    alpha = f(alpha)
    max_sparse = f(0.0)
    if len(sparse):
        max_sparse = f(sparse[0][1])
        for _, s in sparse[1:]:
            # f32::max: NaN-ignoring
            s = f(s)
            if np.isnan(max_sparse) or (not np.isnan(s) and s > max_sparse):
                max_sparse = s
    dense_scores_m = {}
    for i, s in dense:
        dense_scores_m[i] = f(s)  # later duplicates overwrite, as HashMap::insert
    sparse_raw, sparse_norm = {}, {}
    with np.errstate(all="ignore"):
        for i, s in sparse:
            sparse_raw[i] = f(s)
            sparse_norm[i] = f(f(s) / max_sparse) if max_sparse > 0 else f(0.0)
    all_ids, seen = [], set()
    for i, _ in list(dense) + list(sparse):
        if i not in seen:
            seen.add(i)
            all_ids.append(i)
    out = []
    with np.errstate(all="ignore"):
        for i in all_ids:
            d = dense_scores_m.get(i, f(0.0))
            s = sparse_norm.get(i, f(0.0))
            if alpha <= 0:
                score = f(d + f(s * f(0.1)))
            else:
                score = f(f(alpha * d) + f(f(f(1.0) - alpha) * s))
            out.append(dict(id=i, fused=score, dense=d, sparse_raw=sparse_raw.get(i, f(0.0)),
                            sparse_norm=s, in_dense=i in dense_scores_m, in_sparse=i in sparse_raw))
    out.sort(key=lambda r: (-_total_key(r["fused"]), r["id"]))
    return out[:candidate_count]


def candidate_count_for(limit: int, floor: int | None = None) -> int:
    """max(5*limit, floor=500)   src/limits.rs:315-320 (env CQS_SEARCH_CANDIDATE_FLOOR)."""
    if floor is None:
        try:
            floor = int(os.environ.get("CQS_SEARCH_CANDIDATE_FLOOR", "500"))
        except ValueError:
            floor = 500
    usize_max = (1 << 64) - 1
    return max(min(limit * 5, usize_max), floor)


def cap_k_to_backend(max_k, k: int) -> int:
    """src/search/query.rs:232-245"""
    return max_k if (max_k is not None and k > max_k) else k


# ---------------------------------------------------------------------------
# a12  per-category alpha                 src/search/router.rs:126-175, :708-833
# ---------------------------------------------------------------------------

CATEGORIES = (
    "identifier_lookup", "structural", "behavioral", "conceptual", "multi_step",
    "negation", "type_filtered", "cross_language", "unknown",
)
DEFAULT_ALPHA = {
    "identifier_lookup": 0.85, "structural": 0.60, "behavioral": 1.00,
    "conceptual": 0.80, "multi_step": 0.10, "negation": 0.80,
    "type_filtered": 0.00, "cross_language": 0.70, "unknown": 0.80,
}
CENTROID_ALPHA_FLOOR = 0.7  # src/cli/commands/search/query.rs:655-657


def _parse_f32(val: str):
    """Rust ``str::parse::<f32>``: no surrounding whitespace, no underscores."""
    if val != val.strip() or "_" in val or val == "":
        return None
    try:
        with np.errstate(all="ignore"):
            return np.float32(float(val))
    except ValueError:
        return None


def resolve_splade_alpha(category: str, env=None, slot_table=None) -> np.float32:
    """Precedence per-category env > global env > slot.toml > default; finite
    values clamp to [0,1]; NaN/Inf/garbage fall through (router.rs:708-833)."""
    env = os.environ if env is None else env
    for key in (f"CQS_SPLADE_ALPHA_{category.upper()}", "CQS_SPLADE_ALPHA"):
        val = env.get(key)
        if val is not None:
            a = _parse_f32(val)
            if a is not None and np.isfinite(a):
                return np.float32(min(max(a, np.float32(0.0)), np.float32(1.0)))
    if slot_table:
        a = slot_table.get(category.lower())
        if a is not None and np.isfinite(np.float32(a)):
            return np.float32(min(max(np.float32(a), np.float32(0.0)), np.float32(1.0)))
    return np.float32(DEFAULT_ALPHA[category])


def apply_centroid_floor(alpha, centroid_applied: bool) -> np.float32:
    """src/cli/commands/search/query.rs:655-657"""
    a = np.float32(alpha)
    return np.float32(max(a, np.float32(CENTROID_ALPHA_FLOOR))) if centroid_applied else a


# ---------------------------------------------------------------------------
# a13  CentroidClassifier::classify               src/search/router.rs:1415-1444
# ---------------------------------------------------------------------------


def centroid_scores(centroids: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """score[q,c] = sequential f32 sum_i e_i*c_i (mul, then add, index order)."""
    c = np.asarray(centroids, dtype=np.float32)
    q = np.asarray(queries, dtype=np.float32)
    acc = np.zeros((q.shape[0], c.shape[0]), dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(c.shape[1]):
            acc = acc + (q[:, i:i + 1] * c[None, :, i]).astype(np.float32)
    return acc


def centroid_classify(centroids: np.ndarray, queries: np.ndarray, threshold: float = 0.01):
    """Returns (cat int32[nq] (-1 below margin), margin f32[nq]).

    Iteration is in centroid index order (the reference iterates a HashMap, so
    the winner among *exactly equal* scores is unspecified there)."""
    sc = centroid_scores(centroids, queries)
    nq, nc = sc.shape
    cat = np.full(nq, -1, dtype=np.int32)
    margin = np.zeros(nq, dtype=np.float32)
    thr = np.float32(threshold)
    for qi in range(nq):
        best, second, bc = np.float32(-np.inf), np.float32(-np.inf), -1
        for ci in range(nc):
            s = sc[qi, ci]
            if s > best:
                second, best, bc = best, s, ci
            elif s > second:
                second = s
        with np.errstate(all="ignore"):
            m = np.float32(best - second)
        margin[qi] = m
        if m >= thr:
            cat[qi] = bc
    return cat, margin


# ---------------------------------------------------------------------------
# a5  apply_scoring_pipeline (numeric fold)  src/search/scoring/candidate.rs:420-562
# ---------------------------------------------------------------------------


def apply_scoring_pipeline(embedding_score, *, name_boost=None, name_score=0.0, glob_ok=True,
                           note_boost=1.0, importance=None, threshold=0.0):
    """Order-sensitive f32 fold.  ``name_boost=None`` = no name matcher;
    ``importance=None`` = demotion disabled.  Returns f32 or None."""
    f = np.float32
    s = f(min(max(f(embedding_score), f(0.0)), f(1.0)))
    if name_boost is not None:
        nb = f(min(max(f(name_boost), f(0.0)), f(1.0)))
        s = f(f(f(f(1.0) - nb) * s) + f(nb * f(name_score)))
    if not glob_ok:
        return None
    s = f(f(max(s, f(0.0))) * f(note_boost))
    if importance is not None:
        s = f(s * f(importance))
    if s >= f(threshold):
        return s
    return None


# ---------------------------------------------------------------------------
# a14  rrf_fuse_n                           src/search/scoring/fusion.rs:36-68
# ---------------------------------------------------------------------------


def rrf_fuse_n(ranked_lists: Sequence[Sequence[object]], limit: int, k: float = 60.0):
    f = np.float32
    scores: dict[object, np.float32] = {}
    for lst in ranked_lists:
        seen = set()
        for rank, ident in enumerate(lst):
            if ident in seen:
                continue
            seen.add(ident)
            contribution = f(f(1.0) / f(f(f(k) + f(rank)) + f(1.0)))
            scores[ident] = f(scores.get(ident, f(0.0)) + contribution)
    heap = BoundedScoreHeap(limit)
    for ident, s in scores.items():
        heap.push(ident, float(s))
    return [(i, np.float32(s)) for i, s in heap.into_sorted_vec()]


# ---------------------------------------------------------------------------
# Synthetic generator                      examples/exp_level_scale.rs:200-224
# ---------------------------------------------------------------------------

XORSHIFT_SEED = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def xorshift_stream(count: int, state: int = XORSHIFT_SEED):
    """`count` draws of ``(state >> 11) as f32 / 2^53`` after each xorshift64
    step (x^=x<<13; x^=x>>7; x^=x<<17).  Returns (f32[count], final_state)."""
    out = np.empty(count, dtype=np.uint64)
    s = state
    for i in range(count):
        s ^= (s << 13) & _M64
        s ^= s >> 7
        s ^= (s << 17) & _M64
        out[i] = s >> 11
    # u64 -> f32 cast rounds to nearest even; then divide by 2^53 (exact in f32)
    vals = out.astype(np.float32) / np.float32(2.0 ** 53)
    return vals, s


def synth_vectors(n: int, dim: int = 768, state: int = XORSHIFT_SEED):
    """Reference recipe verbatim: x = next()*2-1 (f32), norm = sqrt(sequential
    f32 sum of x*x), x /= norm.  Returns (f32[n,dim], final_state)."""
    u, state = xorshift_stream(n * dim, state)
    v = (u * np.float32(2.0) - np.float32(1.0)).astype(np.float32).reshape(n, dim)
    sq = (v * v).astype(np.float32)
    acc = np.zeros(n, dtype=np.float32)
    for i in range(dim):  # sequential f32 sum, as Iterator::sum::<f32>
        acc = acc + sq[:, i]
    norm = np.sqrt(acc).astype(np.float32)
    ok = norm > 0
    v[ok] = (v[ok] / norm[ok, None]).astype(np.float32)
    return v, state


def brute_force_topk_f32(query: np.ndarray, vectors: np.ndarray, k: int) -> np.ndarray:
    """exp_level_scale.rs:466-477 ground truth: sequential f32 dot, sort desc."""
    q = np.asarray(query, dtype=np.float32)
    v = np.asarray(vectors, dtype=np.float32)
    acc = np.zeros(v.shape[0], dtype=np.float32)
    for i in range(v.shape[1]):
        acc = acc + (v[:, i] * q[i]).astype(np.float32)
    return np.argsort(-acc, kind="stable")[:k]


# ---------------------------------------------------------------------------
# Fast synthetic corpora for the large configs (not a reference recipe; the
# generator is splitmix64-seeded per row block so shards can be produced
# independently — SURVEY.md §8d).  Used by tests and bench only.
# ---------------------------------------------------------------------------


def fast_unit_rows(n: int, dim: int, seed: int, clustered: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if clustered:
        centres = rng.standard_normal((256, dim)).astype(np.float32)
        centres /= np.linalg.norm(centres, axis=1, keepdims=True)
        pick = rng.integers(0, 256, size=n)
        # noise of norm ~0.6 around a unit centre: same-cluster cosine ~0.7
        x = centres[pick] + rng.standard_normal((n, dim)).astype(np.float32) * np.float32(0.6 / math.sqrt(dim))
    else:
        x = rng.uniform(-1.0, 1.0, size=(n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def synth_sparse(n_docs: int, vocab: int = 30522, mean_nnz: int = 200, seed: int = 7,
                 lo: int = 20, hi: int = 400, zipf_s: float = 1.1):
    """Doc-major CSR: nnz ~ Poisson(mean) clipped [lo,hi]; token ids Zipf(s)
    without replacement within a doc, ascending; weights log1p(relu(N(.8,.5)))
    kept > 0.01 (mirrors src/splade/mod.rs:721-727, threshold :405-413)."""
    rng = np.random.default_rng(seed)
    nnz = np.clip(rng.poisson(mean_nnz, size=n_docs), lo, min(hi, vocab)).astype(np.int64)
    p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** zipf_s
    cdf = np.cumsum(p / p.sum())
    indptr = np.zeros(n_docs + 1, dtype=np.int64)
    toks, ws = [], []
    for d in range(n_docs):
        want = int(nnz[d])
        got = np.unique(np.searchsorted(cdf, rng.random(want * 2)).clip(0, vocab - 1))
        while got.size < want:
            extra = np.searchsorted(cdf, rng.random(want * 2)).clip(0, vocab - 1)
            got = np.unique(np.concatenate([got, extra]))
        if got.size > want:
            got = np.sort(rng.choice(got, size=want, replace=False))
        wt = np.log1p(np.maximum(rng.normal(0.8, 0.5, size=want), 0.0)).astype(np.float32)
        keep = wt > np.float32(0.01)
        toks.append(got[keep].astype(np.uint32))
        ws.append(wt[keep])
        indptr[d + 1] = indptr[d] + int(keep.sum())
    return indptr, np.concatenate(toks) if toks else np.empty(0, np.uint32), \
        np.concatenate(ws) if ws else np.empty(0, np.float32)
