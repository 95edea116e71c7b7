"""Mirror of the SPLADE leg and the hybrid fusion entry
(src/splade/index.rs:177-291, src/search/query.rs:811-1005) on top of the
B200 index: same call shapes, id strings on this side, arithmetic on the GPU."""
from __future__ import annotations

import os
from typing import Callable, Optional, Sequence

import numpy as np

from . import capi
from .index import B200Index, IndexResult


def candidate_count_for(limit: int) -> int:
    """src/limits.rs:315-320"""
    try:
        floor = int(os.environ.get("CQS_SEARCH_CANDIDATE_FLOOR", "500"))
    except ValueError:
        floor = 500
    return max(min(limit * 5, (1 << 64) - 1), floor)


def cap_k_to_backend(index, k: int) -> int:
    """src/search/query.rs:232-245"""
    cap = index.max_k()
    return cap if (cap is not None and k > cap) else k


class SpladeIndex:
    """SpladeIndex::build over the SAME chunk ids as a B200Index: the sparse
    vectors are uploaded as a CSR aligned with the dense rows."""

    def __init__(self, index: B200Index, chunks: Sequence[tuple[str, Sequence[tuple[int, float]]]],
                 vocab: int = 30522):
        self.index = index
        by_id = {cid: sv for cid, sv in chunks}
        indptr = np.zeros(len(index.id_map) + 1, np.uint64)
        toks, ws = [], []
        for r, cid in enumerate(index.id_map):
            sv = by_id.get(cid, ())
            toks.extend(int(t) for t, _ in sv)
            ws.extend(float(w) for _, w in sv)
            indptr[r + 1] = len(toks)
        self.n_chunks = len(by_id)
        index.sparse_attach(indptr, np.asarray(toks, np.uint32), np.asarray(ws, np.float32), vocab)

    def __len__(self):
        return self.n_chunks

    def search_with_filter(self, query: Sequence[tuple[int, float]], k: int,
                           flt: Optional[Callable[[str], bool]] = None) -> list[IndexResult]:
        ix = self.index
        if len(query) == 0 or len(ix.id_map) == 0 or k == 0:
            return []
        bitset = None
        if flt is not None:
            bitset, included = ix.bitset_for(flt)
            if included == 0:
                return []
        t = np.asarray([q[0] for q in query], np.uint32)
        w = np.asarray([q[1] for q in query], np.float32)
        try:
            rows, scores = ix.search_sparse_rows(t, w, min(k, capi.MAX_K), bitset)
        except capi.B200Error:
            return []
        return ix._results(rows, scores)

    def search(self, query, k: int):
        return self.search_with_filter(query, k, None)


def search_hybrid(index: B200Index, query: np.ndarray, sparse_query: Sequence[tuple[int, float]],
                  alpha: float, limit: int, flt: Optional[Callable[[str], bool]] = None):
    """The two leg calls + fusion of search_hybrid_inner (query.rs:880-1005) as ONE
    library call.  Returns the fused pool: list of dicts {id, fused, dense,
    sparse_raw, in_dense, in_sparse}, (fused desc, id asc), len <= candidate_count."""
    candidate_count = candidate_count_for(limit)
    pool_k = cap_k_to_backend(index, candidate_count)
    bitset = None
    if flt is not None:
        bitset, included = index.bitset_for(flt)
        if included == 0:
            return []
    t = np.asarray([q[0] for q in sparse_query], np.uint32)
    w = np.asarray([q[1] for q in sparse_query], np.float32)
    try:
        r = index.search_hybrid_rows(query, t, w, alpha, pool_k, bitset)
    except capi.B200Error:
        return []
    out = []
    for i in range(r["rows"].shape[0]):
        local = int(r["rows"][i]) - index.row_base
        if 0 <= local < len(index.id_map):
            out.append(dict(id=index.id_map[local], fused=np.float32(r["fused"][i]),
                            dense=np.float32(r["dense"][i]), sparse_raw=np.float32(r["sparse_raw"][i]),
                            in_dense=bool(r["present"][i] & 1), in_sparse=bool(r["present"][i] & 2)))
    return out
