"""Python mirror of the reference's ``VectorIndex`` trait for the B200 backend
(src/index.rs:139-239) — the shape the Rust shim in INTEGRATION.md has.

Chunk-id strings stay on this side (``id_map``), the library speaks row
indices (src/cagra.rs:268).  Rows are sorted by chunk id once at build so the
device tie-break (row asc) equals the reference's (id asc).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import capi
from .capi import lib, check, ptr


@dataclass
class IndexResult:  # src/index.rs:129
    id: str
    score: float


def _sort_key(s: str) -> bytes:
    return s.encode("utf-8")  # Rust String ordering is bytewise


class B200Index:
    """impl VectorIndex for B200Index (name() == "B200")."""

    def __init__(self, dim: int, metric: str = "cosine", storage: str = "f32",
                 devices: Optional[Sequence[int]] = None, row_base: int = 0):
        self._h = C.c_void_p()
        self._dim = int(dim)
        self.storage = storage
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        check(lib.cqs_b200_create(arr, len(devs), self._dim,
                                  capi.METRIC_COSINE if metric == "cosine" else capi.METRIC_DOT,
                                  {"f32": capi.STORAGE_F32, "bf16": capi.STORAGE_BF16,
                                   "bf16+f32": capi.STORAGE_BF16_F32}[storage],
                                  C.byref(self._h)))
        self.id_map: list[str] = []
        self.row_base = int(row_base)
        if row_base:
            check(lib.cqs_b200_set_row_base(self._h, int(row_base)))
        self._bitset_cache: dict = {}

    # ---- build -----------------------------------------------------------
    @classmethod
    def build(cls, ids: Sequence[str], embeddings: np.ndarray, **kw) -> "B200Index":
        """The try_open / build_from_store analogue (src/cagra.rs:842-919): takes the
        (chunk_id, embedding) feed in any order, drops zero / non-finite rows like
        prepare_index_data (src/hnsw/mod.rs:716-736), sorts by chunk id, uploads."""
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        ok = np.isfinite(emb).all(axis=1) & (np.abs(emb).sum(axis=1) > 0)
        keep = np.nonzero(ok)[0]
        order = sorted(keep.tolist(), key=lambda i: _sort_key(ids[i]))
        ix = cls(emb.shape[1], **kw)
        ix.reserve(len(order))
        ix.append([ids[i] for i in order], emb[order])
        ix.finalize()
        return ix

    # ---- persistence (src/cagra.rs:963-1652 analogue: blob + id sidecar) -----------------
    def save(self, path: str) -> None:
        import json
        check(lib.cqs_b200_save(self._h, path.encode()))
        with open(path + ".ids.json.tmp", "w") as f:
            json.dump({"magic": "cqs-b200-ids-v1", "dim": self._dim, "chunk_count": len(self.id_map),
                       "storage": self.storage, "id_map": self.id_map}, f)
        import os
        os.replace(path + ".ids.json.tmp", path + ".ids.json")

    @classmethod
    def load(cls, path: str, devices: Optional[Sequence[int]] = None) -> Optional["B200Index"]:
        """None on any mismatch (caller rebuilds from the store, src/cagra.rs:1676-1802)."""
        import json
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        if lib.cqs_b200_load(path.encode(), arr, len(devs), C.byref(h)) != capi.OK:
            return None
        self = cls.__new__(cls)
        self._h = h
        self._dim = int(lib.cqs_b200_dim(h))
        self.row_base = 0
        self._bitset_cache = {}
        self.id_map = []
        self.storage = "?"
        try:
            with open(path + ".ids.json") as f:
                meta = json.load(f)
            if meta.get("magic") != "cqs-b200-ids-v1" or meta.get("chunk_count") != len(self):
                raise ValueError("sidecar mismatch")
            self.id_map = list(meta["id_map"])
            self.storage = meta.get("storage", "?")
        except (OSError, ValueError, KeyError):
            self.close()
            return None
        return self

    def reserve(self, n: int) -> None:
        check(lib.cqs_b200_reserve(self._h, int(n)))

    def append(self, ids: Optional[Sequence[str]], rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        assert rows.ndim == 2 and rows.shape[1] == self._dim
        check(lib.cqs_b200_append_rows_f32(self._h, ptr(rows), rows.shape[0]))
        if ids is not None:
            self.id_map.extend(ids)

    def append_device(self, d_ptr: int, n_rows: int) -> None:
        check(lib.cqs_b200_append_rows_f32_device(self._h, C.c_void_p(d_ptr), int(n_rows)))

    def finalize(self) -> None:
        check(lib.cqs_b200_finalize(self._h))

    def close(self) -> None:
        if self._h:
            lib.cqs_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- VectorIndex -------------------------------------------------------
    def __len__(self) -> int:
        return int(lib.cqs_b200_len(self._h))

    def is_empty(self) -> bool:
        return len(self) == 0

    def name(self) -> str:
        return lib.cqs_b200_name().decode()

    def dim(self) -> int:
        return int(lib.cqs_b200_dim(self._h))

    def is_poisoned(self) -> bool:
        return bool(lib.cqs_b200_is_poisoned(self._h))

    def max_k(self) -> Optional[int]:
        return int(lib.cqs_b200_max_k(self._h))

    def index_scores_are_cosine(self) -> bool:
        return bool(lib.cqs_b200_scores_are_cosine(self._h))

    def set_timing(self, enable: bool = True) -> None:
        check(lib.cqs_b200_set_timing(self._h, int(enable)))

    def last_kernel_ms(self) -> float:
        return float(lib.cqs_b200_last_kernel_ms(self._h))

    def search_rows(self, query: np.ndarray, k: int, bitset: Optional[np.ndarray] = None):
        """Row-level search: (rows u64[n], scores f32[n]).  Errors surface as
        exceptions here; ``search`` maps them to an empty result like the shim."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim != 1 or q.shape[0] != self._dim:
            return np.empty(0, np.uint64), np.empty(0, np.float32)  # src/cagra.rs:449-456
        kk = max(int(k), 0)
        rows = np.empty(max(kk, 1), np.uint64)
        scores = np.empty(max(kk, 1), np.float32)
        n = C.c_uint32(0)
        bs = None if bitset is None else np.ascontiguousarray(bitset, dtype=np.uint32)
        check(lib.cqs_b200_search(self._h, ptr(q), kk, ptr(bs), ptr(rows), ptr(scores), C.byref(n)))
        return rows[: n.value].copy(), scores[: n.value].copy()

    def _results(self, rows, scores) -> list[IndexResult]:
        out = []
        for r, s in zip(rows.tolist(), scores.tolist()):
            local = r - self.row_base
            if 0 <= local < len(self.id_map):  # drop out-of-range slots (src/cagra.rs:642-669)
                out.append(IndexResult(self.id_map[local], float(np.float32(s))))
        return out

    def search(self, query: np.ndarray, k: int) -> list[IndexResult]:
        try:
            return self._results(*self.search_rows(query, k))
        except capi.B200Error:
            return []  # every device failure -> empty Vec (src/cagra.rs:541-627)

    def bitset_for(self, flt: Callable[[str], bool]):
        """Predicate -> (bitset, included): bit i%32 of word i/32 (src/cagra.rs:747-757)."""
        n = len(self.id_map)
        mask = np.fromiter((bool(flt(i)) for i in self.id_map), dtype=bool, count=n)
        pad = (-n) % 32
        bits = np.concatenate([mask, np.zeros(pad, bool)]).astype(np.uint8)
        return np.packbits(bits, bitorder="little").view("<u4").copy(), int(mask.sum())

    def search_with_filter(self, query: np.ndarray, k: int, flt: Callable[[str], bool]) -> list[IndexResult]:
        n = len(self.id_map)
        if n == 0 or k == 0:
            return []
        bitset, included = self.bitset_for(flt)
        if included == n:               # all pass -> unfiltered path (src/cagra.rs:759-762)
            return self.search(query, k)
        if included == 0:               # none pass (:764-767)
            return []
        try:
            return self._results(*self.search_rows(query, min(k, included), bitset))
        except capi.B200Error:
            return []

    # ---- Store::search_filtered on device (src/search/query.rs:316-510) -------------
    def set_row_meta(self, chunk_type, lang) -> None:
        ct = None if chunk_type is None else np.ascontiguousarray(chunk_type, dtype=np.uint8)
        lg = None if lang is None else np.ascontiguousarray(lang, dtype=np.uint8)
        check(lib.cqs_b200_set_row_meta(self._h, ptr(ct), ptr(lg), len(self)))

    def set_row_signals(self, note_boost=None, importance=None) -> None:
        nb = None if note_boost is None else np.ascontiguousarray(note_boost, dtype=np.float32)
        im = None if importance is None else np.ascontiguousarray(importance, dtype=np.float32)
        check(lib.cqs_b200_set_row_signals(self._h, ptr(nb), ptr(im), len(self)))

    @staticmethod
    def _mask256(codes):
        if codes is None:
            return None
        m = np.zeros(4, np.uint64)
        for c in codes:
            m[int(c) >> 6] |= np.uint64(1) << np.uint64(int(c) & 63)
        return m

    def search_filtered_rows(self, query, limit: int, threshold: float = 0.0, include_types=None,
                             languages=None, enable_demotion: bool = True):
        """search_filtered(query, filter, limit, threshold) with SearchFilter::default()-style
        signals (no name matcher / glob): returns (rows, folded scores)."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim != 1 or q.shape[0] != self._dim:
            return np.empty(0, np.uint64), np.empty(0, np.float32)
        kk = max(int(limit), 0)
        rows = np.empty(max(kk, 1), np.uint64)
        scores = np.empty(max(kk, 1), np.float32)
        n = C.c_uint32(0)
        tm, lm = self._mask256(include_types), self._mask256(languages)
        check(lib.cqs_b200_search_filtered(self._h, ptr(q), kk, float(threshold), ptr(tm), ptr(lm),
                                           int(bool(enable_demotion)), ptr(rows), ptr(scores), C.byref(n)))
        return rows[: n.value].copy(), scores[: n.value].copy()

    def search_typed_rows(self, query, k: int, include_types=None, languages=None):
        """search_with_filter for an include_types / languages predicate, tested on device."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        if q.ndim != 1 or q.shape[0] != self._dim:
            return np.empty(0, np.uint64), np.empty(0, np.float32)
        kk = max(int(k), 0)
        rows = np.empty(max(kk, 1), np.uint64)
        scores = np.empty(max(kk, 1), np.float32)
        n = C.c_uint32(0)
        tm, lm = self._mask256(include_types), self._mask256(languages)
        check(lib.cqs_b200_search_typed(self._h, ptr(q), kk, ptr(tm), ptr(lm), ptr(rows), ptr(scores), C.byref(n)))
        return rows[: n.value].copy(), scores[: n.value].copy()

    # ---- inherent (no trait counterpart) ---------------------------------------
    def search_batch_rows(self, queries: np.ndarray, k: int, bitset: Optional[np.ndarray] = None):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self._dim
        nq = q.shape[0]
        rows = np.empty((nq, max(k, 1)), np.uint64)
        scores = np.empty((nq, max(k, 1)), np.float32)
        n = np.zeros(nq, np.uint32)
        bs = None if bitset is None else np.ascontiguousarray(bitset, dtype=np.uint32)
        check(lib.cqs_b200_search_batch(self._h, ptr(q), nq, int(k), ptr(bs), ptr(rows), ptr(scores), ptr(n)))
        return rows, scores, n

    # ---- SPLADE leg / hybrid (row level) ------------------------------------------
    def sparse_attach(self, indptr: np.ndarray, tok: np.ndarray, w: np.ndarray, vocab: int) -> None:
        ip = np.ascontiguousarray(indptr, dtype=np.uint64)
        t = np.ascontiguousarray(tok, dtype=np.uint32)
        ww = np.ascontiguousarray(w, dtype=np.float32)
        check(lib.cqs_b200_sparse_attach(self._h, ptr(ip), ptr(t), ptr(ww), int(vocab)))

    def sparse_attach_device(self, d_indptr: int, d_tok: int, d_w: int, nnz: int, vocab: int) -> None:
        """Doc-major CSR already on the device (raw pointers); the inverted index is built there."""
        check(lib.cqs_b200_sparse_attach_device(self._h, C.c_void_p(d_indptr), C.c_void_p(d_tok), C.c_void_p(d_w),
                                                int(nnz), int(vocab)))

    def sparse_save(self, path: str, generation: int) -> None:
        """SpladeIndex::save (src/splade/index.rs:308): posting lists + the store's splade_generation."""
        check(lib.cqs_b200_sparse_save(self._h, path.encode(), int(generation)))

    def sparse_load(self, path: str, expected_generation: int) -> bool:
        """SpladeIndex::load: False when the file is missing, damaged or stale (caller rebuilds)."""
        return lib.cqs_b200_sparse_load(self._h, path.encode(), int(expected_generation)) == 0

    def search_sparse_rows(self, q_tok, q_w, k: int, bitset=None):
        t = np.ascontiguousarray(q_tok, dtype=np.uint32)
        w = np.ascontiguousarray(q_w, dtype=np.float32)
        rows = np.empty(max(k, 1), np.uint64)
        scores = np.empty(max(k, 1), np.float32)
        n = C.c_uint32(0)
        bs = None if bitset is None else np.ascontiguousarray(bitset, dtype=np.uint32)
        check(lib.cqs_b200_search_sparse(self._h, ptr(t), ptr(w), t.shape[0], int(k), ptr(bs),
                                         ptr(rows), ptr(scores), C.byref(n)))
        return rows[: n.value].copy(), scores[: n.value].copy()

    def search_hybrid_rows(self, query, q_tok, q_w, alpha: float, pool_k: int, bitset=None, peer=None):
        q = np.ascontiguousarray(query, dtype=np.float32)
        t = np.ascontiguousarray(q_tok, dtype=np.uint32)
        w = np.ascontiguousarray(q_w, dtype=np.float32)
        kk = max(int(pool_k), 1)
        rows = np.empty(kk, np.uint64)
        fused = np.empty(kk, np.float32)
        dense = np.empty(kk, np.float32)
        sraw = np.empty(kk, np.float32)
        present = np.empty(kk, np.uint8)
        n = C.c_uint32(0)
        bs = None if bitset is None else np.ascontiguousarray(bitset, dtype=np.uint32)
        if peer is not None:   # row-sharded corpus: every rank gets the global fused list
            check(lib.cqs_b200_search_hybrid_sharded(self._h, peer._h, ptr(q), ptr(t), ptr(w), t.shape[0],
                                                     float(alpha), int(pool_k), ptr(bs), ptr(rows), ptr(fused),
                                                     ptr(dense), ptr(sraw), ptr(present), C.byref(n)))
        else:
            check(lib.cqs_b200_search_hybrid(self._h, ptr(q), ptr(t), ptr(w), t.shape[0], float(alpha),
                                             int(pool_k), ptr(bs), ptr(rows), ptr(fused), ptr(dense),
                                             ptr(sraw), ptr(present), C.byref(n)))
        m = n.value
        return dict(rows=rows[:m].copy(), fused=fused[:m].copy(), dense=dense[:m].copy(),
                    sparse_raw=sraw[:m].copy(), present=present[:m].copy())


def rrf_fuse_n(ranked_lists, limit: int, k: float = 60.0, device: int = 0):
    """rrf_fuse_n over lists of row ids (src/search/scoring/fusion.rs:36-68) -> (ids, scores)."""
    lens = np.asarray([len(l) for l in ranked_lists], np.uint32)
    ids = np.ascontiguousarray(np.concatenate([np.asarray(l, np.uint64) for l in ranked_lists])
                               if len(ranked_lists) else np.empty(0, np.uint64))
    cap = max(min(int(limit), int(lens.sum())), 1)
    out_ids = np.empty(cap, np.uint64)
    out_sc = np.empty(cap, np.float32)
    n = C.c_uint32(0)
    check(lib.cqs_b200_rrf_fuse(device, ptr(ids), ptr(lens), len(ranked_lists), float(k), int(limit),
                                ptr(out_ids), ptr(out_sc), C.byref(n)))
    return out_ids[: n.value].copy(), out_sc[: n.value].copy()


def fuse_pools(dense_rows, dense_scores, sparse_rows, sparse_scores, alpha: float, pool_k: int, device: int = 0):
    """a11 on caller-supplied pools (src/search/query.rs:914-1005)."""
    dr = np.ascontiguousarray(dense_rows, dtype=np.uint64)
    ds = np.ascontiguousarray(dense_scores, dtype=np.float32)
    sr = np.ascontiguousarray(sparse_rows, dtype=np.uint64)
    ss = np.ascontiguousarray(sparse_scores, dtype=np.float32)
    cap = max(min(int(pool_k), dr.shape[0] + sr.shape[0]), 1)
    rows = np.empty(cap, np.uint64)
    fused = np.empty(cap, np.float32)
    dense = np.empty(cap, np.float32)
    sraw = np.empty(cap, np.float32)
    present = np.empty(cap, np.uint8)
    n = C.c_uint32(0)
    check(lib.cqs_b200_fuse_pools(device, ptr(dr), ptr(ds), dr.shape[0], ptr(sr), ptr(ss), sr.shape[0],
                                  float(alpha), int(pool_k), ptr(rows), ptr(fused), ptr(dense), ptr(sraw),
                                  ptr(present), C.byref(n)))
    m = n.value
    return dict(rows=rows[:m].copy(), fused=fused[:m].copy(), dense=dense[:m].copy(),
                sparse_raw=sraw[:m].copy(), present=present[:m].copy())
