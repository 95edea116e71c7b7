"""Row-sharded corpora across processes: one process per GPU,
contiguous row blocks in ascending-chunk-id order, per-shard top-k left on the
device, ONE small all-gather, deterministic merge with the reference's ordering
rule (score desc by f32 total order, row asc — candidate.rs:321-329).  The
reference has no multi-GPU path; this is the B200-side extension of
VectorIndex::search for corpora that do not fit (or should not sit on) one GPU.

torch is used for the plumbing only (device buffers, torch.distributed)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .capi import lib, check


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of rank `rank`: (first_row, n_rows); blocks differ by <= 1 chunk."""
    per = (n_total + world - 1) // world
    row0 = min(rank * per, n_total)
    return row0, min(per, n_total - row0)


def merge_topk_host(scores: np.ndarray, rows: np.ndarray, k: int):
    """Host reference of the merge rule for gathered lists [n_lists][k]; empty slots
    carry row == UINT64_MAX (or a non-finite score).  Returns (scores, rows)."""
    s = np.asarray(scores, np.float32).reshape(-1)
    r = np.asarray(rows).astype(np.uint64).reshape(-1)
    ok = np.isfinite(s) & (r != np.uint64(0xFFFFFFFFFFFFFFFF))
    s, r = s[ok], r[ok]
    u = s.view(np.uint32)
    key = np.where((u & np.uint32(0x80000000)) != 0, ~u, u | np.uint32(0x80000000)).astype(np.int64)
    order = np.lexsort((r, -key))[:k]
    return s[order], r[order]


class ShardedSearcher:
    """search_batch over a row-sharded corpus: local fused scan+top-k per query, then
    all_gather_into_tensor(scores) + all_gather_into_tensor(rows) and one merge kernel."""

    def __init__(self, index, dist=None, device=None, max_queries: int = 64, k: int = 20):
        import torch
        self.torch = torch
        self.index = index
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.k = k
        self.Q = max_queries
        t = torch
        self.d_sc = t.empty((self.Q, k), dtype=t.float32, device=self.device)
        self.d_rw = t.empty((self.Q, k), dtype=t.int64, device=self.device)
        self.d_n = t.empty((self.Q,), dtype=t.int32, device=self.device)
        if self.world > 1:
            self.g_sc = t.empty((self.world, self.Q, k), dtype=t.float32, device=self.device)
            self.g_rw = t.empty((self.world, self.Q, k), dtype=t.int64, device=self.device)
            self.m_sc = t.empty((self.Q, k), dtype=t.float32, device=self.device)
            self.m_rw = t.empty((self.Q, k), dtype=t.int64, device=self.device)
            self.m_n = t.empty((self.Q,), dtype=t.int32, device=self.device)

    def search_device(self, d_queries, nq: int):
        """d_queries: f32 [nq][dim] CUDA tensor, identical on every rank.  Asynchronous on the
        current torch stream (must not be the legacy default stream).  Returns device
        tensors (scores [nq][k], rows [nq][k], n [nq]) holding the GLOBAL top-k on every rank."""
        t = self.torch
        assert nq <= self.Q
        stream = C.c_void_p(t.cuda.current_stream().cuda_stream)
        dim = self.index.dim()
        for i in range(nq):
            check(lib.cqs_b200_search_device(
                self.index._h, C.c_void_p(d_queries.data_ptr() + i * dim * 4), self.k, None,
                C.c_void_p(self.d_sc.data_ptr() + i * self.k * 4),
                C.c_void_p(self.d_rw.data_ptr() + i * self.k * 8),
                C.c_void_p(self.d_n.data_ptr() + i * 4), stream))
        if self.world == 1:
            return self.d_sc[:nq], self.d_rw[:nq], self.d_n[:nq]
        self.dist.all_gather_into_tensor(self.g_sc, self.d_sc)
        self.dist.all_gather_into_tensor(self.g_rw, self.d_rw)
        check(lib.cqs_b200_merge_topk_device(
            self.device.index or 0, C.c_void_p(self.g_sc.data_ptr()), C.c_void_p(self.g_rw.data_ptr()),
            self.world, self.Q, self.k, C.c_void_p(self.m_sc.data_ptr()), C.c_void_p(self.m_rw.data_ptr()),
            C.c_void_p(self.m_n.data_ptr()), stream))
        return self.m_sc[:nq], self.m_rw[:nq], self.m_n[:nq]
