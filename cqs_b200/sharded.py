"""Row-sharded corpora across processes: one process per GPU,
contiguous row blocks in ascending-chunk-id order, per-shard top-k, then the
lists are exchanged and merged with the reference's ordering rule (score desc
by f32 total order, row asc — candidate.rs:321-329).  The reference has no
multi-GPU path; this is the B200-side extension of VectorIndex::search for
corpora that do not fit (or should not sit on) one GPU.

Two transports for the exchange:
  * ``PeerGroup`` (the product path): mailboxes in peer HBM over NVLink; the scan
    kernel itself pushes its list to every peer, waits for theirs and merges —
    no collective call, no extra launch (csrc/peer.cuh).
  * ``torch.distributed`` all-gather + merge kernel (works on any backend; what the
    gloo CPU tests of the host logic use).

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


class PeerGroup:
    """cqs_b200_peer: this rank's mailbox + the mappings of every peer's (include/cqs_b200.h)."""

    def __init__(self, device: int, world: int, rank: int, max_elems: int = 0):
        self._h = C.c_void_p()
        self.device, self.world, self.rank = int(device), int(world), int(rank)
        check(lib.cqs_b200_peer_create(self.device, self.world, self.rank, int(max_elems), C.byref(self._h)))

    def handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        check(lib.cqs_b200_peer_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles) -> None:
        """handles: the 64-byte handle of every rank, in rank order (other PROCESSES)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        check(lib.cqs_b200_peer_connect(self._h, (C.c_uint8 * len(blob)).from_buffer_copy(blob)))

    @staticmethod
    def connect_local(groups) -> None:
        """All ranks live in THIS process (one group per device or per emulated shard)."""
        arr = (C.c_void_p * len(groups))(*[g._h for g in groups])
        check(lib.cqs_b200_peer_connect_local(arr, len(groups)))

    @classmethod
    def from_dist(cls, dist, device: int, max_elems: int = 0, strict: bool = True) -> Optional["PeerGroup"]:
        """One process per GPU under torch.distributed: create, swap handles, connect.

        Every step is followed by an agreement (an all-gather of per-rank success) so that a failure on
        one rank — no peer access, CUDA IPC unavailable in the container — never leaves the others stuck
        in a collective.  strict: raise on failure; otherwise return None on EVERY rank (the caller
        falls back to the all-gather transport)."""
        world = dist.get_world_size()

        def agree(ok: bool, why: str):
            flags = [None] * world
            dist.all_gather_object(flags, (bool(ok), why))
            bad = [f"rank {r}: {w}" for r, (o, w) in enumerate(flags) if not o]
            return (not bad), "; ".join(bad)

        g, why = None, ""
        try:
            g = cls(device, world, dist.get_rank(), max_elems)
        except Exception as e:  # noqa: BLE001 - reported to all ranks
            why = repr(e)
        ok, msg = agree(g is not None, why)
        if ok and world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, g.handle())
            why = ""
            try:
                g.connect(handles)
            except Exception as e:  # noqa: BLE001
                why = repr(e)
            ok, msg = agree(not why, why)
        if not ok:
            if g is not None:
                g.close()
            if strict:
                raise RuntimeError("peer group setup failed: " + msg)
            return None
        return g

    def set_timeout_ms(self, ms: int) -> None:
        check(lib.cqs_b200_peer_set_timeout_ms(self._h, int(ms)))

    def status(self) -> int:
        rc = lib.cqs_b200_peer_status(self._h)
        if rc < 0:
            check(rc)
        return rc

    def close(self) -> None:
        if self._h:
            lib.cqs_b200_peer_destroy(self._h)
            self._h = C.c_void_p()


def search_sharded(index, peer: PeerGroup, query: np.ndarray, k: int, bitset: Optional[np.ndarray] = None):
    """VectorIndex::search over the whole sharded corpus (host in, GLOBAL host top-k out)."""
    q = np.ascontiguousarray(query, np.float32)
    rows = np.empty(k, np.uint64)
    sc = np.empty(k, np.float32)
    n = C.c_uint32(0)
    check(lib.cqs_b200_search_sharded(index._h, peer._h, q.ctypes.data_as(C.c_void_p), k,
                                      None if bitset is None else bitset.ctypes.data_as(C.c_void_p),
                                      rows.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), C.byref(n)))
    return rows[:n.value], sc[:n.value]


def search_batch_sharded(index, peer: PeerGroup, queries: np.ndarray, k: int, bitset: Optional[np.ndarray] = None):
    """cqs_b200_search_batch over the sharded corpus: (rows [nq][k], scores [nq][k], n [nq])."""
    q = np.ascontiguousarray(queries, np.float32)
    nq = q.shape[0]
    rows = np.full((nq, k), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    sc = np.full((nq, k), -np.inf, np.float32)
    n = np.zeros(nq, np.uint32)
    check(lib.cqs_b200_search_batch_sharded(index._h, peer._h, q.ctypes.data_as(C.c_void_p), nq, k,
                                            None if bitset is None else bitset.ctypes.data_as(C.c_void_p),
                                            rows.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                            n.ctypes.data_as(C.c_void_p)))
    return rows, sc, n


class ShardedSearcher:
    """search_batch over a row-sharded corpus: local fused scan+top-k per query, then
    all_gather_into_tensor(scores) + all_gather_into_tensor(rows) and one merge kernel."""

    def __init__(self, index, dist=None, device=None, max_queries: int = 64, k: int = 20,
                 peer: Optional[PeerGroup] = None):
        import torch
        self.torch = torch
        self.index = index
        self.dist = dist
        self.peer = peer
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
        if self.peer is not None and self.world > 1:
            # product path: every launch scans, exchanges over NVLink and merges
            for i in range(nq):
                check(lib.cqs_b200_search_sharded_device(
                    self.index._h, self.peer._h, C.c_void_p(d_queries.data_ptr() + i * dim * 4), self.k, None,
                    C.c_void_p(self.m_sc.data_ptr() + i * self.k * 4),
                    C.c_void_p(self.m_rw.data_ptr() + i * self.k * 8),
                    C.c_void_p(self.m_n.data_ptr() + i * 4), stream))
            return self.m_sc[:nq], self.m_rw[:nq], self.m_n[:nq]
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
