"""ctypes wrapper of oracle/cqs_oracle.c (TEST INFRASTRUCTURE ONLY — the C port
of the reference's CPU path; used as checker at sizes numpy handles slowly and
as the timed CPU baseline in bench.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libcqs_oracle.so")
vp = C.c_void_p


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "cqs_oracle.c"), os.path.join(_HERE, "hnsw_baseline.c")]
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _PATH


def _load():
    if not os.path.exists(_PATH):
        build()
    dll = C.CDLL(_PATH)
    dll.oracle_synth_vectors.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), vp]
    dll.oracle_dense_scores.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp, C.c_int]
    dll.oracle_brute_force.argtypes = [vp, C.c_uint64, C.c_uint32, vp, C.c_uint32, vp, C.c_int, vp, vp, vp]
    dll.oracle_brute_force.restype = C.c_int
    dll.oracle_brute_force_batch.argtypes = [vp, C.c_uint64, C.c_uint32, vp, C.c_uint32, C.c_uint32,
                                             C.c_int, C.c_int, vp, vp, vp]
    dll.oracle_brute_force_batch.restype = C.c_int
    dll.oracle_num_threads.restype = C.c_int
    dll.oracle_sparse_search.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp, vp, C.c_uint32,
                                         C.c_uint32, vp, vp, vp, vp]
    dll.oracle_sparse_search.restype = C.c_int
    dll.hnsw_tier.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    dll.hnsw_build.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64]
    dll.hnsw_build.restype = vp
    dll.hnsw_free.argtypes = [vp]
    dll.hnsw_search_batch.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, vp, vp, vp]
    return dll


_dll = None


def dll():
    global _dll
    if _dll is None:
        _dll = _load()
    return _dll


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def synth_vectors(n: int, dim: int = 768, state: int = 0x9E3779B97F4A7C15):
    out = np.empty((n, dim), np.float32)
    st = C.c_uint64(state)
    dll().oracle_synth_vectors(n, dim, C.byref(st), _p(out))
    return out, st.value


def dense_scores(rows: np.ndarray, query: np.ndarray, use_f64: bool = True) -> np.ndarray:
    rows = np.ascontiguousarray(rows, np.float32)
    q = np.ascontiguousarray(query, np.float32)
    out = np.empty(rows.shape[0], np.float32)
    dll().oracle_dense_scores(_p(rows), rows.shape[0], rows.shape[1], _p(q), _p(out), int(use_f64))
    return out


def brute_force(rows, query, k, bitset=None, use_f64=True):
    rows = np.ascontiguousarray(rows, np.float32)
    q = np.ascontiguousarray(query, np.float32)
    o_r = np.empty(max(k, 1), np.uint64)
    o_s = np.empty(max(k, 1), np.float32)
    n = C.c_uint32(0)
    bs = None if bitset is None else np.ascontiguousarray(bitset, np.uint32)
    rc = dll().oracle_brute_force(_p(rows), rows.shape[0], rows.shape[1], _p(q), k, _p(bs), int(use_f64),
                                  _p(o_r), _p(o_s), C.byref(n))
    assert rc == 0
    return o_r[: n.value].astype(np.int64), o_s[: n.value].copy()


def brute_force_batch(rows, queries, k, use_f64=False, threads=1):
    rows = np.ascontiguousarray(rows, np.float32)
    q = np.ascontiguousarray(queries, np.float32)
    nq = q.shape[0]
    o_r = np.empty((nq, max(k, 1)), np.uint64)
    o_s = np.empty((nq, max(k, 1)), np.float32)
    n = np.zeros(nq, np.uint32)
    rc = dll().oracle_brute_force_batch(_p(rows), rows.shape[0], rows.shape[1], _p(q), nq, k, int(use_f64),
                                        int(threads), _p(o_r), _p(o_s), _p(n))
    assert rc == 0
    return o_r, o_s, n


def num_threads() -> int:
    return int(dll().oracle_num_threads())


def csr_to_postings(indptr, tok, w, vocab):
    """SpladeIndex::build (src/splade/index.rs:191-211): token -> [(doc, w)] in doc order."""
    indptr = np.asarray(indptr, np.int64)
    tok = np.asarray(tok, np.uint32)
    w = np.asarray(w, np.float32)
    n = indptr.shape[0] - 1
    doc_of = np.repeat(np.arange(n, dtype=np.uint32), np.diff(indptr))
    order = np.argsort(tok, kind="stable")
    tptr = np.zeros(vocab + 1, np.uint64)
    np.add.at(tptr, tok.astype(np.int64) + 1, 1)
    tptr = np.cumsum(tptr).astype(np.uint64)
    return tptr, np.ascontiguousarray(doc_of[order]), np.ascontiguousarray(w[order])


def sparse_search(tptr, doc, w, vocab, n_docs, q_tok, q_w, k, bitset=None):
    qt = np.ascontiguousarray(q_tok, np.uint32)
    qw = np.ascontiguousarray(q_w, np.float32)
    o_r = np.empty(max(k, 1), np.uint64)
    o_s = np.empty(max(k, 1), np.float32)
    n = C.c_uint32(0)
    bs = None if bitset is None else np.ascontiguousarray(bitset, np.uint32)
    rc = dll().oracle_sparse_search(_p(tptr), _p(doc), _p(w), vocab, n_docs, _p(qt), _p(qw), qt.shape[0], k,
                                    _p(bs), _p(o_r), _p(o_s), C.byref(n))
    assert rc == 0
    return o_r[: n.value].astype(np.int64), o_s[: n.value].copy()


class Hnsw:
    """CPU HNSW restatement with the reference's tier parameters (oracle/hnsw_baseline.c):
    a TIMING baseline (approximate; never a parity oracle)."""

    def __init__(self, rows: np.ndarray, threads: int = 0, seed: int = 0):
        self.rows = np.ascontiguousarray(rows, np.float32)   # must outlive the graph
        n, dim = self.rows.shape
        M, efC, efS = C.c_uint32(), C.c_uint32(), C.c_uint32()
        dll().hnsw_tier(n, C.byref(M), C.byref(efC), C.byref(efS))
        self.M, self.efC, self.efS = M.value, efC.value, efS.value
        self.threads = threads or num_threads()
        self._h = dll().hnsw_build(_p(self.rows), n, dim, self.M, self.efC, self.threads, seed)

    def search(self, queries: np.ndarray, k: int, threads: int = 1):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.zeros((nq, k), np.uint32)
        sc = np.zeros((nq, k), np.float32)
        n = np.zeros(nq, np.uint32)
        lat = np.zeros(nq, np.float64)
        dll().hnsw_search_batch(self._h, _p(q), nq, k, self.efS, threads, _p(ids), _p(sc), _p(n), _p(lat))
        return ids, sc, n, lat

    def close(self):
        if self._h:
            dll().hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
