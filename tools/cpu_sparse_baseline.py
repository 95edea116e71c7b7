"""CPU baseline of the SPLADE leg + alpha fusion: the oracle's C port of
SpladeIndex::search_with_filter (src/splade/index.rs:223-291) and the numpy restatement of the
fusion (src/search/query.rs:914-1005), timed on the host cores.  A reported baseline, not a target.

Stand-alone (no GPU needed):  python tools/cpu_sparse_baseline.py [docs] [mean_nnz] [q_nnz]
tools/bench_hybrid.py calls time_cpu_sparse() on the corpus it built so both numbers come from one run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle as CO
from oracle import cqs_oracle as O


def time_cpu_sparse(indptr, tok, w, vocab, queries, k=500, dense_pool=None, alpha=0.8):
    """queries: [(q_tok, q_w)].  Returns dict with build seconds, per-query ms (median), results of the last query."""
    n_docs = indptr.shape[0] - 1
    t0 = time.perf_counter()
    tptr, pdoc, pw = CO.csr_to_postings(indptr, tok, w, vocab)       # SpladeIndex::build, numpy stable argsort
    build_s = time.perf_counter() - t0
    lat, last = [], None
    for qt, qw in queries:
        t0 = time.perf_counter()
        last = CO.sparse_search(tptr, pdoc, pw, vocab, n_docs, qt, qw, k)
        lat.append(time.perf_counter() - t0)
    fuse_ms = None
    if dense_pool is not None and last is not None:
        sp = list(zip(last[0].tolist(), last[1].tolist()))
        t0 = time.perf_counter()
        O.fuse_hybrid(dense_pool, sp, alpha, k)
        fuse_ms = (time.perf_counter() - t0) * 1e3
    return {"build_s": build_s, "sparse_ms_p50": float(np.median(lat) * 1e3), "fuse_ms": fuse_ms,
            "cores": 1, "last": last}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    mean_nnz = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    q_nnz = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    vocab = 30522
    rng = np.random.default_rng(5)
    p = 1.0 / np.arange(1, vocab + 1) ** 1.1
    cdf = np.cumsum(p / p.sum())
    W = int(mean_nnz * 1.6)
    toks, ws, indptr = [], [], [0]
    for b in range(0, n, 20_000):
        m = min(20_000, n - b)
        t = np.sort(np.searchsorted(cdf, rng.random((m, W))).clip(0, vocab - 1), axis=1)
        keep = np.ones_like(t, bool); keep[:, 1:] = t[:, 1:] != t[:, :-1]
        ww = np.log1p(np.maximum(rng.normal(0.8, 0.5, (m, W)), 0.0)).astype(np.float32)
        keep &= ww > 0.01
        toks.append(t[keep].astype(np.uint32)); ws.append(ww[keep])
        indptr.extend((np.cumsum(keep.sum(1)) + indptr[-1]).tolist())
    indptr = np.asarray(indptr, np.uint64); tok = np.concatenate(toks); w = np.concatenate(ws)

    def query():
        t = np.unique(np.searchsorted(cdf, rng.random(q_nnz * 2)).clip(0, vocab - 1))[:q_nnz]
        return t.astype(np.uint32), np.log1p(np.maximum(rng.normal(0.8, 0.5, t.shape[0]), 0.02)).astype(np.float32)

    qs = [query() for _ in range(12)]
    dense_pool = [(int(i), float(s)) for i, s in zip(rng.choice(n, 500, replace=False), np.sort(rng.random(500))[::-1])]
    r = time_cpu_sparse(indptr, tok, w, vocab, qs, 500, dense_pool)
    print(f"{n} docs, {tok.shape[0]/1e6:.1f}M postings: CPU build {r['build_s']:.2f} s (numpy stable argsort), "
          f"sparse leg p50 {r['sparse_ms_p50']:.2f} ms on 1 core, fusion of two 500-pools {r['fuse_ms']:.2f} ms (numpy restatement)")
