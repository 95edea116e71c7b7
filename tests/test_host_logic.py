"""CPU-only checks: the C-ABI library loads and exports every symbol the header
declares, the host-side mirror logic (alpha table, candidate count, layout
choices) matches the oracle, and the oracle's C port agrees with the numpy one.
No compute calls that need a GPU."""
import ctypes
import os
import re

import numpy as np

from oracle import cqs_oracle as O, c_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cqs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cqs_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import cqs_b200.capi as capi
    names = _header_symbols()
    assert len(names) >= 25
    dll = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/cqs_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "ctypes binding and header disagree"
    assert capi.lib.cqs_b200_name() == b"B200"
    assert capi.lib.cqs_b200_max_k(None) == 1024


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    import cqs_b200
    if torch.cuda.is_available():
        return
    try:
        cqs_b200.B200Index(768)
    except cqs_b200.B200Error as e:
        assert e.code == cqs_b200.capi.ERR_CUDA and "no CPU fallback" in str(e)
    else:
        raise AssertionError("index creation must fail without a CUDA device")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cqs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libcqs_oracle", "oracle/", "cqs_oracle"):
                    assert needle not in txt, f"{f} references the oracle ({needle})"


def test_router_mirror_matches_oracle_table():
    import cqs_b200.router as R
    assert R.CATEGORIES == O.CATEGORIES and R.DEFAULT_ALPHA == O.DEFAULT_ALPHA
    envs = [{}, {"CQS_SPLADE_ALPHA": "0.3"}, {"CQS_SPLADE_ALPHA_STRUCTURAL": "0.9", "CQS_SPLADE_ALPHA": "0.2"},
            {"CQS_SPLADE_ALPHA": "NaN"}, {"CQS_SPLADE_ALPHA": "2"}, {"CQS_SPLADE_ALPHA": " 0.5"},
            {"CQS_SPLADE_ALPHA_UNKNOWN": "-1"}, {"CQS_SPLADE_ALPHA": "1e-1"}, {"CQS_SPLADE_ALPHA": "inf"}]
    for env in envs:
        for cat in R.CATEGORIES:
            for slot in (None, {"structural": 0.42, "unknown": 7.0}):
                assert R.resolve_splade_alpha(cat, env, slot) == O.resolve_splade_alpha(cat, env, slot)
    assert R.apply_centroid_floor(0.1, True) == O.apply_centroid_floor(0.1, True)


def test_hybrid_mirror_counts():
    import cqs_b200.hybrid as H
    assert H.candidate_count_for(5) == O.candidate_count_for(5, 500) == 500
    assert H.candidate_count_for(300) == 1500

    class Fake:
        def __init__(self, m): self.m = m
        def max_k(self): return self.m
    assert H.cap_k_to_backend(Fake(1024), 1500) == 1024 and H.cap_k_to_backend(Fake(None), 1500) == 1500


def test_centroid_loader(tmp_path):
    import json
    import cqs_b200.router as R
    p = tmp_path / "classifier_centroids.v1.json"
    p.write_text(json.dumps({"dim": 4, "categories": {"structural": {"centroid": [1, 0, 0, 0]},
                                                      "behavioral_search": {"centroid": [0, 1, 0, 0]},
                                                      "negation": {"centroid": [1, 2]}}}))
    clf = R.CentroidClassifier.load(str(p), env={})
    assert clf is not None and clf.names == ["structural", "behavioral"] and clf.threshold == float(np.float32(0.01))
    assert R.CentroidClassifier.load(str(p), env={"CQS_CENTROID_CLASSIFIER": "0"}) is None
    assert R.CentroidClassifier.load(str(tmp_path / "missing.json"), env={}) is None
    assert R.CentroidClassifier.load(str(p), env={"CQS_CENTROID_THRESHOLD": "0.05"}).threshold == float(np.float32(0.05))


def test_c_port_agrees_with_numpy_oracle():
    v, st = O.synth_vectors(40, 768)
    c, st2 = CO.synth_vectors(40, 768)
    assert st == st2 and np.array_equal(v.view(np.uint32), c.view(np.uint32))
    rows, _ = CO.synth_vectors(3000, 768)
    rows[10:14] = rows[500:504]
    rng = np.random.default_rng(0)
    for qi in range(5):
        q = rows[500 + qi] if qi < 2 else rows[int(rng.integers(0, 3000))]
        mask = rng.random(3000) > 0.5 if qi % 2 else None
        bs = None if mask is None else O.mask_to_bitset(mask)
        for k in (1, 20, 500):
            r1, s1 = O.brute_force_search(rows, q, k, mask)
            r2, s2 = CO.brute_force(rows, q, k, bs, use_f64=True)
            assert np.array_equal(r1, r2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
            r3, s3 = CO.brute_force(rows, q, k, bs, use_f64=False)   # SIMD f32 arm: within tolerance
            assert np.allclose(s3, s1, rtol=1e-5, atol=1e-7)
    rb, sb, nb = CO.brute_force_batch(rows, rows[:6], 10, use_f64=True, threads=3)
    for i in range(6):
        r1, s1 = O.brute_force_search(rows, rows[i], 10)
        assert nb[i] == 10 and np.array_equal(rb[i].astype(np.int64), r1)


def test_c_port_sparse_agrees_with_numpy_oracle():
    rng = np.random.default_rng(4)
    n_docs, vocab = 500, 80
    indptr, tok, w = [0], [], []
    for d in range(n_docs):
        t = np.sort(rng.choice(vocab, size=int(rng.integers(0, 15)), replace=False))
        tok += t.tolist(); w += rng.random(t.shape[0]).astype(np.float32).tolist(); indptr.append(len(tok))
    tptr, pdoc, pw = CO.csr_to_postings(indptr, tok, w, vocab)
    for trial in range(6):
        qt = rng.choice(vocab + 3, size=int(rng.integers(1, 12)), replace=False).astype(np.uint32)
        qw = rng.random(qt.shape[0]).astype(np.float32)
        mask = rng.random(n_docs) > 0.3 if trial % 2 else None
        bs = None if mask is None else O.mask_to_bitset(mask)
        r1, s1 = O.sparse_search_csr(indptr, tok, w, qt, qw, n_docs, 30, mask)
        r2, s2 = CO.sparse_search(tptr, pdoc, pw, vocab, n_docs, qt, qw, 30, bs)
        assert np.array_equal(r1, r2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))


def test_hnsw_baseline_port_reaches_reference_like_recall():
    """oracle/hnsw_baseline.c is a timing baseline (approximate, never a parity oracle); this
    only guards that the restated graph works: tier parameters of src/hnsw/mod.rs:104-112,
    self-match reachable (src/hnsw/mod.rs:962-984), high recall on clustered data."""
    rows = O.fast_unit_rows(2500, 128, seed=7, clustered=True)
    h = CO.Hnsw(rows, threads=4)
    assert (h.M, h.efC, h.efS) == (16, 100, 50)
    ids, sc, n, lat = h.search(rows[:100], 10, threads=2)
    assert (ids[:, 0] == np.arange(100)).mean() > 0.97 and np.allclose(sc[ids[:, 0] == np.arange(100), 0], 1.0, atol=1e-4)
    hits = 0
    for i in range(100):
        r, _ = CO.brute_force(rows, rows[i], 10)
        hits += len(set(r.tolist()) & set(ids[i, :n[i]].tolist()))
    assert hits / 1000 > 0.9
    assert np.all(np.diff(sc, axis=1) <= 1e-6)      # sorted descending
    h.close()
