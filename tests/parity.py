"""Shared comparison helpers for the GPU parity tests.

Contract (BASELINE.json north_star): f32 scores within 1e-5 relative of the
oracle (f64-accumulated dot rounded to f32); top-k ids identical wherever the
oracle scores involved are separated by more than that tolerance.
"""
import numpy as np

REL_TOL = 1e-5      # the north star's stated tolerance, written here once
# Absolute floor for scores near zero: a dot of unit vectors has sum|a_i*b_i| <= 1, so
# any f32 accumulation (ours, or simsimd's in the reference) carries ~1e-7 absolute
# error; a purely relative bound is meaningless where the terms cancel (|score| < 0.01).
# Ranked results (|score| >= 0.01) are held to the relative bound alone.
ABS_FLOOR = 1e-7


def assert_topk_parity(gpu_rows, gpu_scores, ora_rows, ora_scores, full_oracle_scores=None, row_base=0):
    gpu_rows = np.asarray(gpu_rows).astype(np.int64) - row_base
    ora_rows = np.asarray(ora_rows).astype(np.int64)
    gpu_scores = np.asarray(gpu_scores, np.float32)
    ora_scores = np.asarray(ora_scores, np.float32)
    assert gpu_rows.shape[0] == ora_rows.shape[0], (gpu_rows.shape, ora_rows.shape)
    n = gpu_rows.shape[0]
    if n == 0:
        return
    assert np.all(np.isfinite(gpu_scores))
    # sorted descending, unique rows
    assert np.all(gpu_scores[:-1] >= gpu_scores[1:]), "GPU result not sorted by score desc"
    assert np.unique(gpu_rows).shape[0] == n, "duplicate rows in GPU result"
    tol = np.maximum(REL_TOL * np.abs(ora_scores.astype(np.float64)), ABS_FLOOR)
    diff = np.abs(gpu_scores.astype(np.float64) - ora_scores.astype(np.float64))
    assert np.all(diff <= tol), f"score mismatch: max rel {np.max(diff / (np.abs(ora_scores) + 1e-30))}"
    mism = np.nonzero(gpu_rows != ora_rows)[0]
    for i in mism:
        # a different id is only allowed inside a near-tie of ORACLE scores
        assert full_oracle_scores is not None, f"row mismatch at rank {i}: {gpu_rows[i]} vs {ora_rows[i]}"
        a = float(full_oracle_scores[gpu_rows[i]])
        b = float(ora_scores[i])
        assert abs(a - b) <= 2 * max(REL_TOL * abs(b), ABS_FLOOR), \
            f"rank {i}: GPU row {gpu_rows[i]} (oracle score {a}) vs oracle row {ora_rows[i]} ({b}) is not a near-tie"
    # exact ties must be ordered by row ascending
    for i in range(n - 1):
        if gpu_scores[i] == gpu_scores[i + 1]:
            assert gpu_rows[i] < gpu_rows[i + 1], f"tie at rank {i} not ordered by row asc"


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)
