"""CPU tests of bench.py's checker plumbing (no GPU): the block-fed CPU oracle, the list comparison
with its near-tie rule, recall, the shard split and the CLI contract of the two arms."""
import numpy as np

import bench as B
from oracle import cqs_oracle as O


def test_block_oracle_equals_oracle_topk_over_the_whole_corpus():
    rows = O.fast_unit_rows(25_000, 768, seed=3, clustered=True)
    rows[100] = rows[20_000]                                    # exact duplicate across blocks: tie by row asc
    q = B.make_queries(5, 9, "clustered")
    q[0] = rows[100]
    orc = B.BlockOracle(q, 20, threads=2)
    for b in range(0, 25_000, 6_000):                           # ragged last block
        orc.feed(rows[b:b + 6_000], 1_000_000 + b)
    assert orc.rows_fed == 25_000 and orc.cpu_s > 0
    for i in range(5):
        full = O.dense_scores(rows, q[i])
        r, s = O.topk_rows(full, 20)
        assert (orc.rows[i].astype(np.int64) - 1_000_000).tolist() == r.tolist()
        assert np.allclose(orc.scores[i], s, rtol=1e-6)
    assert orc.rows[0, 0] == 1_000_100 and orc.rows[0, 1] == 1_020_000


def test_compare_lists_identical_near_tie_and_real_mismatch():
    o_rows = np.arange(10, dtype=np.uint64)
    o_sc = np.linspace(0.9, 0.1, 10).astype(np.float32)
    assert B.compare_lists(o_rows, o_sc, o_rows, o_sc)[:2] == (True, True)
    # swap of two rows whose oracle scores differ by less than the tolerance: allowed, flagged as near-tie
    tie_sc = o_sc.copy(); tie_sc[4] = tie_sc[3] - np.float32(1e-7)
    g = o_rows.copy(); g[3], g[4] = g[4], g[3]
    ident, near, err = B.compare_lists(g, tie_sc, o_rows, tie_sc)
    assert (ident, near) == (False, True)
    # the same swap with well separated scores is a real mismatch
    ident, near, _ = B.compare_lists(g, o_sc, o_rows, o_sc)
    assert (ident, near) == (False, False)
    # a wrong score is caught by the relative error even when the ids agree
    bad = o_sc.copy(); bad[2] *= np.float32(1.001)
    assert B.compare_lists(o_rows, bad, o_rows, o_sc)[2] > 1e-5
    # empty slots of the oracle are ignored; a length mismatch is a failure
    pad_rows = np.concatenate([o_rows, np.full(3, B.EMPTY, np.uint64)])
    pad_sc = np.concatenate([o_sc, np.full(3, -np.inf, np.float32)])
    assert B.compare_lists(o_rows, o_sc, pad_rows, pad_sc)[0]
    assert B.compare_lists(o_rows[:5], o_sc[:5], o_rows, o_sc)[:2] == (False, False)
    s = B.parity_summary([(o_rows, o_sc), (g, tie_sc), (g, o_sc)], [o_rows] * 3, [o_sc, tie_sc, o_sc])
    assert s["ids_identical_queries"] == "1/3" and s["near_tie_only_mismatches"] == 1 and not s["ok"]


def test_recall_and_generators_are_deterministic():
    ex = np.asarray([[1, 2, 3, 4], [5, 6, 7, B.EMPTY]], np.uint64)
    got = [(np.asarray([1, 2, 3, 9], np.uint64), None), (np.asarray([5, 6, 7], np.uint64), None)]
    assert B.recall_at_k(got, ex) == 6 / 7
    a, b = B.make_queries(4, 7), B.make_queries(4, 7)
    assert np.array_equal(a, b) and np.allclose(np.linalg.norm(a, axis=1), 1, atol=1e-6)
    c = B.make_queries(64, 7, "clustered")
    cos = c @ B.centres_np().T
    assert (cos.max(axis=1) > 0.8).all()                        # every clustered query sits near one of the 256 centres


def test_cli_contract_and_shared_workload_string():
    p = B.build_parser()
    a = p.parse_args([])
    assert (a.gpus, a.impl, a.big_storage, a.records) == (1, "b200", "bf16+f32", "all")
    assert a.steps > 0 and a.warmup >= 3
    # both arms print the SAME config.workload string (the driver's same_config check)
    src = open(B.__file__).read()
    assert src.count('"workload": WORKLOAD') >= 1 and "head_wl = WORKLOAD" in src
    assert "BASELINE configs[1]" in B.WORKLOAD
    for name in ("single_k500", "hybrid_1M", "batch_10M", "sharded_single", "sharded_batch"):
        assert name in B.ALL_RECORDS
