"""Regenerates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

The reference is Rust and cannot be built in this image (no cargo), so no output of the reference
itself can be recorded.  What is recorded:
  * reference_literals.json — the literal input/expected-output values of the reference's own unit
    tests for this path (file:line cited per entry), transcribed by hand from /root/reference;
  * config1_top20.json — BASELINE configs[0]: the reference's synthetic generator
    (examples/exp_level_scale.rs:200-224, xorshift64 seed 0x9E3779B97F4A7C15; 17,523 rows then 218
    queries from the same stream) with the exact top-20 computed by the f64 oracle
    (src/math.rs:17-22 fallback arithmetic, candidate.rs:303-329 ordering) for 8 of the queries.
    It pins the oracle, the C port and the CUDA library against drift; ids must match exactly,
    scores within the north star's 1e-5 relative."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as CO      # noqa: E402
from oracle import cqs_oracle as O     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def config1():
    rows, st = CO.synth_vectors(17523, 768)
    q, _ = CO.synth_vectors(218, 768, st)
    out = {"generator": "examples/exp_level_scale.rs:200-224 (xorshift64, seed 0x9E3779B97F4A7C15)",
           "n_rows": 17523, "dim": 768, "n_queries": 218, "k": 20, "queries": []}
    for qi in (0, 1, 7, 50, 100, 150, 200, 217):
        r, s = O.brute_force_search(rows, q[qi], 20)
        out["queries"].append({"query_index": qi, "rows": [int(x) for x in r],
                               "score_bits": [int(x) for x in np.asarray(s, np.float32).view(np.uint32)]})
    out["row0_first4_bits"] = [int(x) for x in rows[0, :4].view(np.uint32)]
    return out


LITERALS = {
    "splade_fixture": {
        "cite": "src/splade/index.rs:1114-1241 (test_build_and_search, test_dot_product_correct)",
        "index": {"a": [[1, 0.5], [2, 0.3], [3, 0.8]], "b": [[1, 0.7], [4, 0.6]], "c": [[2, 0.9], [3, 0.1], [5, 0.4]]},
        "query": [[1, 1.0], [2, 1.0]],
        "expected": [["c", 0.9], ["a", 0.8], ["b", 0.7]], "tolerance": 1e-5},
    "hybrid_legs": {
        "cite": "tests/search_test.rs:549-791 (search_hybrid_legs_dense_raw_cosine_sparse_minmax_and_raw_dot)",
        "dense_pool": [["A", 0.92], ["B", 0.61]], "sparse_pool": [["B", 0.8], ["C", 0.5]], "alpha": 0.5,
        "minmax": {"B": 1.0, "C": 0.625}, "order": ["B", "A", "C"],
        "fused_from_formula": {"A": 0.46, "B": 0.805, "C": 0.3125}},
    "alpha_table": {
        "cite": "src/search/router.rs:126-175, tests/router_test.rs:92-226",
        "values": {"identifier_lookup": 0.85, "structural": 0.60, "behavioral": 1.00, "conceptual": 0.80,
                   "multi_step": 0.10, "negation": 0.80, "type_filtered": 0.00, "cross_language": 0.70,
                   "unknown": 0.80}},
    "rrf": {"cite": "src/search/scoring/fusion.rs:208-331", "k": 60,
            "rank1_in_both_lists": 2.0 / 61.0, "rank2_single_list": 1.0 / 62.0},
    "heap_ties": {"cite": "src/search/scoring/candidate.rs:588-707", "capacity": 2,
                  "pushes": [["c", 0.5], ["b", 0.5], ["a", 0.5]], "expected": ["a", "b"]},
}

if __name__ == "__main__":
    json.dump(config1(), open(os.path.join(HERE, "config1_top20.json"), "w"), indent=1)
    json.dump(LITERALS, open(os.path.join(HERE, "reference_literals.json"), "w"), indent=1)
    print("written")
