"""Committed fixtures under tests/golden/ (see tests/golden/make_golden.py for what they are and
how they were produced).  CPU: the numpy oracle and its C port reproduce them.  GPU: the CUDA
library reproduces them through the C ABI — ids exactly, scores within the north star's 1e-5."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import cqs_oracle as O

f32 = np.float32
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def config1():
    g = _load("config1_top20.json")
    rows, st = CO.synth_vectors(g["n_rows"], g["dim"])
    q, _ = CO.synth_vectors(g["n_queries"], g["dim"], st)
    assert [int(x) for x in rows[0, :4].view(np.uint32)] == g["row0_first4_bits"]
    return g, rows, q


def test_oracle_and_c_port_reproduce_config1_golden(config1):
    g, rows, q = config1
    for e in g["queries"]:
        r, s = O.brute_force_search(rows, q[e["query_index"]], g["k"])
        assert [int(x) for x in r] == e["rows"]
        assert [int(x) for x in np.asarray(s, f32).view(np.uint32)] == e["score_bits"]
    qi = [e["query_index"] for e in g["queries"]]
    c_r, c_s, c_n = CO.brute_force_batch(rows, q[qi], g["k"], use_f64=True, threads=2)
    for j, e in enumerate(g["queries"]):
        assert int(c_n[j]) == g["k"] and [int(x) for x in c_r[j]] == e["rows"]


def test_oracle_reproduces_reference_literals():
    lit = _load("reference_literals.json")
    sp = lit["splade_fixture"]
    ix = O.SpladeIndex([(k, [tuple(p) for p in v]) for k, v in sp["index"].items()])
    got = ix.search([tuple(p) for p in sp["query"]], 10)
    assert [i for i, _ in got] == [i for i, _ in sp["expected"]]
    for (_, s), (_, want) in zip(got, sp["expected"]):
        assert abs(float(s) - want) < sp["tolerance"]
    hl = lit["hybrid_legs"]
    fused = O.fuse_hybrid([tuple(x) for x in hl["dense_pool"]], [tuple(x) for x in hl["sparse_pool"]], hl["alpha"], 10)
    assert [x["id"] for x in fused] == hl["order"]
    for x in fused:
        assert abs(float(x["fused"]) - hl["fused_from_formula"][x["id"]]) < 1e-6
        if x["id"] in hl["minmax"]:
            assert abs(float(x["sparse_norm"]) - hl["minmax"][x["id"]]) < 1e-6
    for cat, a in lit["alpha_table"]["values"].items():
        assert float(O.resolve_splade_alpha(cat, env={})) == pytest.approx(a, abs=1e-7)
    rr = lit["rrf"]
    out = dict(O.rrf_fuse_n([["x", "y"], ["x"]], 10, rr["k"]))
    assert abs(float(out["x"]) - rr["rank1_in_both_lists"]) < 1e-6 and abs(float(out["y"]) - rr["rank2_single_list"]) < 1e-6
    hp = lit["heap_ties"]
    h = O.BoundedScoreHeap(hp["capacity"])
    for i, s in hp["pushes"]:
        h.push(i, s)
    assert [i for i, _ in h.into_sorted_vec()] == hp["expected"]


@pytest.mark.gpu
@pytest.mark.parametrize("entry", ["search", "search_batch", "search_sharded_world1"])
def test_cuda_library_reproduces_config1_golden(config1, entry):
    import cqs_b200
    from cqs_b200.sharded import PeerGroup, search_sharded
    g, rows, q = config1
    ix = cqs_b200.B200Index(g["dim"])
    ix.append(None, rows); ix.finalize()
    qi = [e["query_index"] for e in g["queries"]]
    if entry == "search":
        got = [ix.search_rows(q[i], g["k"]) for i in qi]
    elif entry == "search_batch":
        r, s, n = ix.search_batch_rows(q[qi], g["k"])
        got = [(r[j, :n[j]], s[j, :n[j]]) for j in range(len(qi))]
    else:
        pg = PeerGroup(0, 1, 0)
        got = [search_sharded(ix, pg, q[i], g["k"]) for i in qi]
        pg.close()
    for (r, s), e in zip(got, g["queries"]):
        assert [int(x) for x in r] == e["rows"]
        want = np.asarray(e["score_bits"], np.uint32).view(f32)
        assert np.allclose(s, want, rtol=1e-5, atol=1e-7)      # north star: 1e-5 relative
    ix.close()


@pytest.mark.gpu
def test_cuda_library_reproduces_reference_literals():
    import cqs_b200
    lit = _load("reference_literals.json")
    sp = lit["splade_fixture"]
    ids = sorted(sp["index"])
    ix = cqs_b200.B200Index.build(ids, O.fast_unit_rows(len(ids), 8, seed=0))
    spi = cqs_b200.SpladeIndex(ix, [(k, [tuple(p) for p in sp["index"][k]]) for k in ids])
    got = spi.search([tuple(p) for p in sp["query"]], 10)
    assert [x.id for x in got] == [i for i, _ in sp["expected"]]
    for x, (_, want) in zip(got, sp["expected"]):
        assert abs(x.score - want) < sp["tolerance"]
    rr = lit["rrf"]
    from cqs_b200.index import rrf_fuse_n
    r_ids, r_sc = rrf_fuse_n([[10, 11], [10]], 10, rr["k"])
    out = dict(zip([int(i) for i in r_ids], [float(s) for s in r_sc]))
    assert abs(out[10] - rr["rank1_in_both_lists"]) < 1e-6 and abs(out[11] - rr["rank2_single_list"]) < 1e-6
    ix.close()
