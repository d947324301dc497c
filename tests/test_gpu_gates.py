"""Attribute gates on the GPU (rr_gate_factors / rr_gate_fixed_bitmaps) against the reference's captured
outputs (tests/golden/gate_cases.json) and the oracle restatement; bit-exact (float32 factors, hit counts)."""
import json
import types

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

from oracle import primitives as P
from oracle.gates import build_gate_groups as oracle_groups, calculate_gate_factor
from tests.golden_worlds import make_gate_texts
from tests.parity import assert_ids_match_modulo_ties


@pytest.fixture(scope="module")
def g(golden_dir):
    import review_recommender_b200 as rr
    gc = json.loads((golden_dir / "gate_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    texts = make_gate_texts(z["emb"].shape[0], z["doc_offsets"], z["token_ids"])
    gi = rr.engine.GateIndex(texts, rr.drop_in.GATE_FIXED_GROUPS)
    return gc, z, texts, gi


@pytest.mark.parametrize("use_bitmaps", [True, False])
def test_factor_cases_bit_exact(g, use_bitmaps):
    import review_recommender_b200 as rr
    gc, z, texts, gi = g
    for pen in (0.5, 0.3, 1.0):
        cs = [c for c in gc["factor_cases"] if c["penalty"] == pen]
        cand = np.array([c["rows"] for c in cs], dtype=np.int64)
        groups = [rr.drop_in.build_gate_groups(gc["queries"][c["query_index"]]) for c in cs]
        gate, hits = gi.factors(groups, cand, pen, want_hits=True, use_bitmaps=use_bitmaps)
        gate, hits = gate.cpu().numpy(), hits.cpu().numpy()
        for i, c in enumerate(cs):
            np.testing.assert_array_equal(gate[i], np.float32(c["gate_f32"]))
            np.testing.assert_array_equal(hits[i], np.int32(c["hits"]))


def test_all_rows_against_oracle_and_invalid_candidates(g):
    """Every product row (long / multi-byte / upper-case texts included) for every golden query, plus -1 rows,
    a pattern longer than the staged overlap and many-group queries."""
    gc, z, texts, gi = g
    n = len(texts)
    queries = gc["queries"]
    groups = [oracle_groups(q) for q in queries]
    groups.append([{"t1"}, {"lorem"}, {"x" * 41 + " " + "x" * 41}, {"x" * 30 + " yellow"}, {"über"}, {"中 headphones"}])
    rows = np.concatenate([np.arange(n), [-1, n, -5]]).astype(np.int64)
    cand = np.tile(rows, (len(groups), 1))
    gate, hits = gi.factors(groups, cand, 0.5, want_hits=True)
    gate, hits = gate.cpu().numpy(), hits.cpu().numpy()
    for b, gr in enumerate(groups):
        want = [calculate_gate_factor(str(t)[:6000], gr, 0.5) for t in texts]
        np.testing.assert_array_equal(gate[b, :n], np.array([w[0] for w in want], dtype=np.float32))
        np.testing.assert_array_equal(hits[b, :n], np.array([w[1] for w in want], dtype=np.int32))
        assert (gate[b, n:] == 1.0).all()
    assert (hits[-1] > 0).any()


def test_drivers_with_gates_reproduce_reference(g, golden_dir):
    import review_recommender_b200 as rr
    gc, z, texts, _ = g
    n = len(texts)
    skus = rr.synth.skus(n)
    corpus = rr.synth.corpus_as_lists(z["doc_offsets"], z["token_ids"])
    perm = z["bm25_perm"]
    meta = pd.DataFrame({"sku": skus, "n_reviews": z["n_reviews"], "avg_stars": z["avg_stars"], "agg_text": texts})
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    table = {q: z["queries"][i % len(z["queries"])] for i, q in enumerate(gc["queries"])}
    eng = rr.drop_in.SearchEngine(meta, Vn, [corpus[i] for i in perm], [skus[i] for i in perm], encode=lambda q: table[q])
    assert eng.gate_ix is not None
    for c in gc["cases"]:
        if c["driver"] == "streamlit":
            top, _, dbg = eng.run_search(c["query"], 10, 0, 0.55, 0.20, 0.0, 0.20, 0.0, 20.0, False, 0, 8, c["penalty"])
        else:
            args = types.SimpleNamespace(query=c["query"], k=10, rerank_k=0, w_dense=0.55, w_bm25=0.15, w_rerank=0.0,
                                         w_prior=0.10, w_best=0.0, prior_C=20.0, gate_penalty=c["penalty"])
            top = eng.search(args)
        ref_gate = dict(zip(c["pool_skus"], c["pool_gate"]))
        np.testing.assert_array_equal(top["_gate"].values, np.float32([ref_gate[s] for s in top["sku"]]))
        order = np.argsort(-np.float32(c["pool_final"]), kind="stable")[:len(c["top_skus"])]
        ref_final = np.float32(c["pool_final"])[order]
        np.testing.assert_allclose(top["_final"].values, ref_final, rtol=1e-5, atol=1e-7)
        assert_ids_match_modulo_ties([int(s[3:]) for s in top["sku"]], top["_final"].values,
                                     [int(s[3:]) for s in c["top_skus"]], ref_final, 2e-6,
                                     f"{c['driver']} q{c['query_index']} pen{c['penalty']}")
