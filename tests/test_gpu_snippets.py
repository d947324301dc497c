"""Best-review scoring on the GPU (rr_best_review_scores + K4's raw-best min-max) against the
reference's captured outputs (tests/golden/snippet_cases.json) and the oracle restatement."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.snippets import best_review_snippets
from tests import snippet_world
from tests.parity import assert_ids_match_modulo_ties

SIM_ATOL = 1e-6      # fp32 summation order (BLAS sgemv vs the canonical dot product)


@pytest.fixture(scope="module")
def w(golden_dir):
    return snippet_world.load(golden_dir)


@pytest.fixture(scope="module")
def eng(w):
    import review_recommender_b200 as rr
    return rr.drop_in.SearchEngine(w["meta"], w["Vn"], w["bm25_corpus"], w["bm25_skus"], encode=lambda q: w["table"][q],
                                   reviews=w["reviews"])


def _check_snips(got, want):
    assert set(got) == set(want)
    for sku, v in want.items():
        g = got[sku]
        assert abs(g["score"] - v["score"]) <= SIM_ATOL, sku
        if g["text"] != v["text"]:
            # a different review of the same product may win only on a similarity tie within tolerance
            assert abs(g["score"] - v["score"]) <= SIM_ATOL
        else:
            assert g["stars"] == v["stars"]


def test_direct_calls_match_reference(w, eng):
    q = w["z"]["queries"]
    for c in w["snip"]["direct"]:
        qv = q[c["query_index"]]
        if c["driver"] == "cli":
            if c["max_rows"] <= 0:
                continue
            got = eng.best_review_snippets(qv, c["cand_skus"], max_rows=c["max_rows"])
        else:
            got = eng._best_snippets(qv, c["cand_skus"], max_rows=c["max_rows"])
        _check_snips(got, c["snippets"])


def test_review_index_batch_matches_oracle(w, eng):
    """A batch of queries x arbitrary candidate rows (with invalid rows and products without reviews)."""
    rng = np.random.default_rng(3)
    q = w["z"]["queries"]
    n = len(w["skus"])
    cand = rng.integers(0, n, size=(q.shape[0], 40)).astype(np.int64)
    cand[:, 5] = -1
    cand[:, 6] = 0          # product 0 has no reviews (make_reviews)
    cand[:, 7] = 7          # product 7 has hundreds
    for cap in (None, 90):
        score, pos = eng.review_ix.best(q, cand, max_rows=cap)
        for b in range(q.shape[0]):
            skus = [w["skus"][r] for r in cand[b] if r >= 0]
            want = best_review_snippets(q[b], skus, w["reviews"], max_rows=cap if cap is not None else 10**9)
            for j, r in enumerate(cand[b]):
                ref = want.get(w["skus"][r]) if r >= 0 else None
                if ref is None:
                    assert score[b, j] == 0.0 and pos[b, j] == -1
                else:
                    assert abs(score[b, j] - ref["score"]) <= SIM_ATOL
                    assert pos[b, j] == ref["file_pos"] or abs(score[b, j] - ref["score"]) <= SIM_ATOL


def test_drivers_with_snippets_reproduce_reference(w, eng):
    n_cases = 0
    for c in w["snip"]["cases"]:
        ps = c["params"]
        if c["driver"] == "streamlit":
            top, snips, dbg = eng.run_search(c["query"], ps["k"], ps["rerank_k"], ps["w_dense"], ps["w_bm25"],
                                             ps["w_rerank"], ps["w_prior"], ps["w_best"], ps["prior_C"], True,
                                             c["max_rows"], ps["min_reviews"], 1.0)
            _check_snips(snips, c["snippets"])
            best_col = "_best"
        else:
            args = types.SimpleNamespace(query=c["query"], k=ps["k"], rerank_k=ps["rerank_k"], w_dense=ps["w_dense"],
                                         w_bm25=ps["w_bm25"], w_rerank=ps["w_rerank"], w_prior=ps["w_prior"],
                                         w_best=ps["w_best"], prior_C=ps["prior_C"], gate_penalty=1.0,
                                         no_snippets=False, max_reviews_scan=c["max_rows"])
            top = eng.search(args)
            best_col = "_bestrev"
            for sku, text in zip(c["top_skus"], c["top_snippets"]):
                got = eng.last_snippets.get(sku)
                assert (got["text"] if got else None) == text
        order = np.argsort(-np.float32(c["pool_final"]), kind="stable")[:len(c["top_skus"])]
        ref_final = np.float32(c["pool_final"])[order]
        ref_best = dict(zip(c["pool_skus"], c["pool_best"]))
        np.testing.assert_allclose(top["_final"].values, ref_final, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(top[best_col].values, [ref_best[s] for s in top["sku"]], rtol=1e-5, atol=2e-6)
        got_rows = [int(s[3:]) for s in top["sku"].tolist()]
        ref_rows = [int(s[3:]) for s in c["top_skus"]]
        assert_ids_match_modulo_ties(got_rows, top["_final"].values, ref_rows, ref_final, 2e-6,
                                     f"{c['driver']} q{c['query_index']} cap{c['max_rows']}")
        n_cases += 1
    assert n_cases == 16
