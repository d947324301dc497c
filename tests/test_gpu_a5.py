"""SURVEY 8a row a5 on the CUDA path: drop_in._bm25_for_candidates / bm25_scores / ensure_same_order and
SearchEngine.search (CLI driver, permutation AND identity-order rule) against outputs of the reference's own
functions (tests/golden/a5_cases.json); the dense cache of the pure cosine drop-ins never serves a stale matrix."""
import json
import types

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

from tests.parity import BM25_RTOL, FUSED_RTOL


@pytest.fixture(scope="module")
def a5(golden_dir):
    import review_recommender_b200 as rr
    g = json.loads((golden_dir / "a5_cases.json").read_text())
    z = np.load(golden_dir / "a5_cases.npz")
    c = rr.synth.make_corpus(g["N"], g["D"], g["V"])
    corpus = rr.synth.corpus_as_lists(c.doc_offsets, c.token_ids)
    corpus_a = [corpus[i] for i in z["perm"]]
    return dict(rr=rr, g=g, z=z, c=c, corpus_a=corpus_a, bm25=rr.drop_in.BM25Okapi(corpus_a), skus=rr.synth.skus(g["N"]))


def test_bm25_for_candidates_duplicate_absent_permuted(a5):
    rr, g = a5["rr"], a5["g"]
    blob = {"bm25": a5["bm25"], "skus": g["skus_a"]}
    for case in g["direct"]:
        got = rr.drop_in._bm25_for_candidates(blob, case["query"], case["cand_skus"])
        assert got.dtype == np.float32 and got.shape == (len(case["cand_skus"]),)
        want = np.asarray(case["bm25"], dtype=np.float32)
        np.testing.assert_allclose(got, want, rtol=BM25_RTOL, atol=0)
        assert np.array_equal(got == 0, want == 0)
    np.testing.assert_array_equal(rr.drop_in._bm25_for_candidates(None, "t1", g["direct"][0]["cand_skus"]),
                                  np.asarray(g["none_blob"], dtype=np.float32))


def test_bm25_scores_permutation_and_identity(a5):
    rr, g = a5["rr"], a5["g"]
    meta = pd.DataFrame({"sku": a5["skus"]})
    skus_b = [a5["skus"][i] for i in a5["z"]["perm"]]
    assert rr.drop_in.ensure_same_order(meta, g["skus_a"]) is None
    order_b = rr.drop_in.ensure_same_order(meta, skus_b)
    assert order_b == [skus_b.index(s) for s in a5["skus"]]
    top_idx = np.asarray(g["top_idx"])
    for case in g["cli_direct"]:
        toks = rr.drop_in.tokenize_query(case["query"])
        for order, key in ((None, "identity"), (order_b, "permuted")):
            got = rr.drop_in.bm25_scores(a5["bm25"], toks, order, top_idx)
            want = np.asarray(case[key], dtype=np.float32)
            np.testing.assert_allclose(got, want, rtol=BM25_RTOL, atol=0)
            assert np.array_equal(got == 0, want == 0)


def test_search_engine_cli_driver_on_both_blobs(a5):
    rr, g, c = a5["rr"], a5["g"], a5["c"]
    meta = pd.DataFrame({"sku": a5["skus"], "n_reviews": c.n_reviews.astype(np.float64), "avg_stars": c.avg_stars})
    table = {s: a5["z"]["queries"][i] for i, s in enumerate(g["query_strs"])}
    for blob_name in ("identity", "permuted"):
        blob_skus = g["skus_a"] if blob_name == "identity" else [a5["skus"][i] for i in a5["z"]["perm"]]
        se = rr.drop_in.SearchEngine(meta, c.emb, a5["corpus_a"], blob_skus, encode=lambda q: table[q])
        assert se._cli_identity == (blob_name == "identity")
        for case in [x for x in g["cases"] if x["blob"] == blob_name]:
            args = types.SimpleNamespace(query=case["query"], k=10, rerank_k=0, w_dense=0.4, w_bm25=0.4, w_rerank=0.0,
                                         w_prior=0.2, w_best=0.0, prior_C=20.0, gate_penalty=1.0, no_snippets=True)
            out = se.search(args)
            assert out["sku"].tolist() == case["top_skus"]
            ref = dict(zip(case["pool_skus"], case["pool_final"]))
            np.testing.assert_allclose(out["_final"].values, [ref[s] for s in out["sku"]], rtol=FUSED_RTOL, atol=1e-7)
            refb = dict(zip(case["pool_skus"], case["pool_bm25"]))
            np.testing.assert_allclose(out["_bm25"].values, [refb[s] for s in out["sku"]], rtol=1e-5, atol=1e-7)


def test_dense_cache_never_serves_a_stale_matrix():
    import review_recommender_b200 as rr
    rng = np.random.default_rng(3)
    q = rng.standard_normal(64).astype(np.float32)
    for trial in range(6):                  # fresh same-shaped matrices: NumPy tends to reuse the freed address
        mat = rng.standard_normal((2000, 64)).astype(np.float32)
        idx, sims = rr.drop_in.cosine_similarity_search(q, mat, 5)
        want = np.argsort(-(mat @ q))[:5]
        np.testing.assert_array_equal(idx, want)
        del mat
    mat = rng.standard_normal((2000, 64)).astype(np.float32)
    idx1, _ = rr.drop_in.cosine_similarity_search(q, mat, 5)
    mat[:] = rng.standard_normal((2000, 64)).astype(np.float32)          # in-place edit of the cached object
    idx2, _ = rr.drop_in.cosine_similarity_search(q, mat, 5)
    np.testing.assert_array_equal(idx2, np.argsort(-(mat @ q))[:5])
    assert len(rr.drop_in._dense_cache) <= rr.drop_in._DENSE_CACHE_MAX
