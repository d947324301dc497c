"""SURVEY 8a row a5 (BM25 gather rules): the oracle's copies of _bm25_for_candidates, ensure_same_order and
bm25_scores against outputs of the REFERENCE's own functions (tests/golden/a5_cases.json, written by
tests/golden/make_golden.py with oracle.BM25Okapi as the bm25 object).  CPU only."""
import json

import numpy as np
import pandas as pd
import pytest

import review_recommender_b200 as rr
from oracle.bm25_okapi import BM25Okapi
from oracle.hybrid import bm25_for_candidates, bm25_scores, cli_search_core, ensure_same_order
from oracle.primitives import tokenize_query


@pytest.fixture(scope="module")
def a5(golden_dir):
    g = json.loads((golden_dir / "a5_cases.json").read_text())
    z = np.load(golden_dir / "a5_cases.npz")
    c = rr.synth.make_corpus(g["N"], g["D"], g["V"])
    corpus = rr.synth.corpus_as_lists(c.doc_offsets, c.token_ids)
    corpus_a = [corpus[i] for i in z["perm"]]
    return dict(g=g, z=z, c=c, corpus_a=corpus_a, bm25=BM25Okapi(corpus_a), skus=rr.synth.skus(g["N"]))


def test_streamlit_gather_duplicate_absent_permuted(a5):
    g = a5["g"]
    for case in g["direct"]:
        got = bm25_for_candidates(a5["bm25"], g["skus_a"], tokenize_query(case["query"]), case["cand_skus"])
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, np.asarray(case["bm25"], dtype=np.float32))
    np.testing.assert_array_equal(bm25_for_candidates(None, None, ["x"], g["direct"][0]["cand_skus"]),
                                  np.asarray(g["none_blob"], dtype=np.float32))


def test_cli_permutation_and_identity_rule(a5):
    g = a5["g"]
    skus_b = [a5["skus"][i] for i in a5["z"]["perm"]]
    assert ensure_same_order(a5["skus"], g["skus_a"]) is None          # two meta SKUs have no document
    order_b = ensure_same_order(a5["skus"], skus_b)
    top_idx = np.asarray(g["top_idx"])
    for case in g["cli_direct"]:
        toks = tokenize_query(case["query"])
        np.testing.assert_array_equal(bm25_scores(a5["bm25"], toks, None, top_idx), np.asarray(case["identity"], np.float32))
        np.testing.assert_array_equal(bm25_scores(a5["bm25"], toks, order_b, top_idx), np.asarray(case["permuted"], np.float32))


def test_cli_driver_on_both_blobs(a5):
    g, c = a5["g"], a5["c"]
    meta = pd.DataFrame({"sku": a5["skus"], "n_reviews": c.n_reviews.astype(np.float64), "avg_stars": c.avg_stars,
                         "agg_text": ["" for _ in a5["skus"]]})
    for case in g["cases"]:
        blob_skus = g["skus_a"] if case["blob"] == "identity" else [a5["skus"][i] for i in a5["z"]["perm"]]
        top, pool = cli_search_core(a5["z"]["queries"][case["query_index"]], c.emb, meta, a5["bm25"], blob_skus,
                                    tokenize_query(case["query"]), k=10, rerank_k=0, w_dense=0.4, w_bm25=0.4,
                                    w_rerank=0.0, w_prior=0.2, w_best=0.0, prior_C=20.0)
        assert pool["sku"].tolist() == case["pool_skus"]
        np.testing.assert_array_equal(pool["_bm25"].values.astype(np.float32), np.asarray(case["pool_bm25"], np.float32))
        np.testing.assert_allclose(pool["_final"].values, case["pool_final"], rtol=1e-6, atol=1e-7)
        assert top["sku"].tolist() == case["top_skus"]
