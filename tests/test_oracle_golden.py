"""Oracle restatements against outputs of the reference's own code captured by
tests/golden/make_golden.py (utils.py, app/test.py `search`, app/app_product_search.py
`run_search`).  CPU only; does not need /root/reference."""
import json

import numpy as np
import pandas as pd
import pytest

from oracle import primitives as P
from oracle.bm25_okapi import BM25Okapi
from oracle.hybrid import cli_search_core, run_search_core
import review_recommender_b200 as rr


@pytest.fixture(scope="module")
def prim(golden_dir):
    return np.load(golden_dir / "primitives.npz")


@pytest.fixture(scope="module")
def cases(golden_dir):
    return json.loads((golden_dir / "search_cases.json").read_text())


def test_l2_normalize(prim):
    np.testing.assert_array_equal(P.l2_normalize(prim["l2_in"]), prim["l2_out"])


@pytest.mark.parametrize("name", ["f32", "f64", "const", "nan", "inf", "tiny", "one", "bm25like"])
def test_minmax(prim, name):
    got = P.minmax_normalize(prim[f"mm_in_{name}"])
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, prim[f"mm_out_{name}"])


def test_prior_trust(prim):
    np.testing.assert_array_equal(P.bayesian_prior(prim["prior_avg"], prim["prior_n"], 20.0), prim["prior_out"])
    np.testing.assert_array_equal(P.bayesian_prior(np.array([4.0, 3.0, np.nan]), np.array([0, 5, 10])),
                                  prim["prior_small"])
    np.testing.assert_allclose(prim["prior_small"][:2], [3.5, 3.4], rtol=1e-9)
    n = prim["trust_n"]
    np.testing.assert_array_equal(P.trust_score_from_reviews(n, 8, 50), prim["trust_out_50"])
    np.testing.assert_array_equal(P.trust_score_from_reviews(n, 8, 80), prim["trust_out_80"])
    np.testing.assert_array_equal(P.trust_score_from_reviews(n.astype(np.float64), 0, 80), prim["trust_out_m0"])


def test_cosine_search(prim):
    idx, sims = P.cosine_similarity_search(prim["cos_q"], prim["cos_mat"], 25)
    np.testing.assert_array_equal(idx, prim["cos_idx"])
    np.testing.assert_array_equal(sims, prim["cos_sims"])
    idx, sims = P.cosine_similarity_search(prim["cos_q"], prim["cos_mat"][:10], 50)
    np.testing.assert_array_equal(idx, prim["cos_idx_clamp"])
    idx2, sims2 = P.cosine_topk_canonical(prim["cos_q"], prim["cos_mat"], 25)
    np.testing.assert_array_equal(idx2, prim["cos_idx"])          # no ties in this draw


def test_tokenizer(cases):
    for q, want in zip(cases["tokenize"]["queries"], cases["tokenize"]["tokens"]):
        assert P.tokenize_query(q) == want


@pytest.fixture(scope="module")
def world(golden_dir, cases):
    z = np.load(golden_dir / "search_cases.npz")
    n = z["emb"].shape[0]
    skus = rr.synth.skus(n)
    corpus = rr.synth.corpus_as_lists(z["doc_offsets"], z["token_ids"])
    perm = z["bm25_perm"]
    meta = pd.DataFrame({"sku": skus, "n_reviews": z["n_reviews"], "avg_stars": z["avg_stars"],
                         "agg_text": ["" for _ in range(n)]})
    bm25 = BM25Okapi([corpus[i] for i in perm])
    bm25_skus = [skus[i] for i in perm]
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    return dict(z=z, meta=meta, bm25=bm25, bm25_skus=bm25_skus, Vn=Vn)


def test_search_drivers_match_reference(cases, world):
    z = world["z"]
    n_cli = n_st = 0
    for c in cases["cases"]:
        ps = dict(c["params"])
        q = z["queries"][c["query_index"]]
        toks = P.tokenize_query(c["query"])
        if c["driver"] == "cli":
            top, pool = cli_search_core(q, world["Vn"], world["meta"], world["bm25"], world["bm25_skus"], toks, **ps)
            n_cli += 1
        else:
            top, pool = run_search_core(q, world["Vn"], world["meta"], world["bm25"], world["bm25_skus"], toks, **ps)
            assert len(pool) == c["pool_size"] and toks == c["tokens"]
            np.testing.assert_array_equal(pool["_trust"].values, np.float32(c["pool_trust"]))
            n_st += 1
        assert pool["sku"].tolist() == c["pool_skus"]
        np.testing.assert_array_equal(pool["_final"].values.astype(np.float32), np.float32(c["pool_final"]))
        np.testing.assert_array_equal(pool["_dense"].values.astype(np.float32), np.float32(c["pool_dense"]))
        np.testing.assert_array_equal(pool["_bm25"].values.astype(np.float64), np.float64(c["pool_bm25"]))
        np.testing.assert_array_equal(pool["_prior"].values, np.float64(c["pool_prior"]))
        assert top["sku"].tolist() == c["top_skus"]
    assert n_cli == 18 and n_st == 24


@pytest.mark.parametrize("D", [7, 100, 130, 384, 768])
def test_l2_normalize_long_rows(prim, D):
    np.testing.assert_array_equal(P.l2_normalize(prim[f"l2big_in_{D}"]), prim[f"l2big_out_{D}"])
