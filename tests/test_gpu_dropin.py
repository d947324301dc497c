"""The reference-signature drop-ins (review-recommender_b200/drop_in.py): same calls the reference
makes, NumPy / DataFrame in and out, results against the reference's own tests and golden cases."""
import json
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

from oracle import primitives as P
from oracle.bm25_okapi import BM25Okapi as OracleBM25
from tests.parity import BM25_RTOL, assert_ids_match_modulo_ties


def _rr():
    import review_recommender_b200 as rr
    return rr


def test_cosine_similarity_search_reference_tests():
    """tests/test_utils.py:178-208 of the reference, against the drop-in."""
    d = _rr().drop_in
    emb = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], dtype=np.float32)
    q = np.array([0.0, 1.0], dtype=np.float32)
    idx, sims = d.cosine_similarity_search(q, emb, top_k=2)
    assert len(idx) == 2 and len(sims) == 2 and idx[0] == 1 and sims[0] == 1.0
    emb2 = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.float32)
    idx, sims = d.cosine_similarity_search(q, emb2, top_k=10)
    assert len(idx) == 2 and len(sims) == 2
    for fn in (d.cosine_search, d._cosine_pool):
        i2, s2 = fn(q, emb, 2)
        assert list(i2) == [1, 2]


def test_bm25okapi_class_matches_oracle_on_strings():
    rr = _rr()
    offs, toks = rr.synth.corpus_tokens(2500, 300)
    corpus = rr.synth.corpus_as_lists(offs, toks)
    want = OracleBM25(corpus)
    got = rr.drop_in.BM25Okapi(corpus, tile_docs=512)
    assert got.corpus_size == want.corpus_size and got.avgdl == want.avgdl and got.average_idf == want.average_idf
    assert got.idf == want.idf
    for q in (["t3", "t17"], ["t1", "t1", "zzz"], [], ["t250", "t9", "t40", "t2"]):
        s = got.get_scores(q)
        assert s.dtype == np.float64 and s.shape == (2500,)
        np.testing.assert_allclose(s, want.get_scores(q), rtol=BM25_RTOL)
    # the shim module is what `from rank_bm25 import BM25Okapi` resolves to
    shim_dir = str(Path(rr.drop_in.__file__).parent / "shims")
    sys.path.insert(0, shim_dir)
    try:
        import rank_bm25
        assert rank_bm25.BM25Okapi is rr.drop_in.BM25Okapi
    finally:
        sys.path.remove(shim_dir)
        sys.modules.pop("rank_bm25", None)


def test_search_engine_reproduces_reference_golden_cases(golden_dir):
    rr = _rr()
    cases = json.loads((golden_dir / "search_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    n = z["emb"].shape[0]
    skus = rr.synth.skus(n)
    corpus = rr.synth.corpus_as_lists(z["doc_offsets"], z["token_ids"])
    perm = z["bm25_perm"]
    meta = pd.DataFrame({"sku": skus, "n_reviews": z["n_reviews"], "avg_stars": z["avg_stars"],
                         "agg_text": ["" for _ in range(n)]})
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    table = {s: z["queries"][i] for i, s in enumerate(cases["query_strs"])}
    eng = rr.drop_in.SearchEngine(meta, Vn, [corpus[i] for i in perm], [skus[i] for i in perm],
                                  encode=lambda q: table[q])
    for c in cases["cases"]:
        ps = c["params"]
        if c["driver"] == "streamlit":
            top, snips, dbg = eng.run_search(c["query"], ps["k"], ps["rerank_k"], ps["w_dense"], ps["w_bm25"],
                                             ps["w_rerank"], ps["w_prior"], ps["w_best"], ps["prior_C"], False, 0,
                                             ps["min_reviews"], 1.0)
            assert dbg["pool"] == c["pool_size"] and dbg["tokens"] == c["tokens"]
        else:
            args = types.SimpleNamespace(query=c["query"], k=ps["k"], rerank_k=ps["rerank_k"], w_dense=ps["w_dense"],
                                         w_bm25=ps["w_bm25"], w_rerank=ps["w_rerank"], w_prior=ps["w_prior"],
                                         w_best=ps["w_best"], prior_C=ps["prior_C"], gate_penalty=1.0)
            top = eng.search(args)
        ref_final = np.sort(np.float32(c["pool_final"]))[::-1][:len(c["top_skus"])]
        got_rows = [int(s[3:]) for s in top["sku"].tolist()]
        ref_rows = [int(s[3:]) for s in c["top_skus"]]
        np.testing.assert_allclose(top["_final"].values, ref_final, rtol=1e-5, atol=1e-7)
        assert_ids_match_modulo_ties(got_rows, top["_final"].values, ref_rows, ref_final, 2e-6,
                                     f"{c['driver']} q{c['query_index']}")


def test_batched_search_equals_single_queries_and_artifact_loader(golden_dir, tmp_path):
    """run_search_batch (one GPU batch) == run_search per query, with gates and snippets on; the engine is built
    from artifact files in the reference's formats (product_emb.npy / product_emb_meta.parquet / product_bm25.pkl /
    reviews_with_embeddings.parquet); the reference's evaluate_ranking_methods loop shape is served from the batch."""
    import pickle
    from tests import snippet_world
    from tests.golden_worlds import GATE_QUERIES, make_gate_texts
    rr = _rr()
    w = snippet_world.load(golden_dir)
    z = w["z"]
    meta = w["meta"].copy()
    meta["agg_text"] = make_gate_texts(len(meta), z["doc_offsets"], z["token_ids"])
    np.save(tmp_path / "product_emb.npy", np.array(z["emb"]))                  # un-normalised rows on purpose
    meta.to_parquet(tmp_path / "product_emb_meta.parquet")
    with open(tmp_path / "product_bm25.pkl", "wb") as f:
        pickle.dump({"skus": w["bm25_skus"], "corpus": w["bm25_corpus"], "tokenizer": "simple_en_v1"}, f, protocol=4)
    w["reviews"].to_parquet(tmp_path / "reviews_with_embeddings.parquet")
    queries = list(w["cases"]["query_strs"]) + GATE_QUERIES[:5]
    table = dict(w["table"])
    table.update({q: z["queries"][i % len(z["queries"])] for i, q in enumerate(GATE_QUERIES)})
    eng = rr.drop_in.SearchEngine.from_artifacts(tmp_path / "product_emb.npy", tmp_path / "product_emb_meta.parquet",
                                                 tmp_path / "product_bm25.pkl", tmp_path / "reviews_with_embeddings.parquet",
                                                 encode=lambda q: table[q])
    assert eng.gate_ix is not None and eng.review_ix is not None and eng.bm25_active
    cfg = dict(k=20, rerank_k=0, w_dense=0.45, w_bm25=0.15, w_rerank=0.0, w_prior=0.10, w_best=0.30, prior_C=20.0,
               use_snips=True, max_scan=300_000, min_reviews=8, gate_penalty=0.5)
    batch = eng.run_search_batch(queries, **cfg)
    assert len(batch) == len(queries)
    for q, (top_b, snips_b, dbg_b) in zip(queries, batch):
        top_1, snips_1, dbg_1 = eng.run_search(q, *[cfg[k] for k in ("k", "rerank_k", "w_dense", "w_bm25", "w_rerank",
                                                                      "w_prior", "w_best", "prior_C", "use_snips",
                                                                      "max_scan", "min_reviews", "gate_penalty")])
        assert top_b["sku"].tolist() == top_1["sku"].tolist(), q
        for col in ("_final", "_dense", "_bm25", "_prior", "_best", "_gate", "_trust"):
            np.testing.assert_array_equal(top_b[col].values, top_1[col].values, err_msg=f"{q} {col}")
        assert snips_b == snips_1 and dbg_b == dbg_1
    fn = eng.batched_search_function(queries)
    for q in queries[:3]:
        res, _, _ = fn(q, **cfg)
        assert res["sku"].tolist() == batch[queries.index(q)][0]["sku"].tolist()
    with pytest.raises(SystemExit):
        rr.drop_in.load_product_index(tmp_path / "missing.npy", tmp_path / "product_emb_meta.parquet")
