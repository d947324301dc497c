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
