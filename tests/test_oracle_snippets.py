"""oracle.snippets against the reference's own best_review_snippets / _best_snippets and both drivers
run with snippets on (tests/golden/snippet_cases.json).  CPU only."""
import numpy as np
import pytest

from oracle import primitives as P
from oracle.bm25_okapi import BM25Okapi
from oracle.hybrid import cli_search_core, run_search_core
from oracle.snippets import best_review_snippets, streamlit_best_snippets
from tests import snippet_world


@pytest.fixture(scope="module")
def w(golden_dir):
    return snippet_world.load(golden_dir)


def _strip(d):
    return {k: {f: v[f] for f in ("score", "text", "stars")} for k, v in d.items()}


def test_direct_calls_match_reference(w):
    q = w["z"]["queries"]
    n = 0
    for c in w["snip"]["direct"]:
        qv = q[c["query_index"]]
        if c["driver"] == "cli":
            got = best_review_snippets(qv, c["cand_skus"], w["reviews"], max_rows=c["max_rows"])
        else:
            got = streamlit_best_snippets(qv, c["cand_skus"], w["reviews"], max_rows=c["max_rows"])
        assert _strip(got) == c["snippets"], (c["driver"], c["query_index"], c["max_rows"])
        n += 1
    assert n == 18
    assert any(c["max_rows"] < 1000 and 0 < len(c["snippets"]) < 100 for c in w["snip"]["direct"])   # the cap bites
    assert any(c["max_rows"] == 0 and c["snippets"] == {} for c in w["snip"]["direct"])


def test_drivers_with_snippets_match_reference(w):
    bm25 = BM25Okapi(w["bm25_corpus"])
    for c in w["snip"]["cases"]:
        qv = w["z"]["queries"][c["query_index"]]
        toks = P.tokenize_query(c["query"])
        ps = dict(c["params"])
        if c["driver"] == "cli":
            def best_fn(skus, c=c, qv=qv):
                d = best_review_snippets(qv, skus, w["reviews"], max_rows=c["max_rows"])
                return [d.get(s, {}).get("score") for s in skus]
            top, pool = cli_search_core(qv, w["Vn"], w["meta"], bm25, w["bm25_skus"], toks, best_scores_fn=best_fn, **ps)
            best_col = "_bestrev"
        else:
            def best_fn(skus, c=c, qv=qv):
                d = streamlit_best_snippets(qv, skus, w["reviews"], max_rows=c["max_rows"])
                return [d.get(s, {}).get("score") for s in skus]
            top, pool = run_search_core(qv, w["Vn"], w["meta"], bm25, w["bm25_skus"], toks, best_scores_fn=best_fn, **ps)
            best_col = "_best"
        assert pool["sku"].tolist() == c["pool_skus"]
        np.testing.assert_array_equal(pool[best_col].values.astype(np.float32), np.float32(c["pool_best"]))
        np.testing.assert_array_equal(pool["_final"].values.astype(np.float32), np.float32(c["pool_final"]))
        assert top["sku"].tolist() == c["top_skus"]
