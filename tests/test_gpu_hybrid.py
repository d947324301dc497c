"""End-to-end parity of the one-shot hybrid search (host buffers through rr_hybrid_search_host)
against the oracle drivers, and against the golden cases captured from the reference itself."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import primitives as P
from tests.parity import check_hybrid_against_oracle


@pytest.mark.parametrize("n,d,v,b,l,k,driver", [
    (10_000, 384, 20_000, 4, 4, 10, "streamlit"),      # BASELINE configs[0]: reference CPU case
    (10_000, 384, 20_000, 4, 4, 10, "cli"),
    (3_000, 64, 500, 16, 4, 100, "streamlit"),
    (120, 32, 60, 5, 3, 10, "streamlit"),              # fewer docs than the pool floor
    (40_000, 128, 5_000, 8, 16, 100, "streamlit"),
])
def test_hybrid_matches_oracle(n, d, v, b, l, k, driver):
    frac = check_hybrid_against_oracle(n, d, v, b, l, k, driver=driver, rerank_k=0)
    print(f"bit-exact id fraction {frac:.4f}")
    assert frac >= 0.9


def test_config5_shape_dense_heavy_top1000():
    """BASELINE configs[4] at reduced N: 768-d, fusion weight 0.8 on dense, top-1000 for the reranker
    (pool = 1000), through the tensor path (batch >= 32, N >= 65536) and through the exact path."""
    import review_recommender_b200 as rr
    for mode in (rr._lib.RR_DENSE_TENSOR, rr._lib.RR_DENSE_EXACT):
        frac = check_hybrid_against_oracle(70_000, 768, 20_000, 4, 4, 1000, driver="streamlit", mode=mode, rerank_k=0,
                                           w_dense=0.8, w_bm25=0.1, w_prior=0.1, w_rerank=0.0, w_best=0.0)
        assert frac >= 0.9


def test_golden_cases_from_the_reference(golden_dir):
    """The reference's own run_search / search outputs (tests/golden/search_cases.json) reproduced
    through the C-ABI: same top-k SKUs, fused scores within 1e-5 relative."""
    import review_recommender_b200 as rr
    from tests.parity import assert_ids_match_modulo_ties
    cases = json.loads((golden_dir / "search_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    n = z["emb"].shape[0]
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    ix = rr.engine.HybridIndex(Vn, z["doc_offsets"], z["token_ids"], cases["V"], z["n_reviews"], z["avg_stars"],
                               device="cuda:0", tile_docs=256)
    for c in cases["cases"]:
        ps = dict(c["params"])
        toks = P.tokenize_query(c["query"])
        ids = [int(t[1:]) - 1 if (t.startswith("t") and t[1:].isdigit()) else -1 for t in toks]
        tid, nt = rr.engine.HybridIndex.pack_terms([ids])
        fusion = rr.engine.Fusion(driver=c["driver"], **ps)
        rows, final = ix.hybrid_search_host(z["queries"][c["query_index"]][None], tid, nt, fusion,
                                            mode=rr._lib.RR_DENSE_EXACT)
        ref_rows = [int(s[3:]) for s in c["top_skus"]]
        ref_final = np.sort(np.float32(c["pool_final"]))[::-1][:len(ref_rows)]
        np.testing.assert_allclose(final[0, :len(ref_rows)], ref_final, rtol=1e-5, atol=1e-7)
        assert_ids_match_modulo_ties(rows[0, :len(ref_rows)], final[0, :len(ref_rows)], ref_rows, ref_final, 2e-6,
                                     f"{c['driver']} q{c['query_index']}")
    ix.close()
