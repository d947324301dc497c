"""Small and ragged shapes through every kernel of the single-query / small-batch path (fp32 GEMV per batch width,
one-kernel and tree top-k, warp-per-candidate BM25 gather, fusion with rank sorts): each result must equal, bit for bit,
what the radix-select pipeline and the thread-per-candidate gather return for the same call."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OLD_PATH = {"RR_NO_SMALL_TOPK": "1", "RR_NO_CHUNKED_TOPK": "1", "RR_BM25_CAND_WARP": "0"}


def _set(monkeypatch, old):
    for name, val in OLD_PATH.items():
        if old:
            monkeypatch.setenv(name, val)
        else:
            monkeypatch.delenv(name, raising=False)


@pytest.mark.parametrize("n", [3000, 20_000, 40_001])     # one-kernel top-k, two-level tree, ragged last chunk
def test_small_batch_path_equals_the_general_path(n, monkeypatch):
    import torch
    import review_recommender_b200 as rr
    monkeypatch.setenv("RR_NO_GRAPHS", "1")
    D, V, L = 32, 2000, 4
    c = rr.synth.make_corpus(n, D, V)
    ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), V,
                               c.n_reviews, c.avg_stars, make_bf16=False)
    for b in (1, 3, 8, 20):
        q = rr.synth.queries(b, D)
        qt = rr.synth.query_terms(b, L, c.doc_offsets, c.token_ids, V).astype(np.int32)
        nt = np.full(b, L, dtype=np.int32)
        for k, rerank_k in ((10, 0), (100, 1000)):            # pools of 150 and 1000 candidates
            fusion = rr.engine.Fusion(k=k, rerank_k=rerank_k, w_rerank=0.0, w_best=0.0)
            got = []
            for old in (False, True):
                _set(monkeypatch, old)
                rows, fin = ix.hybrid_search_host(q, qt, nt, fusion, mode=rr._lib.RR_DENSE_EXACT)
                assert rows.shape == (b, k) and np.all(np.diff(fin.astype(np.float64), axis=1) <= 0)
                got.append((rows.copy(), fin.copy()))
            np.testing.assert_array_equal(got[0][0], got[1][0])
            np.testing.assert_array_equal(got[0][1], got[1][1])
        for k in (1, 150, 1024):
            got = []
            for old in (False, True):
                _set(monkeypatch, old)
                idx, sims, cnt = ix.dense_topk(q, k, rr._lib.RR_DENSE_EXACT)
                assert int(cnt.min()) == min(k, n)
                got.append((idx.cpu().numpy(), sims.cpu().numpy()))
            np.testing.assert_array_equal(got[0][0], got[1][0])
            np.testing.assert_array_equal(got[0][1], got[1][1])
    ix.close()
