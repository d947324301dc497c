"""Exact dense path (fp32 GEMV + radix select) against NumPy `mat @ q` and the reference's
known-answer tests for cosine_similarity_search (tests/test_utils.py:178-208)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.primitives import cosine_topk_canonical
from tests.parity import DENSE_ATOL, assert_ids_match_modulo_ties


def _rr():
    import review_recommender_b200 as rr
    return rr


def test_reference_known_answers():
    rr = _rr()
    emb = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], dtype=np.float32)
    ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
    idx, sims, cnt = ix.dense_topk(np.array([0.0, 1.0], dtype=np.float32), 2, rr._lib.RR_DENSE_EXACT)
    idx, sims = idx.cpu().numpy()[0], sims.cpu().numpy()[0]
    # rows 1 and 2 both score exactly 1.0; the documented tie policy returns the lower row first
    assert list(idx) == [1, 2] and sims[0] == 1.0 and sims[1] == 1.0 and int(cnt[0]) == 2
    ix.close()
    emb = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=np.float32)
    ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
    idx, sims, cnt = ix.dense_topk(np.array([0.0, 1.0], dtype=np.float32), 10, rr._lib.RR_DENSE_EXACT)
    assert int(cnt[0]) == 2 and list(idx.cpu().numpy()[0][:2]) == [1, 0] and np.all(idx.cpu().numpy()[0][2:] == -1)
    ix.close()


@pytest.mark.parametrize("n,d,b,k", [(1, 8, 1, 1), (257, 2, 3, 5), (5000, 384, 1, 150), (5000, 384, 19, 150),
                                      (30011, 100, 9, 1000), (4096, 37, 4, 4096), (70000, 64, 2, 100)])
def test_exact_topk_matches_numpy(n, d, b, k):
    rr = _rr()
    emb = rr.synth.embeddings(n, d)
    q = rr.synth.queries(b, d)
    ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
    idx, sims, cnt = ix.dense_topk(q, k, rr._lib.RR_DENSE_EXACT)
    idx, sims, cnt = idx.cpu().numpy(), sims.cpu().numpy(), cnt.cpu().numpy()
    kk = min(k, n)
    for i in range(b):
        ref_idx, ref_sims = cosine_topk_canonical(q[i], emb, k)
        assert cnt[i] == kk
        np.testing.assert_allclose(sims[i, :kk], ref_sims, rtol=0, atol=DENSE_ATOL)
        assert np.all(np.diff(sims[i, :kk]) <= 0)
        assert_ids_match_modulo_ties(idx[i, :kk], sims[i, :kk], ref_idx, ref_sims, 2 * DENSE_ATOL, f"q{i}")
        assert len(set(idx[i, :kk].tolist())) == kk
    ix.close()


def test_ties_and_duplicates_are_ordered_by_row():
    rr = _rr()
    emb = rr.synth.embeddings(3000, 32)
    emb[100:140] = emb[7]                       # 41 identical rows
    q = emb[7:8].copy()
    ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
    idx, sims, _ = ix.dense_topk(q, 20, rr._lib.RR_DENSE_EXACT)
    idx = idx.cpu().numpy()[0]
    assert list(idx) == [7] + list(range(100, 119))
    # batch-size independence: the same query inside a batch of 9 gives bit-identical output
    qb = np.concatenate([rr.synth.queries(8, 32), q])
    idx9, sims9, _ = ix.dense_topk(qb, 20, rr._lib.RR_DENSE_EXACT)
    np.testing.assert_array_equal(idx9.cpu().numpy()[8], idx)
    np.testing.assert_array_equal(sims9.cpu().numpy()[8], sims.cpu().numpy()[0])
    ix.close()


@pytest.mark.parametrize("n,d,b,k", [(16_385, 16, 1, 150), (200_000, 16, 1, 150), (200_000, 16, 8, 1024), (150_000, 8, 3, 1), (200_000, 16, 20, 150),
                                      (2_500_000, 8, 2, 150),         # three levels: 611 chunks -> 12 -> 1
                                      (1_200_000, 8, 1, 1000)])       # four levels: 293 chunks x 1000 survivors -> 36 -> 5 -> 1
def test_topk_tree_equals_the_radix_pipeline(n, d, b, k, monkeypatch):
    """Small batches over long rows select through the shared-memory top-k tree (topk_chunk_kernel); RR_NO_CHUNKED_TOPK=1
    sends the same scores through the radix-select pipeline.  Same rows, same similarities, same order -- also with rows
    that tie across chunk borders."""
    rr = _rr()
    emb = rr.synth.embeddings(n, d)
    q = rr.synth.queries(b, d)
    for r in (3, 16_384, n // 2, n - 1):                 # copies of the best row of query 0, spread over the chunks
        emb[r] = q[0] / np.linalg.norm(q[0])
    ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
    out = {}
    for tree in (True, False):
        if tree:
            monkeypatch.delenv("RR_NO_CHUNKED_TOPK", raising=False)
        else:
            monkeypatch.setenv("RR_NO_CHUNKED_TOPK", "1")
        rr.engine.launch_count(reset=True)
        idx, sims, cnt = ix.dense_topk(q, k, rr._lib.RR_DENSE_EXACT)
        out[tree] = (idx.cpu().numpy(), sims.cpu().numpy(), cnt.cpu().numpy(), rr.engine.launch_count())
    np.testing.assert_array_equal(out[True][0], out[False][0])
    np.testing.assert_array_equal(out[True][1], out[False][1])
    np.testing.assert_array_equal(out[True][2], out[False][2])
    assert out[True][3] + 8 <= out[False][3], "the tree must replace the 15-launch pipeline by 2-4 launches"
    twins = sorted({3, 16_384, n // 2, n - 1})
    assert list(out[True][0][0][:min(k, len(twins))]) == twins[:min(k, len(twins))]
    ref_idx, ref_sims = cosine_topk_canonical(q[b - 1], emb, k)
    np.testing.assert_allclose(out[True][1][b - 1], ref_sims, rtol=0, atol=DENSE_ATOL)
    assert_ids_match_modulo_ties(out[True][0][b - 1], out[True][1][b - 1], ref_idx, ref_sims, 2 * DENSE_ATOL)
    ix.close()
