"""Tensor-core shortlist path (tcgen05 bf16 GEMM + threshold filter + exact rescoring +
certification) against the exact fp32 path: results must be BIT-IDENTICAL (same rows, same
similarities), because every similarity returned is the canonical fp32 dot product and
uncertified queries are redone exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.primitives import cosine_topk_canonical
from tests.parity import DENSE_ATOL, assert_ids_match_modulo_ties


def _rr():
    import review_recommender_b200 as rr
    return rr


def _both(ix, q, pool):
    rr = _rr()
    i1, s1, c1 = ix.dense_topk(q, pool, rr._lib.RR_DENSE_EXACT)
    i2, s2, c2 = ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
    st = ix.dense_stats()
    return (i1.cpu().numpy(), s1.cpu().numpy(), c1.cpu().numpy()), (i2.cpu().numpy(), s2.cpu().numpy(), c2.cpu().numpy()), st


@pytest.mark.parametrize("n,d,b,pool", [(5000, 64, 40, 20), (70_000, 384, 200, 150), (33_333, 100, 130, 150),
                                         (300_000, 384, 256, 150), (1000, 128, 3, 150),
                                         (60_000, 768, 130, 150),      # 768-d: query k-blocks streamed (configs[4] shape)
                                         (120_000, 768, 64, 1000),     # top-1000 shortlist (configs[4]: k = pool = 1000)
                                         (50_000, 448, 33, 48)])       # small pool of a sharded round 1
def test_tensor_path_equals_exact_path(n, d, b, pool):
    rr = _rr()
    emb = rr.synth.embeddings(n, d)
    q = rr.synth.queries(b, d)
    ix = rr.engine.HybridIndex(emb, device="cuda:0")
    (i1, s1, c1), (i2, s2, c2), st = _both(ix, q, pool)
    print("tensor path stats:", st)
    assert st["path"] == 2
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(s1, s2)
    assert st["n_uncertified"] <= max(1, b // 10)
    # and both agree with NumPy on a few queries
    for i in range(min(b, 3)):
        ref_idx, ref_sims = cosine_topk_canonical(q[i], emb, pool)
        kk = len(ref_idx)
        np.testing.assert_allclose(s2[i, :kk], ref_sims, rtol=0, atol=DENSE_ATOL)
        assert_ids_match_modulo_ties(i2[i, :kk], s2[i, :kk], ref_idx, ref_sims, 2 * DENSE_ATOL)
    ix.close()


def test_near_duplicates_fall_back_to_exact_and_stay_exact():
    """Rows that differ by less than the bf16 error bound cannot be certified from bf16 scores;
    the query must be redone by the exact path and still return the exact answer."""
    rr = _rr()
    n, d = 40_000, 128
    emb = rr.synth.embeddings(n, d)
    rng = np.random.default_rng(3)
    base = emb[17].copy()
    for r in range(1000, 1600):                       # 600 rows within 1e-4 of row 17
        v = base + 1e-4 * rng.standard_normal(d).astype(np.float32)
        emb[r] = v / np.linalg.norm(v)
    q = np.stack([base, rr.synth.queries(1, d)[0]])
    ix = rr.engine.HybridIndex(emb, device="cuda:0")
    (i1, s1, c1), (i2, s2, c2), st = _both(ix, q, 150)
    print("tensor path stats:", st)
    assert st["n_uncertified"] >= 1
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(s1, s2)
    ix.close()


def test_bf16_scores_within_1e3_of_exact_before_rescoring():
    """north_star: "dense cosine scores must match within 1e-3 absolute before rescoring".  The raw tensor-core scores
    (rr_dense_debug_bf16_scores: tcgen05 bf16 x bf16 -> fp32, what the threshold filter sees) against the exact fp32
    similarities on configs[1]-shaped data (384-d unit rows, recipe of SURVEY 8d)."""
    import torch
    import review_recommender_b200 as rr
    n, d, b = 200_000, 384, 96
    emb = rr.synth.embeddings(n, d)
    q = rr.synth.queries(b, d)
    ix = rr.engine.HybridIndex(emb, device="cuda:0")
    worst = 0.0
    for row0, rows in ((0, 148 * 256), (150_016, 30_000), (n - 256 - (n - 256) % 256, 256 + (n - 256) % 256)):
        got = ix.debug_bf16_scores(q, row0, rows).cpu().numpy()
        assert not np.isnan(got).any(), "every row of the range must have been scored"
        want = (emb[row0:row0 + rows] @ q.T).T
        worst = max(worst, float(np.max(np.abs(got - want))))
    assert worst <= 1e-3, worst
    # and the certification margin the library actually uses is a proven bound, looser than the stated tolerance
    _, _, _ = ix.dense_topk(q, 150, rr._lib.RR_DENSE_TENSOR)
    assert ix.dense_stats()["eps"] >= worst
    ix.close()


def test_cta_pair_kernel_equals_one_cta_kernel(monkeypatch):
    """tc_filter_pair_kernel (tcgen05 cta_group::2, an ODD number of query tiles padded to an even one) and
    tc_filter_kernel return the same exact pools; both equal the fp32 path."""
    rr = _rr()
    n, d, b, pool = 150_000, 384, 1100, 150            # 9 query tiles -> 10 for the pair kernel
    emb = rr.synth.embeddings(n, d)
    q = rr.synth.queries(b, d)
    ix = rr.engine.HybridIndex(emb, device="cuda:0")
    out = {}
    for pair in ("1", "0"):
        monkeypatch.setenv("RR_TC_PAIR", pair)
        i2, s2, c2 = ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
        assert ix.dense_stats()["path"] == 2
        out[pair] = (i2.cpu().numpy(), s2.cpu().numpy())
    np.testing.assert_array_equal(out["1"][0], out["0"][0])
    np.testing.assert_array_equal(out["1"][1], out["0"][1])
    i1, s1, _ = ix.dense_topk(q[:64], pool, rr._lib.RR_DENSE_EXACT)
    np.testing.assert_array_equal(out["1"][0][:64], i1.cpu().numpy())
    np.testing.assert_array_equal(out["1"][1][:64], s1.cpu().numpy())
    ix.close()


def test_rescoring_prune_keeps_the_exact_answer(monkeypatch):
    """Small pools (a sharded round 1), opt-in RR_TC_PRUNE=1: rows more than 2 eps below the m-th bf16 score are not rescored.  The result must
    equal the unpruned tensor path and the exact path, also when many rows crowd the cut-off (near-duplicates)."""
    rr = _rr()
    n, d, b, pool = 300_000, 384, 256, 48
    emb = rr.synth.embeddings(n, d)
    rng = np.random.default_rng(5)
    base = emb[123].copy()
    for r in range(5000, 5040):                       # 40 rows within 1e-3 of row 123: inside the 2-eps band of its query
        v = base + 1e-3 * rng.standard_normal(d).astype(np.float32)
        emb[r] = v / np.linalg.norm(v)
    q = rr.synth.queries(b, d)
    q[7] = base
    ix = rr.engine.HybridIndex(emb, device="cuda:0")
    i0, s0, _ = ix.dense_topk(q, pool, rr._lib.RR_DENSE_EXACT)
    out = {}
    for prune in (True, False):
        if prune:
            monkeypatch.setenv("RR_TC_PRUNE", "1")
        else:
            monkeypatch.delenv("RR_TC_PRUNE", raising=False)
        i2, s2, _ = ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
        assert ix.dense_stats()["path"] == 2 and ix.dense_stats()["shortlist"] <= 512
        np.testing.assert_array_equal(i2.cpu().numpy(), i0.cpu().numpy())
        np.testing.assert_array_equal(s2.cpu().numpy(), s0.cpu().numpy())
    ix.close()
