"""K1 parity: rr_bm25_get_scores / rr_bm25_candidates against the oracle BM25Okapi (float64)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.bm25_okapi import BM25Okapi, BM25OkapiCSR, flatten_corpus
from tests.parity import BM25_RTOL


def _rr():
    import review_recommender_b200 as rr
    return rr


def _index(offs, toks, v, tile=16384, forward_index=True):
    rr = _rr()
    n = offs.shape[0] - 1
    emb = np.zeros((n, 4), dtype=np.float32)
    emb[:, 0] = 1.0
    return rr.engine.HybridIndex(emb, offs, toks, v, device="cuda:0", tile_docs=tile, make_bf16=False,
                                 forward_index=forward_index)


def _check(ix, csr, term_lists, atol=0.0):
    rr = _rr()
    ids, nt = rr.engine.HybridIndex.pack_terms(term_lists)
    got = ix.bm25_get_scores(ids, nt).cpu().numpy()
    assert got.shape == (len(term_lists), csr.corpus_size)
    for i, terms in enumerate(term_lists):
        want = csr.get_scores(terms)
        np.testing.assert_allclose(got[i], want, rtol=BM25_RTOL, atol=atol)
        assert np.array_equal(got[i] == 0, want.astype(np.float32) == 0)
    return got


def test_reference_fixture_corpus(golden_dir):
    fx = json.loads((golden_dir / "bm25_fixture.json").read_text())
    for corpus, scores in ((fx["fixture_corpus"], fx["fixture_scores"]), (fx["neg_corpus"], fx["neg_scores"])):
        offs, ids, vocab = flatten_corpus(corpus)
        ix = _index(offs, ids, len(vocab), tile=16)
        lut = {w: i for i, w in enumerate(vocab)}
        for q, want in scores.items():
            terms = [lut.get(w, -1) for w in q.split()]
            t, n = _rr().engine.HybridIndex.pack_terms([terms])
            got = ix.bm25_get_scores(t, n).cpu().numpy()[0]
            np.testing.assert_allclose(got, want, rtol=BM25_RTOL, atol=1e-8)
        ix.close()


@pytest.mark.parametrize("n,v,tile", [(1, 5, 16), (1000, 300, 64), (20000, 2000, 4096), (50001, 5000, 16384)])
def test_get_scores_matches_oracle(n, v, tile):
    rr = _rr()
    offs, toks = rr.synth.corpus_tokens(n, v)
    csr = BM25OkapiCSR(offs, toks, v)
    ix = _index(offs, toks, v, tile)
    rng = np.random.default_rng(n)
    qt = rr.synth.query_terms(6, min(4, v), offs, toks, v)
    term_lists = [list(map(int, r)) for r in qt]
    term_lists += [[0], [0, 0, 0], [v - 1, -1, v + 7, 1], [], list(map(int, rng.integers(0, v, 16)))]
    term_lists.append(list(map(int, rng.integers(0, min(v, 50), 150))))      # > 64 terms: multi-pass staging
    got = _check(ix, csr, term_lists)
    # candidate mode is bit-identical to the full scores at those rows
    cand = rng.integers(0, n, size=(len(term_lists), 37)).astype(np.int64)
    cand[0, :3] = [-1, n, 0]
    ids, nt = rr.engine.HybridIndex.pack_terms(term_lists)
    ix_search = _index(offs, toks, v, tile, forward_index=False)     # candidate mode by postings search
    for index in (ix, ix_search):                                    # ... and through the forward index
        cm = index.bm25_candidates(ids, nt, cand).cpu().numpy()
        for i in range(len(term_lists)):
            ok = (cand[i] >= 0) & (cand[i] < n)
            np.testing.assert_array_equal(cm[i][ok], got[i][cand[i][ok]])
            assert np.all(cm[i][~ok] == 0)
    ix_search.close()
    ix.close()


def test_duplicate_tokens_are_summed_per_occurrence():
    rr = _rr()
    offs, toks = rr.synth.corpus_tokens(3000, 200)
    ix = _index(offs, toks, 200, 1024)
    ids, nt = rr.engine.HybridIndex.pack_terms([[5], [5, 5], [5, 9], [9, 5]])
    s = ix.bm25_get_scores(ids, nt).cpu().numpy()
    np.testing.assert_array_equal(s[1], s[0] + s[0])
    # fp32 addition is commutative for two terms: order must not matter here
    np.testing.assert_array_equal(s[2], s[3])
    ix.close()


def test_dict_based_oracle_agrees_on_strings():
    rr = _rr()
    offs, toks = rr.synth.corpus_tokens(800, 120)
    corpus = rr.synth.corpus_as_lists(offs, toks)
    bm = BM25Okapi(corpus)
    ix = _index(offs, toks, 120, 256)
    ids, nt = rr.engine.HybridIndex.pack_terms([[3, 17, 3, 60]])
    got = ix.bm25_get_scores(ids, nt).cpu().numpy()[0]
    want = bm.get_scores(["t4", "t18", "t4", "t61"])
    np.testing.assert_allclose(got, want, rtol=BM25_RTOL)
    ix.close()


def test_fp32_accumulation_error_bound_is_pinned():
    """rank_bm25 adds float64 terms and the callers cast once to float32; K1 / K1c add the correctly rounded fp32
    of every term in fp32, in query-token order.  The deviation is bounded by ~(L + 1) * 2^-24 relative -- far inside
    north_star's 1e-5 -- and K1c equals K1 bit for bit.  16-term queries (configs[3] shape) over a Zipf corpus."""
    import torch
    rr = _rr()
    n, v, l = 60_000, 4000, 16
    offs, toks = rr.synth.corpus_tokens(n, v)
    csr = BM25OkapiCSR(offs, toks, v)
    ix = _index(offs, toks, v, tile=12288)
    qt = rr.synth.query_terms(8, l, offs, toks, v).astype(np.int32)
    nt = np.full(8, l, dtype=np.int32)
    got = ix.bm25_get_scores(qt, nt).cpu().numpy().astype(np.float64)
    worst = 0.0
    for i in range(8):
        want = csr.get_scores(qt[i].tolist())
        nz = want != 0
        worst = max(worst, float(np.max(np.abs(got[i][nz] - want[nz]) / want[nz])))
        assert np.array_equal(got[i] == 0, want == 0)
    bound = (l + 1) * 2.0 ** -24
    print(f"max relative deviation from the float64 sum: {worst:.3e} (bound {bound:.3e})")
    assert worst <= bound
    cand = torch.randint(0, n, (8, 150), device="cuda", dtype=torch.int64)
    full = ix.bm25_get_scores(qt, nt)
    assert torch.equal(ix.bm25_candidates(qt, nt, cand), torch.gather(full, 1, cand))
    ix.close()
