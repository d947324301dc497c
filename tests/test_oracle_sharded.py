"""The sharded oracle (oracle/sharded.py) answers a query exactly like run_search_core over the whole arrays.  CPU only."""
import numpy as np
import pandas as pd
import pytest

import review_recommender_b200 as rr
from oracle.bm25_okapi import BM25OkapiCSR
from oracle.hybrid import cli_search_core, run_search_core
from oracle.sharded import ChunkedBM25Scores, ShardedOracle
from tests.parity import CSRBm25Adapter


def _world(n, d, v, b, l):
    s = rr.synth
    c = s.make_corpus(n, d, v)
    return c, s.queries(b, d), s.query_terms(b, l, c.doc_offsets, c.token_ids, v)


@pytest.mark.parametrize("driver,k,cuts", [("streamlit", 10, [0, 1700, 4100, 6000]), ("streamlit", 100, [0, 6000]),
                                           ("cli", 25, [0, 150, 3000, 3100, 6000])])
def test_sharded_oracle_equals_whole_corpus_oracle(driver, k, cuts):
    n, d, v, b, l = 6000, 48, 900, 6, 4
    c, q, qt = _world(n, d, v, b, l)
    qt[1, 2] = qt[1, 0]                       # a duplicate query term is summed twice
    qt[2, 1] = v + 5                          # unknown term id
    pool = max(k, 150 if driver == "streamlit" else 100)
    so = ShardedOracle(q, qt, v, pool, margin=8)
    lens = np.diff(c.doc_offsets)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        so.add_chunk(lo, c.emb[lo:hi], lens[lo:hi], c.token_ids[c.doc_offsets[lo]:c.doc_offsets[hi]],
                     c.n_reviews[lo:hi], c.avg_stars[lo:hi])
    so.finalize()
    csr = BM25OkapiCSR(c.doc_offsets, c.token_ids, v)
    assert so.avgdl == csr.avgdl and so.average_idf == csr.average_idf
    np.testing.assert_array_equal(so.idf, csr.idf)
    skus = rr.synth.skus(n)
    meta = pd.DataFrame({"sku": skus, "n_reviews": c.n_reviews, "avg_stars": c.avg_stars})
    fn = run_search_core if driver == "streamlit" else cli_search_core
    for i in range(b):
        toks = [f"t{int(t) + 1}" for t in qt[i] if t >= 0]
        want, want_pool = fn(q[i], c.emb, meta, CSRBm25Adapter(csr), skus, toks, k=k, rerank_k=0)
        got, got_pool = so.run(i, k, driver, rerank_k=0)
        np.testing.assert_array_equal(got["_grow"].values, want["_row"].values)
        np.testing.assert_allclose(got["_final"].values, want["_final"].values, rtol=2e-6, atol=1e-7)
        # BM25 at the pool members: same float64 expression, same statistics -> bit-identical
        a = got_pool.set_index("_grow")["_bm25_raw"]
        w = want_pool.set_index("_row")["_bm25_raw"]
        np.testing.assert_array_equal(a.loc[w.index].values, w.values)


def test_chunked_get_scores_is_bit_identical_to_csr():
    n, v = 5000, 700
    offs, toks = rr.synth.corpus_tokens(n, v)
    lens = np.diff(offs)
    ch = ChunkedBM25Scores(v)
    for lo, hi in [(0, 1234), (1234, 1235), (1235, 5000)]:
        ch.add_chunk(lo, lens[lo:hi], toks[offs[lo]:offs[hi]])
    ch.finalize()
    csr = BM25OkapiCSR(offs, toks, v)
    for q in ([3, 17, 3, 250], [0], [699, v + 1, -1], []):
        np.testing.assert_array_equal(ch.get_scores(q), csr.get_scores(q))


def test_chunk_stream_matches_whole_recipe():
    s = rr.synth
    old = s.CHUNK
    s.CHUNK = 1000                           # exercise chunk boundaries without 1 M-row chunks
    try:
        ref = s.make_corpus(3500, 16, 300, row0=700)
        pieces = list(s.chunk_stream(700, 3500, 16, 300, workers=3, per_chunk=lambda p: p.row0))
        assert [p.row0 for p in pieces] == [700, 1000, 2000, 3000, 4000] and all(p.extra == p.row0 for p in pieces)
        np.testing.assert_array_equal(np.concatenate([p.emb for p in pieces]), ref.emb)
        np.testing.assert_array_equal(np.concatenate([p.lens for p in pieces]), np.diff(ref.doc_offsets))
        np.testing.assert_array_equal(np.concatenate([p.token_ids for p in pieces]), ref.token_ids)
        np.testing.assert_array_equal(np.concatenate([p.n_reviews for p in pieces]), ref.n_reviews)
        np.testing.assert_array_equal(s.query_terms_global(5, 4, 3500, 300),
                                      s.query_terms(5, 4, *s.corpus_tokens(1000, 300), 300))
    finally:
        s.CHUNK = old
