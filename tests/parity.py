"""Shared parity helpers: run the CPU oracle and the CUDA path on the same seeded inputs and
compare under the tolerances BASELINE.json's north_star states:

  * BM25 and fused scores  within 1e-5 relative (fp32)
  * dense cosine           within 1e-6 absolute (exact / rescored path)
  * top-k ids              bit-exact, except ties within tolerance (documented tie policy:
                           score descending, then row / pool position ascending)
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import pandas as pd

from oracle.bm25_okapi import BM25OkapiCSR
from oracle.hybrid import cli_search_core, run_search_core

BM25_RTOL = 1e-5
FUSED_RTOL = 1e-5
DENSE_ATOL = 1e-6


class CSRBm25Adapter:
    """Gives BM25OkapiCSR the `get_scores(list[str])` face of rank_bm25 for the oracle drivers;
    token strings are synth's f"t{rank}" (term id = rank-1)."""

    def __init__(self, csr: BM25OkapiCSR):
        self.csr = csr

    def get_scores(self, tokens: Sequence[str]) -> np.ndarray:
        ids = []
        for t in tokens:
            try:
                ids.append(int(t[1:]) - 1 if t.startswith("t") else -1)
            except ValueError:
                ids.append(-1)
        return self.csr.get_scores(ids)


def assert_ids_match_modulo_ties(got_ids, got_scores, ref_ids, ref_scores, tol, what=""):
    """Same ids position by position, except inside groups whose reference scores are within
    `tol` of each other (there any permutation / boundary choice is accepted as long as the
    returned score agrees)."""
    got_ids, ref_ids = np.asarray(got_ids), np.asarray(ref_ids)
    got_scores, ref_scores = np.asarray(got_scores, dtype=np.float64), np.asarray(ref_scores, dtype=np.float64)
    assert got_ids.shape == ref_ids.shape, (what, got_ids.shape, ref_ids.shape)
    n_exact = int(np.sum(got_ids == ref_ids))
    for i in np.nonzero(got_ids != ref_ids)[0]:
        # the id differs: legal only if the reference has another entry (or the cut-off) tied with it
        near = np.abs(ref_scores - ref_scores[i]) <= tol
        assert near.sum() > 1 or i == len(ref_ids) - 1, f"{what}: id mismatch at {i} without a tie"
        assert abs(got_scores[i] - ref_scores[i]) <= tol, f"{what}: score mismatch at {i}"
    return n_exact


def make_world(n, d, v, b, l, seed_shift=0):
    import review_recommender_b200 as rr
    s = rr.synth
    c = s.make_corpus(n, d, v)
    q = s.queries(b, d)
    qt = s.query_terms(b, l, c.doc_offsets, c.token_ids, v)
    return c, q, qt


def oracle_hybrid(c, q, qt, k, driver="streamlit", **kw) -> List[pd.DataFrame]:
    import review_recommender_b200 as rr
    n = c.emb.shape[0]
    skus = rr.synth.skus(n)
    meta = pd.DataFrame({"sku": skus, "n_reviews": c.n_reviews, "avg_stars": c.avg_stars})
    bm25 = CSRBm25Adapter(BM25OkapiCSR(c.doc_offsets, c.token_ids, c.vocab_size))
    out = []
    for i in range(q.shape[0]):
        toks = [f"t{int(t) + 1}" for t in qt[i] if t >= 0]
        if driver == "streamlit":
            top, pool = run_search_core(q[i], c.emb, meta, bm25, skus, toks, k=k, **kw)
        else:
            top, pool = cli_search_core(q[i], c.emb, meta, bm25, skus, toks, k=k, **kw)
        out.append((top, pool))
    return out


def check_hybrid_against_oracle(n, d, v, b, l, k, device="cuda:0", driver="streamlit", mode=0, **kw):
    import review_recommender_b200 as rr
    c, q, qt = make_world(n, d, v, b, l)
    ix = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, v, c.n_reviews, c.avg_stars, device=device)
    fusion = rr.engine.Fusion(k=k, driver=driver, **kw)
    nterms = np.full(b, l, dtype=np.int32)
    rows, final = ix.hybrid_search_host(q, qt.astype(np.int32), nterms, fusion, mode=mode)
    okw = dict(rerank_k=fusion.rerank_k, w_dense=fusion.w_dense, w_bm25=fusion.w_bm25, w_rerank=fusion.w_rerank,
               w_prior=fusion.w_prior, w_best=fusion.w_best, prior_C=fusion.prior_C)
    if driver == "streamlit":
        okw["min_reviews"] = fusion.min_reviews
    ref = oracle_hybrid(c, q, qt, k, driver, **okw)
    exact = 0
    for i, (top, pool) in enumerate(ref):
        ref_rows = top["_row"].values
        ref_final = top["_final"].values.astype(np.float64)
        kk = len(ref_rows)
        np.testing.assert_allclose(final[i, :kk], ref_final, rtol=FUSED_RTOL, atol=1e-7)
        tol = FUSED_RTOL * max(1e-3, float(np.max(np.abs(ref_final)))) + 1e-7
        exact += assert_ids_match_modulo_ties(rows[i, :kk], final[i, :kk], ref_rows, ref_final, tol, f"query {i}")
        assert np.all(rows[i, kk:] == -1)
    ix.close()
    return exact / max(1, sum(len(t) for t, _ in ref))
