"""Headless restatement of the reference's two search drivers.  TEST INFRASTRUCTURE ONLY.

`run_search_core`  follows `run_search`  app/app_product_search.py:245-317 (Streamlit, production)
`cli_search_core`  follows `search(args)` app/test.py:228-309            (CLI)

with the parts that are outside the hot path replaced by inputs:
  * the sentence-transformer encode (app/app_product_search.py:250-251, app/test.py:231-233)
    -> `qvec` is passed in;
  * the cross-encoder (app/app_product_search.py:271-282, app/test.py:262-271)
    -> `rerank_fn(texts) -> scores` (None = "model unavailable": zeros, as :275 / app/test.py:221-222);
  * best-review snippets (:285-294, app/test.py:274-288) -> `best_scores` per candidate or None;
  * attribute gates (:297-302, app/test.py:291-297) -> `gate_fn(agg_text) -> factor` or None (=1.0).
pandas is used exactly where the reference uses it so that dtype promotion (f32 columns vs
the f64 `0.0` columns, f64 `_prior`) and `sort_values` behave identically.

PINNED by tests/golden/make_golden.py, which runs the reference's own `run_search` (under a
stub `streamlit`) and `search` on seeded inputs and stores their outputs.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import pandas as pd

from .primitives import (bayesian_prior, cosine_similarity_search, minmax_normalize,
                         trust_score_from_reviews)


def bm25_for_candidates(bm25, skus: Optional[Sequence[str]], tokens: List[str],
                        cand_skus: Sequence[str]) -> np.ndarray:
    """_bm25_for_candidates app/app_product_search.py:201-208: zeros if no index or no
    tokens; else full get_scores cast to f32, gathered through a SKU dict (duplicate SKUs:
    last one wins; SKU absent from the BM25 blob: 0.0)."""
    if bm25 is None:
        return np.zeros(len(cand_skus), dtype=np.float32)
    if not tokens:
        return np.zeros(len(cand_skus), dtype=np.float32)
    scores_all = np.array(bm25.get_scores(tokens), dtype=np.float32)
    by_sku = {skus[i]: scores_all[i] for i in range(len(skus))}
    return np.array([by_sku.get(str(s), 0.0) for s in cand_skus], dtype=np.float32)


def ensure_same_order(meta_skus: Sequence[str], bm25_skus: Sequence[str]):
    """app/test.py:159-166: permutation that reorders BM25 docs into meta order, or None
    if some meta SKU is missing from the BM25 blob."""
    idx_map = {s: i for i, s in enumerate(bm25_skus)}
    try:
        return [idx_map[str(s)] for s in meta_skus]
    except KeyError:
        return None


def bm25_scores(bm25, tokens: List[str], order_idx, top_idx: np.ndarray) -> np.ndarray:
    """app/test.py:168-173."""
    scores_all = np.array(bm25.get_scores(tokens), dtype=np.float32)
    if order_idx is not None:
        scores_all = scores_all[np.array(order_idx)]
    return scores_all[top_idx]


def _prior_block(cand: pd.DataFrame, prior_C: float):
    # app/app_product_search.py:264-268 == app/test.py:255-259
    n = pd.to_numeric(cand.get("n_reviews", pd.Series([np.nan] * len(cand))), errors="coerce").fillna(0).values
    r = pd.to_numeric(cand.get("avg_stars", pd.Series([np.nan] * len(cand))), errors="coerce").fillna(np.nan).values
    prior_rating = bayesian_prior(r, n, C=prior_C)
    prior_volume = np.log1p(n) / (np.log1p(n).max() + 1e-9)
    return n, minmax_normalize(prior_rating) * 0.7 + 0.3 * prior_volume


def run_search_core(qvec: np.ndarray, V: np.ndarray, meta: pd.DataFrame,
                    bm25=None, bm25_skus: Optional[Sequence[str]] = None,
                    tokens: Optional[List[str]] = None,
                    k: int = 10, rerank_k: int = 0,
                    w_dense: float = 0.55, w_bm25: float = 0.20, w_rerank: float = 0.20,
                    w_prior: float = 0.20, w_best: float = 0.10,
                    prior_C: float = 20.0, min_reviews: int = 8,
                    rerank_fn: Optional[Callable] = None,
                    best_scores_fn: Optional[Callable] = None,
                    gate_fn: Optional[Callable] = None):
    """Returns (top-k DataFrame in rank order, full pool DataFrame before the sort)."""
    pool = max(k, rerank_k, 150)                                              # :253
    cand_idx, dense_scores = cosine_similarity_search(qvec, V, pool)          # :254 (_cosine_pool)
    cand = meta.iloc[cand_idx].reset_index(drop=True).copy()                  # :255
    cand["_row"] = np.asarray(cand_idx, dtype=np.int64)                       # (oracle bookkeeping)
    cand["_dense_raw"] = dense_scores.astype(np.float32)
    cand["_dense"] = minmax_normalize(dense_scores.astype(np.float32))        # :256

    bm25_raw = bm25_for_candidates(bm25, bm25_skus, tokens or [], cand["sku"].astype(str).tolist())  # :259-260
    cand["_bm25_raw"] = bm25_raw
    cand["_bm25"] = minmax_normalize(bm25_raw)                                # :261

    n, prior = _prior_block(cand, prior_C)                                    # :264-268
    cand["_prior"] = prior

    if rerank_k > 0:                                                          # :271-280
        rr_k = min(rerank_k, len(cand))
        if rerank_fn is None:
            rr = np.zeros(rr_k, dtype=np.float32)
        else:
            rr_texts = cand["agg_text"].astype(str).str.slice(0, 2000).tolist()[:rr_k]
            rr = np.array(rerank_fn(rr_texts), dtype=np.float32)
        z = np.zeros(len(cand), dtype=np.float32)
        z[:rr_k] = minmax_normalize(rr)
        cand["_rerank"] = z
    else:
        cand["_rerank"] = 0.0                                                 # :282 (an f64 column)

    best_contrib = np.zeros(len(cand), dtype=np.float32)                      # :288-294
    if best_scores_fn is not None:
        raw = best_scores_fn(cand["sku"].astype(str).tolist())                # list of float or None
        got = False
        for i, v in enumerate(raw):
            if v is not None:
                best_contrib[i] = v
                got = True
        if got:
            best_contrib = minmax_normalize(best_contrib)
    cand["_best"] = best_contrib

    if gate_fn is None:                                                       # :297-302
        gate_vals = [1.0] * len(cand)
    else:
        gate_vals = [gate_fn(t) for t in cand["agg_text"].astype(str).str.slice(0, 6000).tolist()]
    cand["_gate"] = np.array(gate_vals, dtype=np.float32)
    cand["_trust"] = trust_score_from_reviews(n, min_reviews=min_reviews, saturation=80)   # :303

    final = (w_dense * cand["_dense"].values + w_bm25 * cand["_bm25"].values +            # :306-310
             w_rerank * cand["_rerank"].values + w_prior * cand["_prior"].values +
             w_best * cand["_best"].values).astype(np.float32)
    final = final * cand["_trust"].values * cand["_gate"].values
    cand["_final"] = final
    top = cand.sort_values("_final", ascending=False).head(k).reset_index(drop=True)       # :312
    return top, cand


def cli_search_core(qvec: np.ndarray, V: np.ndarray, meta: pd.DataFrame,
                    bm25=None, bm25_skus: Optional[Sequence[str]] = None,
                    tokens: Optional[List[str]] = None,
                    k: int = 10, rerank_k: int = 50,
                    w_dense: float = 0.55, w_bm25: float = 0.15, w_rerank: float = 0.15,
                    w_prior: float = 0.10, w_best: float = 0.05,
                    prior_C: float = 20.0,
                    rerank_fn: Optional[Callable] = None,
                    best_scores_fn: Optional[Callable] = None,
                    gate_fn: Optional[Callable] = None):
    """app/test.py:238-309.  Differences from the Streamlit driver: pool floor 100, BM25
    aligned by permutation, no trust factor, `_bm25 = 0.0` (f64) when BM25 is absent."""
    topK0 = max(k, rerank_k, 100)                                             # :238
    cand_idx, dense_scores = cosine_similarity_search(qvec, V, topK0)         # :239
    cand = meta.iloc[cand_idx].reset_index(drop=True).copy()
    cand["_row"] = np.asarray(cand_idx, dtype=np.int64)
    cand["_dense_raw"] = dense_scores.astype(np.float32)
    cand["_dense"] = dense_scores.astype(np.float32)                          # :241

    if bm25 is not None:                                                      # :244-252
        order = ensure_same_order(meta["sku"].astype(str).tolist(), bm25_skus)
        cand["_bm25_raw"] = bm25_scores(bm25, tokens or [], order, cand_idx)
        cand["_bm25"] = _cli_minmax(cand["_bm25_raw"].values)
    else:
        cand["_bm25"] = 0.0

    n, prior = _prior_block(cand, prior_C)                                    # :255-259
    cand["_prior"] = prior

    if rerank_k > 0:                                                          # :262-271
        k_rr = min(rerank_k, len(cand))
        if rerank_fn is None:
            rr_scores = np.zeros(k_rr, dtype=np.float32)
        else:
            rr_texts = cand["agg_text"].astype(str).str.slice(0, 2000).tolist()[:k_rr]
            rr_scores = np.array(rerank_fn(rr_texts), dtype=np.float32)
        zeros = np.zeros(len(cand), dtype=np.float32)
        zeros[:k_rr] = _cli_minmax(rr_scores)
        cand["_rerank"] = zeros
    else:
        cand["_rerank"] = 0.0

    cand["_dense"] = _cli_minmax(cand["_dense"].values)                       # :280

    best_contrib = np.zeros(len(cand), dtype=np.float32)                      # :283-289
    if best_scores_fn is not None:
        raw = best_scores_fn(cand["sku"].astype(str).tolist())
        got = False
        for i, v in enumerate(raw):
            if v is not None:
                best_contrib[i] = v
                got = True
        if got:
            best_contrib = _cli_minmax(best_contrib)
    cand["_bestrev"] = best_contrib

    if gate_fn is None:                                                       # :292-297
        gate_vals = [1.0] * len(cand)
    else:
        gate_vals = [gate_fn(t) for t in cand["agg_text"].astype(str).str.slice(0, 6000).tolist()]
    cand["_gate"] = np.array(gate_vals, dtype=np.float32)

    final = (w_dense * cand["_dense"].values + w_bm25 * cand["_bm25"].values +            # :300-306
             w_rerank * cand["_rerank"].values + w_prior * cand["_prior"].values +
             w_best * cand["_bestrev"].values).astype(np.float32)
    cand["_final"] = final * cand["_gate"].values                             # :308
    top = cand.sort_values("_final", ascending=False).head(k).reset_index(drop=True)       # :309
    return top, cand


def _cli_minmax(arr: np.ndarray) -> np.ndarray:
    """app/test.py:114-119 (same arithmetic as minmax_normalize; empty input is returned as is)."""
    if arr.size == 0:
        return arr
    return minmax_normalize(arr)


def canonical_topk(pool_df: pd.DataFrame, k: int) -> pd.DataFrame:
    """The pool sorted under the repo's documented tie policy: `_final` descending (NaN last),
    then pool position ascending (pool position = dense rank)."""
    f = pool_df["_final"].values.astype(np.float64)
    key = np.where(np.isnan(f), -np.inf, f)
    order = np.lexsort((np.arange(len(f)), -key))
    return pool_df.iloc[order[:k]].reset_index(drop=True)
