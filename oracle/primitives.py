"""NumPy restatement of the reference's scoring primitives.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows.  The same arithmetic appears three
times in the reference (utils.py, app/test.py, app/app_product_search.py); the copies
are identical except where noted.  PINNED: tests/test_oracle_vs_reference.py runs these
against the reference's own functions when /root/reference is present, and
tests/golden/primitives.npz holds outputs of the reference captured by
tests/golden/make_golden.py.
"""
from __future__ import annotations

import math
import re
from typing import List, Tuple

import numpy as np

# utils.py:11-12 / app/test.py:32-33 / app/app_product_search.py:151-152
TOKEN_RE = re.compile(r"[a-z0-9]+(?:'[a-z0-9]+)?")
STOP_WORDS = {"a", "an", "the", "and", "or", "of", "for", "to", "in", "on", "with",
              "is", "are", "it", "this", "that"}


def tokenize_query(query: str) -> List[str]:
    """utils.py:57-60 (= app/test.py:59-60, app/app_product_search.py:189-190)."""
    return [t for t in TOKEN_RE.findall(query.lower()) if t not in STOP_WORDS]


def l2_normalize(x: np.ndarray, axis: int = 1, eps: float = 1e-12) -> np.ndarray:
    """utils.py:40-44: x / max(||x||_2, eps) along `axis` (dtype of x is kept)."""
    n = np.linalg.norm(x, axis=axis, keepdims=True)
    return x / np.maximum(n, eps)


def minmax_normalize(x: np.ndarray) -> np.ndarray:
    """utils.py:46-55 (= _minmax app/app_product_search.py:182-187).

    lo/hi are Python floats, so under NumPy-2 weak-scalar promotion an f32 input is
    processed in f32 (lo and the divisor are rounded to f32) and an f64 input in f64;
    the result is cast to f32.  Non-finite lo/hi (any NaN or inf) or hi-lo < 1e-12 -> zeros.
    NOTE app/test.py:114-119 `minmax` differs only for EMPTY input (returns it uncast).
    """
    if x.size == 0:
        return x.astype(np.float32)
    lo, hi = float(np.min(x)), float(np.max(x))
    if not math.isfinite(lo) or not math.isfinite(hi) or hi - lo < 1e-12:
        return np.zeros_like(x, dtype=np.float32)
    return ((x - lo) / (hi - lo + 1e-12)).astype(np.float32)


def bayesian_prior(avg: np.ndarray, n: np.ndarray, C: float = 20.0, global_mean=None) -> np.ndarray:
    """utils.py:103-109 (= _bayes_prior app/app_product_search.py:197-199, app/test.py:121-123).
    global_mean defaults to nanmean over the array that is passed (the candidate pool)."""
    g = float(np.nanmean(avg)) if global_mean is None else float(global_mean)
    return ((avg * n) + (g * C)) / (n + C + 1e-9)


def trust_score_from_reviews(n: np.ndarray, min_reviews: int = 8, saturation: int = 50) -> np.ndarray:
    """utils.py:126-133 (= _trust_from_reviews app/app_product_search.py:238-242; the
    Streamlit caller passes sat=80, app/app_product_search.py:303)."""
    ramp = np.clip(n / max(min_reviews, 1), 0, 1)
    satv = np.minimum(1.0, np.log1p(n) / np.log1p(max(saturation, 1)))
    return (0.6 * ramp + 0.4 * satv).astype(np.float32)


def cosine_similarity_search(qvec: np.ndarray, mat: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:111-124 (= cosine_search app/test.py:125-132, _cosine_pool
    app/app_product_search.py:192-195): sims = mat @ q; k clamped to N; argpartition then
    argsort of the k survivors, descending.  Order among equal sims is unspecified."""
    sims = mat @ qvec
    if top_k >= len(sims):
        top_k = len(sims)
    idx = np.argpartition(-sims, top_k - 1)[:top_k]
    idx = idx[np.argsort(-sims[idx])]
    return idx, sims[idx]


def cosine_topk_canonical(qvec: np.ndarray, mat: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Same result set as `cosine_similarity_search`, with the repo's documented tie
    policy (similarity descending, then row index ascending).  Used by parity tests so
    that ties compare deterministically."""
    sims = mat @ qvec
    top_k = min(int(top_k), len(sims))
    order = np.lexsort((np.arange(len(sims)), -sims.astype(np.float64)))[:top_k]
    return order.astype(np.int64), sims[order]
