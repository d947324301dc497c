"""Best-review snippet scoring, restated.  TEST INFRASTRUCTURE ONLY.

Follows `_best_snippets` app/app_product_search.py:320-370 (Streamlit; text cap 600, any exception
-> {}) and `best_review_snippets` app/test.py:181-215 (CLI; text cap 400), with the parquet read
replaced by the frame itself:

    reviews   DataFrame with columns sku, text, stars, embedding (file order = row order)

PINNED by tests/golden/make_golden.py::golden_snippets, which runs the reference's own two
functions on a parquet written from the same frame and stores their outputs
(tests/golden/snippet_cases.json).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import pandas as pd

from .primitives import l2_normalize


def best_review_snippets(qvec: np.ndarray, cand_skus: List[str], reviews: pd.DataFrame,
                         max_rows: int = 1_000_000, text_cap: int = 400) -> Dict[str, Dict]:
    """app/test.py:181-215 / app/app_product_search.py:320-361.  Also returns, per SKU, `file_pos` =
    the file position of the winning review (oracle bookkeeping, not in the reference's dict)."""
    if "sku" not in reviews.columns:                                          # :190-192 / :329-331
        return {}
    sel = reviews["sku"].astype(str).isin(set(cand_skus))                     # :194 / :333
    sub_meta = reviews[sel]
    if sub_meta.empty:                                                        # :196 / :335-337
        return {}
    emb_series = reviews[["embedding"]].iloc[sub_meta.index]                 # :199 / :341
    if len(sub_meta) > max_rows:                                              # :201-203 / :343-346
        sub_meta = sub_meta.iloc[:max_rows]
        emb_series = emb_series.iloc[:max_rows]
    file_pos = np.asarray(sub_meta.index)
    E = np.stack(emb_series["embedding"].values).astype(np.float32)           # :205 / :348 (raises on 0 rows)
    En = l2_normalize(E, axis=1)                                              # :206 / :349
    sims = En @ qvec                                                          # :207 / :350
    sub_meta = sub_meta.reset_index(drop=True)
    sub_meta["__sim"] = sims
    sub_meta["__file"] = file_pos
    best = {}
    for sku, grp in sub_meta.groupby("sku"):                                  # :212 / :355
        j = int(grp["__sim"].values.argmax())
        row = grp.iloc[j]
        best[str(sku)] = {"score": float(row["__sim"]), "text": str(row["text"])[:text_cap],
                          "stars": float(row.get("stars", np.nan)), "file_pos": int(row["__file"])}
    return best


def streamlit_best_snippets(qvec, cand_skus, reviews, max_rows: int = 300_000) -> Dict[str, Dict]:
    """_best_snippets: same arithmetic, text cap 600, every exception swallowed (:363-367)."""
    try:
        return best_review_snippets(qvec, cand_skus, reviews, max_rows=max_rows, text_cap=600)
    except Exception:
        return {}
