"""Shared loader of the snippet golden world (tests/golden/snippet_cases.* made by make_golden.py)."""
from __future__ import annotations

import json

import numpy as np
import pandas as pd

from oracle import primitives as P
import review_recommender_b200 as rr


def load(golden_dir):
    cases = json.loads((golden_dir / "search_cases.json").read_text())
    snip = json.loads((golden_dir / "snippet_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    s = np.load(golden_dir / "snippet_cases.npz")
    n = z["emb"].shape[0]
    skus = rr.synth.skus(n)
    prod, E, stars = s["rev_product"], s["rev_emb"], s["rev_stars"]
    sku_col = [skus[i] if i >= 0 else f"UNKNOWN{j % 5}" for j, i in enumerate(prod)]
    text = [f"review {j} of {sku_col[j]} " + "lorem ipsum " * (j % 70) for j in range(len(prod))]   # make_reviews
    reviews = pd.DataFrame({"sku": sku_col, "text": text, "stars": stars, "embedding": list(E)})
    corpus = rr.synth.corpus_as_lists(z["doc_offsets"], z["token_ids"])
    perm = z["bm25_perm"]
    meta = pd.DataFrame({"sku": skus, "n_reviews": z["n_reviews"], "avg_stars": z["avg_stars"],
                         "agg_text": ["" for _ in range(n)]})
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    return dict(cases=cases, snip=snip, z=z, skus=skus, reviews=reviews, meta=meta, Vn=Vn,
                bm25_corpus=[corpus[i] for i in perm], bm25_skus=[skus[i] for i in perm],
                table={q: z["queries"][i] for i, q in enumerate(cases["query_strs"])})
