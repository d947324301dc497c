#!/usr/bin/env python3
"""Which queries of the in-flight test batch does the tensor path fail to certify, per segment schedule?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

n, d, v = 90_000, 128, 3000
c = rr.synth.make_corpus(n, d, v)
rng = np.random.default_rng(3)
base = c.emb[17].copy()
for r in range(1000, 1600):
    x = base + 1e-4 * rng.standard_normal(d).astype(np.float32)
    c.emb[r] = x / np.linalg.norm(x)
q = rr.synth.queries(64, d)
for g in ("4", "8", "0"):
    os.environ["RR_TC_GROWTH"] = g
    ix = rr.engine.HybridIndex(c.emb, device="cuda:0")
    idx, sims, cnt, unc = ix.dense_topk(q, 150, rr._lib.RR_DENSE_TENSOR, want_uncertified=True)
    torch.cuda.synchronize()
    u = torch.nonzero(unc).view(-1).tolist()
    print("growth", g, "stats", ix.dense_stats(), "uncertified", u)
    s = (c.emb @ q.T)
    for qi in u:
        col = np.sort(s[:, qi])[::-1]
        print("   q", qi, "cluster score", float(s[1000, qi]), "rank of cluster", int((s[:, qi] > s[1000, qi] + 1e-3).sum()),
              "150th", float(col[149]), "416th", float(col[415]), "gap", float(col[149] - col[415]))
    ix.close()
