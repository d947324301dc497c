#!/usr/bin/env python3
"""Two exact-path calls at B = 1 and B = 8 over 1 M x 384 for an ncu capture of the GEMV and the top-k tree."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

n, d = 1_000_000, 384
emb = torch.randn((n, d), device="cuda")
emb /= emb.norm(dim=1, keepdim=True)
ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
for b in (1, 8):
    q = torch.from_numpy(rr.synth.queries(b, d)).cuda()
    for _ in range(2):
        ix.dense_topk(q, 150, rr._lib.RR_DENSE_EXACT)
torch.cuda.synchronize()
ix.close()
