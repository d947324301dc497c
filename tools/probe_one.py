#!/usr/bin/env python3
"""One rr_dense_topk (tensor path) at a given shape, for ncu launch lists: python tools/probe_one.py docs B pool"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import review_recommender_b200 as rr
docs, B, pool = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
emb = torch.randn((docs, 384), device="cuda")
emb /= emb.norm(dim=1, keepdim=True)
ix = rr.engine.HybridIndex(emb, device="cuda:0")
q = torch.from_numpy(rr.synth.queries(B, 384)).cuda()
for _ in range(2):
    ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
torch.cuda.synchronize()
print(ix.dense_stats())
