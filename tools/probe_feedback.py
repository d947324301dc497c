#!/usr/bin/env python3
"""Shortlist feedback on a configs[4]-shaped shard: uncertified queries per call and shortlist length over repeated
synchronous rr_dense_topk calls.   python tools/probe_feedback.py [docs] [dim] [B] [pool]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import review_recommender_b200 as rr
docs, dim, B, pool = (int(x) for x in (sys.argv[1:5] + ["6250000", "768", "8192", "192"][len(sys.argv) - 1:]))
emb = torch.randn((docs, dim), device="cuda")
emb /= emb.norm(dim=1, keepdim=True)
ix = rr.engine.HybridIndex(emb, device="cuda:0")
q = torch.from_numpy(rr.synth.queries(B, dim)).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    e0.record()
    ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
    e1.record()
    torch.cuda.synchronize()
    print(it, round(e0.elapsed_time(e1), 2), "ms", ix.dense_stats())
