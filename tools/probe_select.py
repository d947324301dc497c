#!/usr/bin/env python3
"""Per-class device time of rr_dense_topk (tensor path) for a sweep of batch sizes: where does the per-query
selection time go when the batch shrinks?   python tools/probe_select.py [docs] [pool]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import review_recommender_b200 as rr

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 48
emb = torch.randn((docs, 384), device="cuda")
emb /= emb.norm(dim=1, keepdim=True)
ix = rr.engine.HybridIndex(emb, device="cuda:0")
for B in (128, 512, 1024, 2048, 4096):
    q = torch.from_numpy(rr.synth.queries(B, 384)).cuda()
    for _ in range(3):
        ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
    torch.cuda.synchronize()
    rr.engine.profile_enable(True)
    rr.engine.profile_collect()
    reps = 10
    for _ in range(reps):
        ix.dense_topk(q, pool, rr._lib.RR_DENSE_TENSOR)
    torch.cuda.synchronize()
    prof = rr.engine.profile_collect()
    rr.engine.profile_enable(False)
    st = ix.dense_stats()
    print(B, "KP", st["shortlist"], "segs", st["n_segments"],
          {k: (round(v[0] / reps, 3), round(v[0] / max(v[1], 1) * 1000, 1)) for k, v in prof.items() if v[1] > 0})
