#!/usr/bin/env python3
"""Exact-path kernel classes (fp32 GEMV, row top-k) per batch size, from the library's own CUDA-event profile:
narrow GEMV instantiations (default) against RR_GEMV_WIDE=1, top-k tree against RR_NO_CHUNKED_TOPK=1.

usage: probe_gemv.py [n_docs] [dim]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
emb = torch.randn((n, d), device="cuda")
emb /= emb.norm(dim=1, keepdim=True)
ix = rr.engine.HybridIndex(emb, device="cuda:0", make_bf16=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for b in (1, 8, 16, 32, 64):
    q = torch.from_numpy(rr.synth.queries(b, d)).cuda()
    ref = None
    for label, env in (("default", {}), ("wide gemv", {"RR_GEMV_WIDE": "1"}), ("radix top-k", {"RR_NO_CHUNKED_TOPK": "1"}),
                       ("tree <= 64 rows", {"RR_TOPK_TREE_MAX_ROWS": "64"})):
        for k in ("RR_GEMV_WIDE", "RR_NO_CHUNKED_TOPK", "RR_TOPK_TREE_MAX_ROWS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            out = ix.dense_topk(q, 150, rr._lib.RR_DENSE_EXACT)
        rr.engine.profile_enable(True)
        rr.engine.profile_collect()
        reps = 10
        for _ in range(reps):
            flush.zero_()
            out = ix.dense_topk(q, 150, rr._lib.RR_DENSE_EXACT)
        torch.cuda.synchronize()
        prof = rr.engine.profile_collect()
        rr.engine.profile_enable(False)
        got = (out[0].cpu().numpy(), out[1].cpu().numpy())
        if ref is None:
            ref = got
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]), label
        items = {k: (v[0] / reps * 1e3, v[1] // reps) for k, v in prof.items() if v[1]}
        gemv = items.get("dense_gemv", (0, 0))[0]
        print(f"n = {n} d = {d} B = {b:2d} {label:12s}", " ".join(f"{k} {v[0]:.1f} us x{v[1]}" for k, v in items.items()),
              f"| gemv {4e-3 * n * d * ((b + 7) // 8) / max(gemv, 1e-9):.0f} GB/s", flush=True)
ix.close()
