#!/usr/bin/env python3
"""Device time per kernel class of a single-query hybrid search (library CUDA-event profile, graphs off), for the knobs
of the latency path: top-k tree geometry (RR_TOPK_*), warp-per-candidate BM25 gather (RR_BM25_CAND_WARP).

usage: probe_c1_kernels.py [n_docs ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

D, V, L, K, REPS = 384, 20_000, 4, 10, 200
KNOBS = ("RR_TOPK_TREE_MIN_N", "RR_TOPK_CHUNK_MAX", "RR_TOPK_MID_CHUNK", "RR_TOPK_FINAL_MAX", "RR_BM25_CAND_WARP",
         "RR_NO_CHUNKED_TOPK")
VARIANTS = [
    ("default", {}),
    ("gather: thread per candidate", {"RR_BM25_CAND_WARP": "0"}),
    ("tree from 1024, chunk<=1024, final<=2048", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "1024", "RR_TOPK_MID_CHUNK": "2048", "RR_TOPK_FINAL_MAX": "2048"}),
    ("tree from 1024, chunk<=2048, final<=2048", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "2048", "RR_TOPK_MID_CHUNK": "2048", "RR_TOPK_FINAL_MAX": "2048"}),
    ("tree from 1024, chunk<=2048, final<=4096", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "2048", "RR_TOPK_MID_CHUNK": "4096", "RR_TOPK_FINAL_MAX": "4096"}),
    ("tree from 1024, chunk<=4096, final<=4096", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "4096", "RR_TOPK_MID_CHUNK": "4096", "RR_TOPK_FINAL_MAX": "4096"}),
    ("tree from 1024, chunk<=4096, final<=8192", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "4096", "RR_TOPK_MID_CHUNK": "8192", "RR_TOPK_FINAL_MAX": "8192"}),
    ("tree from 1024, chunk<=8192, final<=16384", {"RR_TOPK_TREE_MIN_N": "1024", "RR_TOPK_CHUNK_MAX": "8192"}),
]


def probe(n):
    c = rr.synth.make_corpus(n, D, V)
    q = rr.synth.queries(REPS, D)
    qt = rr.synth.query_terms(REPS, L, c.doc_offsets, c.token_ids, V).astype(np.int32)
    nt = np.full(1, L, dtype=np.int32)
    fusion = rr.engine.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0)
    ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), V,
                               c.n_reviews, c.avg_stars)
    rows = np.empty((1, K), np.int64); fin = np.empty((1, K), np.float32)
    ref = None
    for label, env in VARIANTS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        for i in range(10):
            ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows, out_final=fin)
        rr.engine.profile_enable(True)
        rr.engine.profile_collect()
        got = []
        for i in range(REPS):
            ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows, out_final=fin)
            got.append((rows.copy(), fin.copy()))
        prof = rr.engine.profile_collect()
        rr.engine.profile_enable(False)
        if ref is None:
            ref = got
        assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(ref, got)), label
        items = {k: v[0] / REPS * 1e3 for k, v in prof.items() if v[1]}
        print(f"n = {n:8d}  {label:44s}", " ".join(f"{k} {v:5.1f}" for k, v in items.items()),
              f"| sum {sum(items.values()):.1f} us", flush=True)
    for k in KNOBS:
        os.environ.pop(k, None)
    ix.close()


for n_docs in [int(a) for a in sys.argv[1:]] or [10_000, 100_000, 1_000_000]:
    probe(n_docs)
