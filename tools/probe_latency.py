#!/usr/bin/env python3
"""Single-query latency of rr_hybrid_search_host: with / without CUDA-graph replay, and (corpora of more than 16384
products) with the shared-memory top-k tree against the radix-select pipeline.

usage: probe_latency.py [n_docs ...]        (default: 10000 = configs[0])"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

D, V, L, K, REPS = 384, 20_000, 4, 10, 300


def _env(name, on):
    if on:
        os.environ[name] = "1"
    else:
        os.environ.pop(name, None)


def probe(n):
    c = rr.synth.make_corpus(n, D, V)
    q = rr.synth.queries(REPS, D)
    qt = rr.synth.query_terms(REPS, L, c.doc_offsets, c.token_ids, V).astype(np.int32)
    nt = np.full(1, L, dtype=np.int32)
    fusion = rr.engine.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0)
    ref = None
    for graphs, tree in ((True, True), (False, True), (True, False)):
        if not tree and n <= 16384:
            continue                                   # one-kernel top-k either way
        _env("RR_NO_GRAPHS", not graphs)
        _env("RR_NO_CHUNKED_TOPK", not tree)
        ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), V,
                                   c.n_reviews, c.avg_stars)
        rows = np.empty((1, K), np.int64); fin = np.empty((1, K), np.float32)
        out = []
        for i in range(REPS):
            rr.engine.launch_count(reset=True)
            t0 = time.perf_counter()
            ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows, out_final=fin)
            out.append((time.perf_counter() - t0, rr.engine.launch_count(), rows.copy()))
        lat = np.sort([o[0] for o in out[20:]]) * 1e6
        print("n = %8d" % n, "graphs" if graphs else "plain ", "tree " if tree else "radix",
              "p50 %.1f us  p99 %.1f us  launches/query %d" % (lat[len(lat) // 2], lat[int(len(lat) * .99)], out[-1][1]),
              flush=True)
        if ref is None:
            ref = [o[2] for o in out]
        else:
            assert all(np.array_equal(a, o[2]) for a, o in zip(ref, out)), "every variant must return the same results"
        ix.close()
    _env("RR_NO_GRAPHS", False)
    _env("RR_NO_CHUNKED_TOPK", False)


for n_docs in [int(a) for a in sys.argv[1:]] or [10_000]:
    probe(n_docs)
