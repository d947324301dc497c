#!/usr/bin/env python3
"""Single-query (configs[0]) latency of rr_hybrid_search_host with and without CUDA-graph replay."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr

n, d, v, l, k = 10_000, 384, 20_000, 4, 10
c = rr.synth.make_corpus(n, d, v)
q = rr.synth.queries(300, d)
qt = rr.synth.query_terms(300, l, c.doc_offsets, c.token_ids, v).astype(np.int32)
nt = np.full(1, l, dtype=np.int32)
fusion = rr.engine.Fusion(k=k, rerank_k=0, w_rerank=0.0, w_best=0.0)
for graphs in (True, False):
    if graphs:
        os.environ.pop("RR_NO_GRAPHS", None)
    else:
        os.environ["RR_NO_GRAPHS"] = "1"
    ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), v,
                               c.n_reviews, c.avg_stars)
    rows = np.empty((1, k), np.int64); fin = np.empty((1, k), np.float32)
    out = []
    for i in range(300):
        rr.engine.launch_count(reset=True)
        t0 = time.perf_counter()
        ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows, out_final=fin)
        out.append((time.perf_counter() - t0, rr.engine.launch_count(), rows.copy()))
    lat = np.sort([o[0] for o in out[20:]]) * 1e6
    print("graphs" if graphs else "plain ", "p50 %.1f us  p99 %.1f us  launches/query %d" % (lat[len(lat) // 2], lat[int(len(lat) * .99)], out[-1][1]))
    if graphs:
        ref = [o[2] for o in out]
    else:
        assert all(np.array_equal(a, o[2]) for a, o in zip(ref, out)), "graph replay must return the plain path's results"
    ix.close()
