#!/usr/bin/env python3
"""Where a single-query host call (configs[0] shape) spends its time: the Python wrapper, the C call, the device kernels.

usage: probe_c1_breakdown.py [n_docs]"""
import ctypes as C
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import review_recommender_b200 as rr
from review_recommender_b200.engine import _ptr, _stream, check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
D, V, L, K, REPS = 384, 20_000, 4, 10, 400
c = rr.synth.make_corpus(n, D, V)
q = rr.synth.queries(REPS, D)
qt = rr.synth.query_terms(REPS, L, c.doc_offsets, c.token_ids, V).astype(np.int32)
nt = np.full(1, L, dtype=np.int32)
fusion = rr.engine.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0)
ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), V,
                           c.n_reviews, c.avg_stars)
rows = np.empty((1, K), np.int64); fin = np.empty((1, K), np.float32)


def pct(ts):
    ts = np.sort(np.asarray(ts[20:])) * 1e6
    return "p50 %.1f us  p99 %.1f us" % (ts[len(ts) // 2], ts[int(len(ts) * .99)])


ts = []
for i in range(REPS):
    t0 = time.perf_counter()
    ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows, out_final=fin)
    ts.append(time.perf_counter() - t0)
print(f"n = {n}: python wrapper + C call   ", pct(ts))

p = fusion.to_c()
args = [(_ptr(np.ascontiguousarray(q[i:i + 1])), _ptr(np.ascontiguousarray(qt[i:i + 1]))) for i in range(REPS)]
keep = [(np.ascontiguousarray(q[i:i + 1]), np.ascontiguousarray(qt[i:i + 1])) for i in range(REPS)]
args = [(_ptr(a), _ptr(b)) for a, b in keep]
pn, pr, pf, st, ref = _ptr(nt), _ptr(rows), _ptr(fin), _stream(), C.byref(p)
fn, h = ix.lib.rr_hybrid_search_host, ix._h
ts = []
for i in range(REPS):
    a, b = args[i]
    t0 = time.perf_counter()
    rc = fn(h, a, b, pn, 1, L, ref, rr._lib.RR_DENSE_AUTO, pr, pf, st)
    ts.append(time.perf_counter() - t0)
    check(rc)
print(f"n = {n}: C call alone (ctypes)     ", pct(ts))

os.environ["RR_NO_GRAPHS"] = "1"
ts = []
for i in range(REPS):
    a, b = args[i]
    t0 = time.perf_counter()
    rc = fn(h, a, b, pn, 1, L, ref, rr._lib.RR_DENSE_AUTO, pr, pf, st)
    ts.append(time.perf_counter() - t0)
    check(rc)
print(f"n = {n}: C call alone, no graph    ", pct(ts))

rr.engine.profile_enable(True)
rr.engine.profile_collect()
for i in range(100):
    a, b = args[i]
    check(fn(h, a, b, pn, 1, L, ref, rr._lib.RR_DENSE_AUTO, pr, pf, st))
prof = rr.engine.profile_collect()
rr.engine.profile_enable(False)
print(f"n = {n}: device time per class     ", " ".join(f"{k} {v[0] * 10:.1f} us x{v[1] // 100}" for k, v in prof.items() if v[1]),
      "| sum %.1f us" % (sum(v[0] for v in prof.values()) * 10))

# the floor of this box: an empty stream round trip (one 4-byte H2D + D2H through pinned memory + synchronize)
hbuf = torch.zeros(16, dtype=torch.float32).pin_memory()
dbuf = torch.zeros(16, dtype=torch.float32, device="cuda")
ts = []
for i in range(REPS):
    t0 = time.perf_counter()
    dbuf.copy_(hbuf, non_blocking=True)
    hbuf.copy_(dbuf, non_blocking=True)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"n = {n}: torch H2D + D2H + sync    ", pct(ts))
ix.close()
