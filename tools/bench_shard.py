#!/usr/bin/env python3
"""Per-shard GPU work of round 1 of a row-sharded search, on ONE GPU: rr_shard_tuples (shortlist GEMM + selection +
rescoring + finalize + candidate BM25 written into the send buffer) for the whole batch against a 1/G row shard with the
round-1 pool m = local_pool(pool, G).  This is what every rank of an 8-GPU run does per step besides the exchange and
the fusion of its B/G queries, so kernel schedules can be compared without an 8-GPU box.

    python tools/bench_shard.py [--shards 8] [--docs 10000000] [--batch 4096] [--reps 20]
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--label", default="")
    args = ap.parse_args()
    import torch
    import review_recommender_b200 as rr
    import bench as B
    dev = torch.device("cuda:0")
    n = args.docs // args.shards
    cfg = dict(docs=n, dim=384, vocab=50_000, terms=4)
    emb, offs, toks, nrev, avg = B.device_shard(cfg, 0, n, dev)
    gb = rr.engine.GpuIndexBuilder(offs, toks, 50_000)
    stats = gb.local_stats().finalize()
    ix = rr.engine.HybridIndex(emb, None, None, 50_000, nrev, avg, device=dev, stats=stats, postings=gb.finish(stats))
    q = torch.from_numpy(rr.synth.queries(args.batch, 384)).to(dev)
    qt = torch.from_numpy(rr.synth.query_terms(args.batch, 4, offs.cpu().numpy(), toks.cpu().numpy(), 50_000).astype(np.int32)).to(dev)
    nt = torch.full((args.batch,), 4, dtype=torch.int32, device=dev)
    m = rr.dist.local_pool(150, args.shards)
    send = None
    for _ in range(3):
        send = ix.shard_tuples(q, qt, nt, m, 1, send=send)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        send = ix.shard_tuples(q, qt, nt, m, 1, send=send)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    rr.engine.profile_enable(True)
    rr.engine.profile_collect()
    for _ in range(args.reps):
        send = ix.shard_tuples(q, qt, nt, m, 1, send=send)
    torch.cuda.synchronize()
    prof = rr.engine.profile_collect()
    rr.engine.profile_enable(False)
    out = {"label": args.label, "rows": n, "batch": args.batch, "m": m, "ms_per_call": ms, "dense_path": ix.dense_stats(),
           "kernels": {k: {"ms": v[0] / args.reps, "launches": v[1] / args.reps} for k, v in prof.items() if v[1]}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
