#!/usr/bin/env python3
"""K1 sweep at BASELINE.json configs[3]: BM25-only, 20M docs, 200k Zipf vocabulary, 16-term queries,
get_scores mode (every doc scored), B in {1, 64, 1024}.  Reports achieved HBM GB/s against the
algorithmic bytes of SURVEY.md section 8d:  8 B per posting of the query's terms + 4 B per doc per query.

    python tools/bench_bm25.py [--docs 20000000] [--vocab 200000] [--terms 16] [--tile-docs 12288,8192]
                               [--rings 256x4,512x4] [--batches 1,64,1024] [--out profiles/r02_bm25_c4.json]

--tile-docs and --rings take lists: the index is rebuilt per tile size and every ring geometry
(STAGE_UNITS x NSTAGE, 16-byte units per chunk x chunks in flight per CTA) is timed on it.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=20_000_000)
    ap.add_argument("--vocab", type=int, default=200_000)
    ap.add_argument("--terms", type=int, default=16)
    ap.add_argument("--tile-docs", default="12288")
    ap.add_argument("--rings", default="", help="e.g. 256x4,512x4 (default: the library's default)")
    ap.add_argument("--batches", default="1,64,1024")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import ctypes as C
    import os
    import torch
    import review_recommender_b200 as rr
    import bench as B
    dev = torch.device("cuda:0")
    cfg = dict(docs=args.docs, dim=8, vocab=args.vocab, terms=args.terms)
    t0 = time.perf_counter()
    emb, offs, toks, nrev, avg = B.device_shard(cfg, 0, args.docs, dev)
    t1 = time.perf_counter()
    peaks = B.load_peaks()
    batches = [int(x) for x in args.batches.split(",")]
    qt = rr.synth.query_terms(max(batches), args.terms, offs.cpu().numpy(), toks.cpu().numpy(), args.vocab).astype(np.int32)
    lib = rr._lib.load()
    out = {"docs": args.docs, "vocab": args.vocab, "terms": args.terms, "gen_s": t1 - t0, "hbm_peak_gbs": peaks["hbm_gbs"],
           "peak_source": peaks["source"], "bytes_model": "8 B per posting of the query's terms + 4 B per doc per query",
           "runs": []}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tile in [int(x) for x in args.tile_docs.split(",")]:
        t2 = time.perf_counter()
        gb = rr.engine.GpuIndexBuilder(offs, toks, args.vocab, tile)       # corpus is in device memory
        stats = gb.local_stats().finalize()
        ix = rr.engine.HybridIndex(emb, None, None, args.vocab, device=dev, stats=stats, postings=gb.finish(stats),
                                   make_bf16=False)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t2
        ib = ix.index_bytes()
        for ring in (args.rings.split(",") if args.rings else [""]):
            if ring:
                su, ns = ring.split("x")
                os.environ["RR_BM25_STAGE_UNITS"], os.environ["RR_BM25_STAGES"] = su, ns
            run_rec = {"tile_docs": tile, "ring": ring or "default", "build_s": build_s, "index_bytes": ib,
                       "directory_frac_of_postings": ib["directory"] / max(ib["postings"], 1), "sweep": []}
            for b in batches:
                ids = torch.from_numpy(qt[:b]).to(dev)
                nts = torch.full((b,), args.terms, dtype=torch.int32, device=dev)
                ld = (args.docs + 3) // 4 * 4
                buf = torch.empty((b, ld), dtype=torch.float32, device=dev)

                def run():
                    rr._lib.check(lib.rr_bm25_get_scores(ix._h, C.c_void_p(ids.data_ptr()), C.c_void_p(nts.data_ptr()), b,
                                                         args.terms, C.c_void_p(buf.data_ptr()), ld,
                                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(args.reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                postings = int(stats.df[qt[:b]].sum())
                nbytes = 8 * postings + 4 * args.docs * b
                run_rec["sweep"].append({"B": b, "ms": ms, "postings_per_query": postings / b, "algorithmic_bytes": nbytes,
                                         "achieved_gbs": nbytes / ms / 1e6,
                                         "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"],
                                         "queries_per_s": b / ms * 1e3})
                del buf
            out["runs"].append(run_rec)
            print(json.dumps(run_rec), file=sys.stderr, flush=True)
        ix.close()
        del ix, gb
        torch.cuda.empty_cache()
    text = json.dumps(out)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
