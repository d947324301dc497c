#!/usr/bin/env python3
"""Benchmark of the hybrid-retrieval hot path (BASELINE.json metric: hybrid top-100 queries/sec at
10M x 384 docs, 1/2/4/8 B200; % of HBM / tensor roofline).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                    (the reference's CPU path, rank 0 only)

A step = one pass of the hot path over one batch of B synthetic queries:
dense top-pool (tcgen05 shortlist + exact rescoring) -> BM25 at the candidates -> fusion -> top-k,
against a corpus that is resident in HBM (row-sharded over the N ranks: strong scaling, the corpus
and the batch are fixed as N grows).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

CONFIGS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (fits one B200: 23 GB)
    "c3": dict(docs=10_000_000, dim=384, vocab=50_000, batch=4096, terms=4, k=100,
               workload="configs[2]: 10M products x 384-d, 50k-vocab BM25, batch 4096, hybrid top-100, row-sharded"),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (single query, top-10); here mainly for
    # `--impl reference --config c1`, which then times the WHOLE workload instead of a scaled sample
    "c1": dict(docs=10_000, dim=384, vocab=20_000, batch=1, terms=4, k=10,
               metric="hybrid top-10 queries/sec at 10k x 384 docs",
               workload="configs[0]: 10k products x 384-d, 20k-vocab BM25, single-query hybrid top-10"),
    "c2": dict(docs=1_000_000, dim=384, vocab=50_000, batch=1024, terms=4, k=100,
               workload="configs[1]: 1M products x 384-d, 50k-vocab BM25, batch 1024, hybrid top-100"),
    # BASELINE.json configs[4]: dense-heavy, top-1000 for the reranker (pool = 1000); needs 8 GPUs at full size
    # (use --docs 6250000 --gpus 1 to time one of its eight row shards)
    "c5": dict(docs=50_000_000, dim=768, vocab=50_000, batch=8192, terms=4, k=1000,
               weights=dict(w_dense=0.8, w_bm25=0.1, w_prior=0.1),
               metric="hybrid top-1000 queries/sec at 50M x 768 docs",
               workload="configs[4]: 50M products x 768-d dense-heavy (w_dense 0.8), batch 8192, top-1000, row-sharded"),
}
METRIC = "hybrid top-100 queries/sec at 10M x 384 docs"
UNIT = "queries/s"
OUT = sys.stdout
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--docs", type=int, default=0, help="override corpus size (debug)")
    ap.add_argument("--batch", type=int, default=0, help="override batch size (debug)")
    ap.add_argument("--dense-mode", type=int, default=0, help="0 auto, 1 exact fp32, 2 tensor")
    ap.add_argument("--cpu-sample-docs", type=int, default=100_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=64, help="queries per step of the CPU reference arm")
    ap.add_argument("--cpu-workers", type=int, default=0, help="worker processes of the CPU arm (0 = all host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sparse-queries", type=int, default=32, help="queries of the BM25 get_scores sweep")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="batches in flight (2 = consecutive steps alternate between two handles / CUDA streams over the "
                         "same index, so one step's latency-bound tail -- selection, exchange, host wait -- runs under the next "
                         "step's GEMM).  0 = auto: 1 on a single GPU (power-capped GEMM: measured +0.6 %), 3 on several (r02, 8 GPUs: "
                         "3.57 / 3.65 / 3.44 ms per step with 1 / 2 / 3 in flight on one box, 3.50 / 3.445 with 1 / 2 on another)")
    ap.add_argument("--synth", default="recipe", choices=["recipe", "device", "clustered"],
                    help="recipe = the SURVEY 8d NumPy recipe, generated per 1 M-row chunk on the host (default; the only "
                         "mode with id_parity); device = same distributions drawn with torch generators in HBM (fast set-up "
                         "for profiling runs, not comparable with the oracle); clustered = device generators, 2000 tight "
                         "clusters (in-cluster cosine ~0.6), queries drawn from the clusters: robustness workload")
    ap.add_argument("--parity-queries", type=int, default=32,
                    help="queries of the batch that are also answered by the CPU oracle (oracle/sharded.py) -> id_parity")
    ap.add_argument("--synth-workers", type=int, default=0, help="host threads generating recipe chunks (0 = all cores)")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0] single-query latency block")
    ap.add_argument("--no-side-configs", action="store_true",
                    help="skip the short configs[1] and configs[4]-shard runs appended to the N=1 line")
    ap.add_argument("--sparse-c4-docs", type=int, default=20_000_000,
                    help="documents of the configs[3] BM25-only sweep appended to the N=1 line (0 = skip)")
    ap.add_argument("--query-groups", type=int, default=1,
                    help="Q query groups x (gpus/Q) row shards (dist.GridSearcher); 1 = plain row sharding")
    return ap.parse_args()


def load_traffic(kernel: str):
    """DRAM bytes of one profiled launch of `kernel` from the committed ncu --set full capture (or None)."""
    p = REPO / "profiles" / "r02_traffic.json"
    try:
        d = json.loads(p.read_text())
        rec = dict(d[kernel])
        rec["source"] = d.get("source")
        return rec
    except Exception:
        return None


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {k: float(d[k]) for k in FALLBACK_PEAKS if k in d} | {"source": "measured"}
        except Exception:
            pass
    return dict(FALLBACK_PEAKS) | {"source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference path; rank_bm25 restated, see oracle/)
# ------------------------------------------------------------------------------------------------
_CPU = {}          # state of the CPU reference arm, inherited by the forked workers


def _cpu_worker_init():
    try:                                   # one BLAS thread per worker process: the workers are the parallelism
        from threadpoolctl import threadpool_limits
        _CPU["blas_limit"] = threadpool_limits(limits=1)
    except Exception:
        pass


def _cpu_one_query(i: int) -> int:
    from oracle.hybrid import run_search_core
    c, cfg = _CPU["corpus"], _CPU["cfg"]
    toks = [f"t{int(t) + 1}" for t in _CPU["qt"][i]]
    top, _ = run_search_core(_CPU["q"][i], c.emb, _CPU["meta"], _CPU["bm25"], _CPU["skus"], toks, k=cfg["k"], rerank_k=0)
    return len(top)


def cpu_reference_sample(cfg, sample_docs: int, n_queries: int, repeats: int = 1, workers: int = 0):
    """Times the reference's single-query path (cosine_search -> BM25Okapi.get_scores -> minmax / prior / trust /
    blend / sort, exactly the order of run_search) on the first `sample_docs` docs of the same synthetic recipe.
    The reference answers one query at a time in one interpreter; to use all host cores, `workers` processes
    (forked after the index is built, so they share it) each answer a slice of the step's queries.
    Returns (seconds per repeat, worker count).  Must run before CUDA is initialised in this process (fork)."""
    import multiprocessing as mp
    import pandas as pd
    import review_recommender_b200 as rr
    from oracle.bm25_okapi import BM25Okapi

    syn = rr.synth
    c = syn.make_corpus(sample_docs, cfg["dim"], cfg["vocab"])
    corpus = syn.corpus_as_lists(c.doc_offsets, c.token_ids)
    skus = syn.skus(sample_docs)
    _CPU.update(cfg=cfg, corpus=c, q=syn.queries(n_queries, cfg["dim"]),
                qt=syn.query_terms(n_queries, cfg["terms"], c.doc_offsets, c.token_ids, cfg["vocab"]), skus=skus,
                meta=pd.DataFrame({"sku": skus, "n_reviews": c.n_reviews, "avg_stars": c.avg_stars}),
                bm25=BM25Okapi(corpus))            # index build is not timed (neither is the GPU's)
    workers = max(1, min(workers or (os.cpu_count() or 1), n_queries))
    times = []
    if workers == 1:
        for _ in range(repeats):
            t0 = time.perf_counter()
            for i in range(n_queries):
                _cpu_one_query(i)
            times.append(time.perf_counter() - t0)
        return times, 1
    with mp.get_context("fork").Pool(workers, initializer=_cpu_worker_init) as pool:
        pool.map(_cpu_one_query, range(min(n_queries, workers)))         # start-up of the workers is not timed
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_one_query, range(n_queries), chunksize=max(1, n_queries // (workers * 2)))
            times.append(time.perf_counter() - t0)
    return times, workers


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_obj(cfg, args, times, workers):
    per_rep = statistics.median(times)
    qps_sample = args.cpu_sample_queries / per_rep
    scaled = qps_sample * args.cpu_sample_docs / cfg["docs"]
    return {
        "value": scaled, "unit": UNIT, "cores": workers, "kind": "port",
        "sample": (f"first {args.cpu_sample_docs} docs of the same synthetic recipe, {args.cpu_sample_queries} queries per "
                   f"step, the reference's single-query path in run_search order (NumPy gemv + pure-Python rank_bm25 "
                   f"restatement) in {workers} forked worker processes sharing one index: {qps_sample:.3f} q/s on the "
                   f"sample; value is that figure scaled linearly to {cfg['docs']} docs (per-query cost is O(N)); "
                   f"10M docs is infeasible for the dict-per-document index"),
        "sample_queries_per_s": qps_sample,
    }


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, workers = cpu_reference_sample(cfg, args.cpu_sample_docs, args.cpu_sample_queries,
                                          repeats=args.warmup + args.steps, workers=args.cpu_workers)
    timed = times[args.warmup:]
    base = cpu_baseline_obj(cfg, args, timed, workers)
    ms = 1000.0 * statistics.mean(timed)
    line = {
        "impl": "reference", "metric": cfg.get("metric", METRIC), "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "docs": cfg["docs"], "dim": cfg["dim"], "vocab": cfg["vocab"],
                   "batch": cfg["batch"], "k": cfg["k"], "step": "bounded CPU sample, see cpu_baseline.sample"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# device-side synthetic corpus (same distributions as synth.py, generated in HBM)
# ------------------------------------------------------------------------------------------------
CLUSTERS, CLUSTER_COS = int(os.environ.get("RR_BENCH_CLUSTERS", "2000")), 0.6


def cluster_centres(D: int, dev):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(777)
    c = torch.randn((CLUSTERS, D), generator=g, device=dev, dtype=torch.float32)
    return c / torch.linalg.vector_norm(c, dim=1, keepdim=True)


def clustered_rows(n: int, D: int, seed: int, dev):
    """Rows of a catalogue made of 2000 tight clusters (robustness workload): centre + sigma * noise, normalised, with
    sigma chosen so that two rows of one cluster have cosine ~ 0.6 (1 / (1 + sigma^2 D) = 0.6)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sigma = ((1.0 / CLUSTER_COS - 1.0) / D) ** 0.5
    assign = torch.randint(0, CLUSTERS, (n,), generator=g, device=dev)
    x = cluster_centres(D, dev)[assign] + sigma * torch.randn((n, D), generator=g, device=dev, dtype=torch.float32)
    return x / torch.linalg.vector_norm(x, dim=1, keepdim=True)


def device_shard(cfg, row0: int, n: int, dev, clustered: bool = False):
    import torch
    import review_recommender_b200 as rr
    syn = rr.synth
    D, V = cfg["dim"], cfg["vocab"]
    emb = torch.empty((n, D), dtype=torch.float32, device=dev)
    cdf = torch.from_numpy(syn.zipf_cdf(V)).to(dev)
    lens_all, toks_all = [], []
    r = row0
    while r < row0 + n:
        chunk, within = divmod(r, syn.CHUNK)
        take = min(syn.CHUNK - within, row0 + n - r)
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + chunk)
        if clustered:
            x = clustered_rows(within + take, D, 1000 + chunk, dev)[within:]
        else:
            x = torch.randn((within + take, D), generator=g, device=dev, dtype=torch.float32)[within:]
            x /= torch.linalg.vector_norm(x, dim=1, keepdim=True).clamp_min(1e-12)
        emb[r - row0:r - row0 + take] = x
        del x
        g.manual_seed(3000 + chunk)
        ln = torch.exp(np.log(48.0) + 0.5 * torch.randn(within + take, generator=g, device=dev, dtype=torch.float64))
        ln = torch.clamp(torch.round(ln), 4, 256).to(torch.int64)
        g.manual_seed(4000 + chunk)
        u = torch.rand(int(ln.sum().item()), generator=g, device=dev, dtype=torch.float64)
        tok = torch.searchsorted(cdf, u).clamp_(max=V - 1).to(torch.int32)
        start = int(ln[:within].sum().item())
        lens_all.append(ln[within:].clone())
        toks_all.append(tok[start:].clone())
        del u, tok, ln
        r += take
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.cat(lens_all), 0, out=offs[1:])
    toks = torch.cat(toks_all)
    nrev, avg = syn.metadata(n, row0)
    return emb, offs, toks, nrev, avg          # the tokenised corpus stays in device memory (GpuIndexBuilder)


def recipe_shard(cfg, row0: int, n: int, dev, oracle=None, workers: int = 0):
    """This rank's rows of the SURVEY 8d recipe (rr.synth, NumPy generators, reference l2_normalize), generated chunk
    by chunk on host threads and uploaded as they arrive; every chunk is also folded into the sharded CPU oracle
    (in the worker thread) when one is given.  Returns what device_shard returns."""
    import torch
    import review_recommender_b200 as rr
    syn = rr.synth
    D, V = cfg["dim"], cfg["vocab"]
    emb = torch.empty((n, D), dtype=torch.float32, device=dev)
    lens_all, toks_all, nrev_all, avg_all = [], [], [], []

    def fold(piece):
        if oracle is not None:
            oracle.add_chunk(piece.row0, piece.emb, piece.lens, piece.token_ids, piece.n_reviews, piece.avg_stars)

    for piece in syn.chunk_stream(row0, n, D, V, workers=workers, per_chunk=fold):
        r = piece.row0 - row0
        emb[r:r + piece.emb.shape[0]].copy_(torch.from_numpy(piece.emb))
        lens_all.append(piece.lens)
        toks_all.append(piece.token_ids)
        nrev_all.append(piece.n_reviews)
        avg_all.append(piece.avg_stars)
    offs_h = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.concatenate(lens_all), out=offs_h[1:])
    offs = torch.from_numpy(offs_h).to(dev)
    toks = torch.from_numpy(np.concatenate(toks_all)).to(dev)
    return emb, offs, toks, np.concatenate(nrev_all), np.concatenate(avg_all)


def c1_block(args, fusion_cls):
    """configs[0] WHOLE on both arms, no scaling: 10 k products x 384-d, 20 k-vocab BM25, one query at a time, hybrid
    top-10 (the shape Streamlit issues, app/app_product_search.py:245-261).  GPU arm: host buffers in / out through
    rr_hybrid_search_host per query; CPU arm: the oracle port of run_search (NumPy gemv + pure-Python rank_bm25
    restatement) in one interpreter, as the reference runs it."""
    import pandas as pd
    import torch
    import review_recommender_b200 as rr
    from oracle.bm25_okapi import BM25Okapi
    from oracle.hybrid import run_search_core
    cfg = CONFIGS["c1"]
    syn = rr.synth
    n, d, v, l, k = cfg["docs"], cfg["dim"], cfg["vocab"], cfg["terms"], cfg["k"]
    c = syn.make_corpus(n, d, v)
    nq = 200
    q = syn.queries(nq, d)
    qt = syn.query_terms(nq, l, c.doc_offsets, c.token_ids, v).astype(np.int32)
    fusion = fusion_cls(k=k, rerank_k=0, w_rerank=0.0, w_best=0.0, driver="streamlit")
    ix = rr.engine.HybridIndex(c.emb, torch.from_numpy(c.doc_offsets).cuda(), torch.from_numpy(c.token_ids).cuda(), v,
                               c.n_reviews, c.avg_stars, device="cuda:0")
    nt = np.full(1, l, dtype=np.int32)
    rows = np.empty((nq, k), dtype=np.int64)
    final = np.empty((nq, k), dtype=np.float32)
    for i in range(10):
        ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows[i:i + 1], out_final=final[i:i + 1])
    lat = []
    for i in range(nq):
        t0 = time.perf_counter()
        ix.hybrid_search_host(q[i:i + 1], qt[i:i + 1], nt, fusion, out_rows=rows[i:i + 1], out_final=final[i:i + 1])
        lat.append(time.perf_counter() - t0)
    ix.close()
    lat_us = np.sort(np.asarray(lat)) * 1e6
    # CPU arm: the same queries through the oracle port, whole corpus
    skus = syn.skus(n)
    meta = pd.DataFrame({"sku": skus, "n_reviews": c.n_reviews, "avg_stars": c.avg_stars})
    bm25 = BM25Okapi(syn.corpus_as_lists(c.doc_offsets, c.token_ids))
    n_cpu = 60
    same = 0
    t0 = time.perf_counter()
    tops = []
    for i in range(n_cpu):
        toks = [f"t{int(t) + 1}" for t in qt[i]]
        top, _ = run_search_core(q[i], c.emb, meta, bm25, skus, toks, k=k, rerank_k=0)
        tops.append(top["_row"].values)
    cpu_s = time.perf_counter() - t0
    for i in range(n_cpu):
        same += int(np.sum(rows[i, :len(tops[i])] == tops[i]))
    gpu_qps = nq / float(np.sum(lat))
    cpu_qps = n_cpu / cpu_s
    return {"workload": cfg["workload"], "queries_gpu": nq, "queries_cpu": n_cpu,
            "gpu_p50_us": float(lat_us[nq // 2]), "gpu_p99_us": float(lat_us[min(nq - 1, int(nq * 0.99))]),
            "gpu_qps": gpu_qps, "cpu_qps": cpu_qps, "cpu_cores": 1, "ratio": gpu_qps / cpu_qps,
            "ids_bit_exact_rate": same / float(n_cpu * k),
            "how": "one rr_hybrid_search_host call per query (pageable host buffers in, results out, stream synchronised "
                   "inside the call), wall clock per call; CPU = oracle port of run_search in one interpreter on the same "
                   "queries, whole 10 k-doc corpus, no scaling"}


def side_config_block(name: str, n_docs: int, peaks, dev, steps: int = 10, warmup: int = 3):
    """A short device-resident run of another BASELINE config on this GPU (device-generated corpus of the same
    distributions, so no oracle comparison here -- parity at these shapes is the tests' job): ms / step, q/s and the
    tensor-roofline fraction of K2 measured like the headline's."""
    import torch
    import review_recommender_b200 as rr
    eng = rr.engine
    cfg = dict(CONFIGS[name])
    cfg["docs"] = n_docs
    D, V, B, L, K = cfg["dim"], cfg["vocab"], cfg["batch"], cfg["terms"], cfg["k"]
    emb, offs, toks, nrev, avg = device_shard(cfg, 0, n_docs, dev)
    gb = eng.GpuIndexBuilder(offs, toks, V)
    stats = gb.local_stats().finalize()
    ix = eng.HybridIndex(emb, None, None, V, nrev, avg, device=dev, stats=stats, postings=gb.finish(stats))
    del emb
    q = torch.from_numpy(rr.synth.queries(B, D)).to(dev)
    qt = torch.from_numpy(rr.synth.query_terms(B, L, offs.cpu().numpy(), toks.cpu().numpy(), V).astype(np.int32)).to(dev)
    nt = torch.full((B,), L, dtype=torch.int32, device=dev)
    del offs, toks
    wts = dict(w_dense=0.55, w_bm25=0.20, w_prior=0.20) | cfg.get("weights", {})
    fusion = eng.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0, prior_C=20.0, min_reviews=8, driver="streamlit", **wts)
    for _ in range(warmup):
        ix.hybrid_search_begin(q, qt, nt, fusion).result()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rep = 0
    for _ in range(steps):
        tok = ix.hybrid_search_begin(q, qt, nt, fusion)
        tok.result()
        rep = tok.repeated
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    eng.profile_enable(True)
    eng.profile_collect()
    for _ in range(steps):
        ix.hybrid_search_begin(q, qt, nt, fusion).result()
    torch.cuda.synchronize(dev)
    prof = eng.profile_collect()
    eng.profile_enable(False)
    kernels = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in prof.items() if v[1] > 0}
    out = {"workload": cfg["workload"] + (f" [one row shard of {n_docs} docs on this GPU]" if n_docs != CONFIGS[name]["docs"] else ""),
           "docs": n_docs, "dim": D, "batch": B, "k": K, "pool": fusion.pool, "ms_per_step": ms, "value": B / (ms / 1000.0),
           "unit": UNIT, "repeated_queries_last_step": int(rep), "dense_path": ix.dense_stats(), "kernels": kernels,
           "data": "synthetic (device generators)"}
    if "tc_filter" in kernels:
        ach = 2.0 * B * n_docs * D / (kernels["tc_filter"]["ms_per_step"] / 1000.0) / 1e12
        out["k2_tflops"] = ach
        out["k2_frac_of_sustained"] = ach / peaks["bf16_tflops_sustained"]
    ix.close()
    del ix
    torch.cuda.empty_cache()
    return out


def sparse_c4_block(args, peaks, dev):
    """configs[3]: BM25-only, 20 M documents, 200 k-term Zipf vocabulary, 16-term queries -- rr_bm25_get_scores (every
    document scored) for B in {1, 64, 256}, CUDA events around 5 calls each, against the algorithmic bytes of SURVEY 8d
    (8 B per posting of the query's terms + 4 B per document per query)."""
    import torch
    import review_recommender_b200 as rr
    n, v, l = args.sparse_c4_docs, 200_000, 16
    cfg = dict(docs=n, dim=8, vocab=v, terms=l)
    emb, offs, toks, nrev, avg = device_shard(cfg, 0, n, dev)
    gb = rr.engine.GpuIndexBuilder(offs, toks, v)
    stats = gb.local_stats().finalize()
    ix = rr.engine.HybridIndex(emb, None, None, v, device=dev, stats=stats, postings=gb.finish(stats), make_bf16=False)
    qt = rr.synth.query_terms(256, l, offs.cpu().numpy(), toks.cpu().numpy(), v).astype(np.int32)
    del offs, toks
    ib = ix.index_bytes()
    out = {"workload": f"configs[3]: BM25-only, {n} docs, 200k Zipf vocab, 16-term queries, get_scores over all documents",
           "kernel": "bm25_tile_scores_persistent_kernel", "tile_docs": ix.tile_docs, "index_bytes": ib,
           "directory_frac_of_postings": ib["directory"] / max(ib["postings"], 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "sweep": []}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for b in (1, 64, 256):
        ids = torch.from_numpy(qt[:b]).to(dev)
        nts = torch.full((b,), l, dtype=torch.int32, device=dev)
        buf = torch.empty((b, (n + 3) // 4 * 4), dtype=torch.float32, device=dev)     # reused: no allocation in the timed loop
        for _ in range(2):
            ix.bm25_get_scores(ids, nts, out=buf)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(5):
            ix.bm25_get_scores(ids, nts, out=buf)
        e1.record()
        torch.cuda.synchronize(dev)
        del buf
        ms = e0.elapsed_time(e1) / 5
        postings = int(stats.df[qt[:b]].sum())
        nbytes = 8 * postings + 4 * n * b
        out["sweep"].append({"B": b, "ms": ms, "algorithmic_bytes": nbytes, "achieved": nbytes / ms / 1e6,
                             "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"], "queries_per_s": b / ms * 1e3})
    ix.close()
    return out


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(",") for ln in open(self.f.name) if ln.strip()]
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point file descriptor 1 at stderr, so
    that anything native code prints there (NCCL's version banner at communicator creation) cannot precede the line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global OUT
    OUT = _claim_stdout()
    args = parse_args()
    cfg = dict(CONFIGS[args.config])
    args.cpu_sample_docs = min(args.cpu_sample_docs, args.docs or cfg["docs"])
    if args.docs:
        cfg["docs"] = args.docs
        cfg["workload"] += f" [debug override: docs={args.docs}]"
    if args.batch:
        cfg["batch"] = args.batch
        cfg["workload"] += f" [debug override: batch={args.batch}]"
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    # CPU baseline leg (rank 0 at N = 1 only): first, because its worker processes are forked and that has to happen
    # before this process initialises CUDA
    cpu_base = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
        times, workers = cpu_reference_sample(cfg, args.cpu_sample_docs, args.cpu_sample_queries, repeats=3,
                                              workers=args.cpu_workers)
        cpu_base = cpu_baseline_obj(cfg, args, times[1:], workers)
        _CPU.clear()

    import torch
    import torch.distributed as dist
    import review_recommender_b200 as rr
    from __graft_entry__ import build
    build()                                           # no-op when librr_b200.so is up to date
    eng = rr.engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    N, D, V, B, L, K = cfg["docs"], cfg["dim"], cfg["vocab"], cfg["batch"], cfg["terms"], cfg["k"]
    if B % world:
        raise SystemExit("batch must be a multiple of the number of GPUs")
    Q = args.query_groups
    if Q < 1 or world % Q:
        raise SystemExit("--query-groups must divide the number of GPUs")
    if args.in_flight <= 0:
        args.in_flight = 3 if (world > 1 and Q == 1) else 1
    searcher = rr.dist.GridSearcher(None, Q, lanes=args.in_flight if Q == 1 else 1) if world > 1 else None   # creates the process sub-groups (collective)
    qg, shard, R = rr.dist.GridSearcher.layout(rank, world, Q)
    row0 = N * shard // R
    n_local = N * (shard + 1) // R - row0
    t_setup = time.perf_counter()
    wts = dict(w_dense=0.55, w_bm25=0.20, w_prior=0.20) | cfg.get("weights", {})
    fusion = eng.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0, prior_C=20.0, min_reviews=8, driver="streamlit", **wts)
    # queries first (identical on every rank and for every GPU count: the recipe alone decides them), so that the
    # sampled ones can be folded into the CPU oracle while the corpus chunks stream through the host
    q_np = rr.synth.queries(B, D)
    oracle = None
    sample = np.zeros(0, dtype=np.int64)
    if args.synth == "recipe":
        qt_np = rr.synth.query_terms_global(B, L, N, V).astype(np.int32)
        if args.parity_queries > 0 and qg == 0:
            from oracle.sharded import ShardedOracle
            sample = np.unique(np.linspace(0, B - 1, min(args.parity_queries, B)).astype(np.int64))
            oracle = ShardedOracle(q_np[sample], qt_np[sample], V, fusion.pool)
        emb, offs, toks, nrev, avg = recipe_shard(cfg, row0, n_local, dev, oracle, args.synth_workers)
    else:
        emb, offs, toks, nrev, avg = device_shard(cfg, row0, n_local, dev, clustered=args.synth == "clustered")
        if args.synth == "clustered":
            q_np = clustered_rows(B, D, 2000, dev).cpu().numpy()
    # global BM25 statistics (every query group holds the whole corpus: reduce inside the row group)
    tok_counts = torch.tensor([int(offs[-1].item())], dtype=torch.int64, device=dev)
    pos0 = 0
    if R > 1:
        allc = [torch.zeros_like(tok_counts) for _ in range(R)]
        dist.all_gather(allc, tok_counts, group=searcher.row_group)
        pos0 = int(sum(int(c.item()) for c in allc[:shard]))
    # index build on the GPU: token keys -> unique (doc, term) pairs -> statistics -> impacts, forward index, postings
    t_build = time.perf_counter()
    builder = eng.GpuIndexBuilder(offs, toks, V)
    stats = builder.local_stats(token_pos0=pos0)
    local_df = stats.df.copy()
    if R > 1:
        rr.dist.all_reduce_stats(stats, group=searcher.row_group, device=dev)
    stats.finalize()
    postings = builder.finish(stats)
    torch.cuda.synchronize(dev)
    build_s = time.perf_counter() - t_build
    ix = eng.HybridIndex(emb, None, None, V, nrev, avg, device=dev, row_offset=row0, stats=stats, postings=postings)
    del emb
    if args.synth != "recipe":
        # legacy device corpus: query terms come from rank 0's documents
        if rank == 0:
            qt_np = rr.synth.query_terms(B, L, offs.cpu().numpy(), toks.cpu().numpy(), V).astype(np.int32)
        else:
            qt_np = np.zeros((B, L), dtype=np.int32)
        if world > 1:
            t = torch.from_numpy(qt_np).to(dev)
            dist.broadcast(t, 0)
            qt_np = t.cpu().numpy()
    nt_np = np.full(B, L, dtype=np.int32)
    del toks
    setup_s = time.perf_counter() - t_setup

    q_dev = torch.from_numpy(q_np).to(dev)
    qt_dev = torch.from_numpy(qt_np).to(dev)
    nt_dev = torch.from_numpy(nt_np).to(dev)
    if searcher is not None:
        searcher.ix = ix
        if searcher.inner is not None:
            searcher.inner.ix = ix

    # single GPU: every step is enqueued without a host synchronisation (hybrid_search_begin) and completed by
    # token.result(); with --in-flight 2 consecutive steps alternate between two handles / streams over the same index
    # tensors, so the tail of step i (rescoring, candidate BM25, fusion) overlaps the GEMM of step i+1
    import collections
    lanes = [(ix, torch.cuda.current_stream())]
    if args.in_flight > 1 and searcher is None:
        lanes += [(ix.view(), torch.cuda.Stream()) for _ in range(args.in_flight - 1)]
    step_no = [0]
    pending = collections.deque()
    last = [None]
    repeated = [0]

    pipelined = searcher is not None and Q == 1 and args.in_flight > 1

    def step_device():
        if pipelined:
            # row-sharded, batches in flight: enqueue this step, then complete the oldest one still pending
            pending.append(searcher.begin(q_dev, qt_dev, nt_dev, fusion, mode=args.dense_mode))
            if len(pending) >= args.in_flight:
                last[0] = pending.popleft().result()
            return last[0]
        if searcher is not None:
            last[0] = searcher.search(q_dev, qt_dev, nt_dev, fusion, mode=args.dense_mode)
            return last[0]
        h, st = lanes[step_no[0] % len(lanes)]
        step_no[0] += 1
        with torch.cuda.stream(st):
            pending.append(h.hybrid_search_begin(q_dev, qt_dev, nt_dev, fusion, mode=args.dense_mode))
        if len(pending) >= len(lanes):
            tok = pending.popleft()
            last[0] = tok.result()
            repeated[0] = tok.repeated
        return last[0]

    def drain():
        while pending:
            tok = pending.popleft()
            last[0] = tok.result()
            repeated[0] = tok.repeated
        for _, st in lanes[1:]:
            torch.cuda.current_stream().wait_stream(st)
        return last[0]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- sparse get_scores sweep (K1, HBM-bound): bytes = 8 per posting + 4 per doc ------------------
    sparse = None
    nsq = min(args.sparse_queries, B)
    if nsq > 0 and rank == 0:
        ids = qt_dev[:nsq].contiguous()
        nts = nt_dev[:nsq].contiguous()
        sp_buf = torch.empty((nsq, (n_local + 3) // 4 * 4), dtype=torch.float32, device=dev)
        for _ in range(2):
            ix.bm25_get_scores(ids, nts, out=sp_buf)
        torch.cuda.synchronize(dev)
        reps = 10
        e0.record()
        for _ in range(reps):
            ix.bm25_get_scores(ids, nts, out=sp_buf)
        e1.record()
        torch.cuda.synchronize(dev)
        sp_ms = e0.elapsed_time(e1) / reps
        del sp_buf
        postings = int(local_df[qt_np[:nsq]].sum())
        sp_bytes = 8 * postings + 4 * n_local * nsq
        sparse = {"kernel": "bm25_tile_scores_persistent_kernel (get_scores mode)", "queries": nsq, "docs": n_local,
                  "postings_per_query": postings / nsq, "ms": sp_ms, "achieved": sp_bytes / sp_ms / 1e6,
                  "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": sp_bytes / sp_ms / 1e6 / peaks["hbm_gbs"],
                  "queries_per_s": nsq / (sp_ms / 1000.0)}
    sync_all()

    # ---- device-resident timing ------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # sampled from the warm-up on: every sample is taken under the same load
    for _ in range(args.warmup):
        step_device()
    drain()
    sync_all()
    eng.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_device()
    rows, final = drain()
    e1.record()
    sync_all()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count(reset=True)
    ms_per_step = ms_total / args.steps
    # the same K steps again with every kernel launch bracketed by CUDA events on its stream (per-class
    # device time for the roofline); kept out of the loop above because ~80 event records per step cost
    # 0.1-0.2 ms of host time, which is visible when a step is only a few ms long (multi-GPU)
    eng.profile_enable(True)
    eng.profile_collect()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_device()
    drain()
    e1.record()
    sync_all()
    profiled_ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    prof = eng.profile_collect()
    eng.profile_enable(False)
    clock_info = clocks.stop() if rank == 0 else {}
    value = B / (ms_per_step / 1000.0)
    dstats = ix.dense_stats()

    # ---- end to end: pinned host inputs -> results back on the host ------------------------------------
    q_pin = torch.from_numpy(q_np).pin_memory()
    qt_pin = torch.from_numpy(qt_np).pin_memory()
    nt_pin = torch.from_numpy(nt_np).pin_memory()
    rows_pin = torch.empty((B, K), dtype=torch.int64).pin_memory()
    final_pin = torch.empty((B, K), dtype=torch.float32).pin_memory()

    nb_q, nb_t, nb_n = q_np.nbytes, qt_np.nbytes, nt_np.nbytes
    batch_pin = torch.empty(nb_q + nb_t + nb_n, dtype=torch.uint8).pin_memory()
    batch_pin[:nb_q] = torch.from_numpy(q_np).view(-1).view(torch.uint8)
    batch_pin[nb_q:nb_q + nb_t] = torch.from_numpy(qt_np).view(-1).view(torch.uint8)
    batch_pin[nb_q + nb_t:] = torch.from_numpy(nt_np).view(-1).view(torch.uint8)
    batch_dev = torch.empty_like(batch_pin, device=dev)
    bq = batch_dev[:nb_q].view(torch.float32).view(B, D)
    bqt = batch_dev[nb_q:nb_q + nb_t].view(torch.int32).view(B, L)
    bnt = batch_dev[nb_q + nb_t:].view(torch.int32)

    # multi-GPU ingest is double-buffered: batch i+1 is copied to rank 0's HBM and broadcast (on the default stream)
    # while batch i is still being searched on its lane; results of batch i-1 are read back meanwhile
    ingest = [(batch_dev, bq, bqt, bnt)]
    for _ in range(args.in_flight - 1 if pipelined else 0):
        bd = torch.empty_like(batch_pin, device=dev)
        ingest.append((bd, bd[:nb_q].view(torch.float32).view(B, D), bd[nb_q:nb_q + nb_t].view(torch.int32).view(B, L),
                       bd[nb_q + nb_t:].view(torch.int32)))
    e2e_pending = collections.deque()
    e2e_no = [0]

    def finish_e2e(tok):
        r, f = tok.result()
        if rank == 0:
            rows_pin.copy_(r, non_blocking=True)
            final_pin.copy_(f, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def step_e2e():
        if searcher is None:
            ix.hybrid_search_host(q_pin.numpy(), qt_pin.numpy(), nt_pin.numpy(), fusion, mode=args.dense_mode,
                                  out_rows=rows_pin.numpy(), out_final=final_pin.numpy())
            return
        # the query batch (vectors | term ids | term counts, one packed buffer) enters on rank 0 and is broadcast
        # ONCE; results are read back on rank 0
        bd, xq, xqt, xnt = ingest[e2e_no[0] % len(ingest)]
        e2e_no[0] += 1
        if rank == 0:
            bd.copy_(batch_pin, non_blocking=True)
        dist.broadcast(bd, 0)
        if pipelined:
            e2e_pending.append(searcher.begin(xq, xqt, xnt, fusion, mode=args.dense_mode))
            if len(e2e_pending) >= args.in_flight:
                finish_e2e(e2e_pending.popleft())
            return
        r, f = searcher.search(xq, xqt, xnt, fusion, mode=args.dense_mode)
        if rank == 0:
            rows_pin.copy_(r, non_blocking=True)
            final_pin.copy_(f, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def drain_e2e():
        while e2e_pending:
            finish_e2e(e2e_pending.popleft())

    for _ in range(min(2, args.warmup)):
        step_e2e()
    drain_e2e()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    drain_e2e()
    e1.record()
    sync_all()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = B / (e2e_ms / 1000.0)
    same = bool(np.array_equal(rows_pin.numpy(), rows.cpu().numpy())) if rank == 0 else True

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
               for k, v in prof.items() if v[1] > 0}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    roofline = None
    if dom == "tc_filter":
        flops = 2.0 * (B // Q) * n_local * D            # this GPU: its query group's slice x its row shard
        t = kernels[dom]["ms_per_step"] / 1000.0
        ach = flops / t / 1e12
        tr = load_traffic("tc_filter_pair_kernel")
        n_l = kernels[dom]["launches_per_step"]
        roofline = {"kernel": "tc_filter_pair_kernel (tcgen05 cta_group::2 bf16 GEMM + threshold filter)", "bound": "tensor",
                    "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"],
                    "traffic": tr["dram_bytes"] if tr else None,
                    "traffic_note": (f"DRAM read+write bytes of ONE profiled launch ({tr['launch']}; its bf16 corpus segment is "
                                     f"{tr['corpus_bytes_of_segment']:.3e} B, read once) -- {tr['source']}") if tr else None,
                    "peak_source": f"{peaks['source']} bf16_tflops_sustained (cuBLAS back to back); burst "
                                   f"{peaks['bf16_tflops']:.1f} -> frac {ach / peaks['bf16_tflops']:.3f}",
                    "algorithmic_flops_per_step": flops, "launches_per_step": n_l,
                    "algorithmic_flops_per_launch_avg": flops / max(n_l, 1.0),
                    "avg_launch_ms": kernels[dom]["ms_per_step"] / max(n_l, 1.0)}
    elif dom == "dense_gemv":
        groups = (B // Q + 7) // 8
        nbytes = 4.0 * D * n_local * groups
        t = kernels[dom]["ms_per_step"] / 1000.0
        ach = nbytes / t / 1e9
        roofline = {"kernel": "dense_scores_f32_kernel (exact fp32 multi-query GEMV)", "bound": "hbm",
                    "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "traffic": None, "peak_source": f"{peaks['source']} hbm_gbs",
                    "launches_per_step": kernels[dom]["launches_per_step"]}
    elif dom is not None:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": None, "traffic": None}

    # ---- parity at the benchmark configuration: sampled queries against the CPU oracle ---------------------
    # every rank folded its own chunks into a partial oracle; rank 0 merges them (corpus statistics + the union of
    # the per-chunk dense top-pools) and answers the sampled queries the way run_search would
    id_parity = None
    if args.synth == "recipe" and args.parity_queries > 0:
        parts = [oracle.partial() if oracle is not None else []]
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(parts[0], gathered, dst=0)
            parts = gathered
        if rank == 0:
            from oracle.sharded import compare_with_oracle
            t_or = time.perf_counter()
            oracle.merge(parts).finalize()
            id_parity = compare_with_oracle(oracle, sample, rows.cpu().numpy(), final.cpu().numpy(), K, "streamlit",
                                            rerank_k=0, w_rerank=0.0, w_best=0.0, prior_C=20.0, min_reviews=8, **wts)
            id_parity["oracle_s"] = time.perf_counter() - t_or
            id_parity["how"] = ("oracle/sharded.py: run_search_core (pinned restatement of app/app_product_search.py:253-312) "
                                "over the union of per-chunk dense top-(pool+32) rows with corpus-global BM25 statistics; "
                                "same recipe arrays as the GPU index")
    digest = None
    if rank == 0:
        import hashlib
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(rows.cpu().numpy()).tobytes())
        h.update(np.ascontiguousarray(final.cpu().numpy()).tobytes())
        digest = h.hexdigest()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    c1 = None
    if world == 1 and not args.no_c1 and args.config == "c3":
        ix.close()
        del ix
        torch.cuda.empty_cache()
        c1 = c1_block(args, eng.Fusion)
    sparse_c4 = None
    if world == 1 and args.sparse_c4_docs > 0 and args.config == "c3":
        try:
            ix.close()
        except Exception:
            pass
        torch.cuda.empty_cache()
        sparse_c4 = sparse_c4_block(args, peaks, dev)
    side = None
    if world == 1 and not args.no_side_configs and args.config == "c3":
        try:
            ix.close()
        except Exception:
            pass
        torch.cuda.empty_cache()
        side = {"configs1": side_config_block("c2", CONFIGS["c2"]["docs"], peaks, dev, steps=20, warmup=3),
                "configs4_one_of_8_shards": side_config_block("c5", CONFIGS["c5"]["docs"] // 8, peaks, dev, steps=5, warmup=3)}


    line = {
        "metric": cfg.get("metric", METRIC), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic" + {"recipe": " (SURVEY 8d NumPy recipe)", "device": " (device generators)",
                                "clustered": " (device generators, 2000 clusters with in-cluster cosine 0.6: robustness workload)"}[args.synth],
        "dtype_note": "bf16 tensor-core shortlist (fp32 accumulate), exact f32 rescoring, f32/f64 fusion as the reference",
        "config": {"workload": cfg["workload"], "docs": N, "docs_per_gpu": n_local, "dim": D, "vocab": V, "batch": B,
                   "query_terms": L, "k": K, "pool": fusion.pool, "weights": wts, "parallelism": f"row-sharded x{R}" + (f", query groups x{Q}" if Q > 1 else ""),
                   "l2": "inputs larger than L2 (bf16 corpus shard read every step)",
                   "dense_path": dstats, "repeated_queries_last_step": int(repeated[0] if searcher is None else searcher.inner.last_repeated if searcher.inner is not None else 0),
                   "setup_s": setup_s, "index_build_s": build_s, "batches_in_flight": args.in_flight if searcher is not None else len(lanes)},
        "roofline": roofline, "kernels": kernels, "profiled_ms_per_step": profiled_ms_per_step, "sparse": sparse,
        "cpu_baseline": cpu_base,
        "clocks": {"sm_mhz": clock_info.get("sm_mhz"), "sm_max_mhz": clock_info.get("sm_max_mhz"),
                   "reasons": clock_info.get("reasons", []), "samples": clock_info.get("samples", 0)},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(q_np.nbytes + qt_np.nbytes + nt_np.nbytes),
                "d2h_bytes_per_step": int(B * K * 12), "results_equal_device_path": same},
        "gpu_launches": int(launches),
        "id_parity": id_parity, "result_digest": digest, "c1": c1, "sparse_c4": sparse_c4, "other_configs": side,
    }
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
