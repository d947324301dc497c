"""Row-sharded search on >= 2 GPUs (NCCL): results must equal the single-GPU search bit for bit
(same global rows, same fused scores).  Skipped on a single-GPU box."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["RR_REPO"])
    import review_recommender_b200 as rr

    local = int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    N, D, V, B, L, K = 200_000, 384, 5000, 256, 4, 100
    syn = rr.synth
    full = syn.make_corpus(N, D, V)
    q = syn.queries(B, D)
    # 600 rows within 1e-4 of row 17 and a query equal to it: the tensor path cannot certify that query, so the
    # shard marks its tuples (global row -2) and the query must take the synchronous second round
    rng = np.random.default_rng(3)
    base = full.emb[17].copy()
    for r in range(1000, 1600):
        v = base + 1e-4 * rng.standard_normal(D).astype(np.float32)
        full.emb[r] = v / np.linalg.norm(v)
    q[0] = base
    qt = syn.query_terms(B, L, full.doc_offsets, full.token_ids, V).astype(np.int32)
    nt = np.full(B, L, dtype=np.int32)
    fusion = rr.engine.Fusion(k=K, rerank_k=0, w_rerank=0.0, w_best=0.0)

    row0 = N * rank // world
    row1 = N * (rank + 1) // world
    offs = full.doc_offsets[row0:row1 + 1] - full.doc_offsets[row0]
    toks = full.token_ids[full.doc_offsets[row0]:full.doc_offsets[row1]]
    st = rr.engine.BM25Stats.local(offs, toks, V, token_pos0=int(full.doc_offsets[row0]))
    rr.dist.all_reduce_stats(st, device=dev)
    st.finalize()
    ix = rr.engine.HybridIndex(full.emb[row0:row1], offs, toks, V, full.n_reviews[row0:row1], full.avg_stars[row0:row1],
                               device=dev, row_offset=row0, stats=st)
    whole_ref = None
    for mode, r1 in ((rr._lib.RR_DENSE_EXACT, None), (rr._lib.RR_DENSE_TENSOR, None), (rr._lib.RR_DENSE_TENSOR, 16)):
        # r1=16: far too few tuples per shard in round 1 -> many queries must take the exact second round
        searcher = rr.dist.ShardedSearcher(ix, round1_pool=r1)
        rows, final = searcher.search(torch.from_numpy(q).to(dev), torch.from_numpy(qt).to(dev),
                                      torch.from_numpy(nt).to(dev), fusion, mode=mode)
        if r1 == 16:
            assert searcher.last_repeated > 1
        if mode == rr._lib.RR_DENSE_TENSOR:
            assert searcher.last_repeated >= 1, "the near-duplicate query must be repeated"
        if mode == rr._lib.RR_DENSE_EXACT and world == 1:
            assert searcher.last_repeated == 0
        print("rank", rank, "mode", mode, "round1", r1 or rr.dist.local_pool(fusion.pool, world), "repeated", searcher.last_repeated)
        if rank == 0:
            whole = rr.engine.HybridIndex(full.emb, full.doc_offsets, full.token_ids, V, full.n_reviews, full.avg_stars,
                                          device=dev)
            r1, f1 = whole.hybrid_search(q, qt, nt, fusion, mode=rr._lib.RR_DENSE_EXACT)
            assert torch.equal(rows, r1), (mode, int((rows != r1).sum()))
            assert torch.equal(final, f1), mode
            whole.close()
    # ---- batches in flight: two lanes, results identical to the one-at-a-time search ----------------------------
    qd, qtd, ntd = torch.from_numpy(q).to(dev), torch.from_numpy(qt).to(dev), torch.from_numpy(nt).to(dev)
    one = rr.dist.ShardedSearcher(ix)
    ref_rows, ref_final = one.search(qd, qtd, ntd, fusion, mode=rr._lib.RR_DENSE_TENSOR)
    two = rr.dist.ShardedSearcher(ix, lanes=2)
    perm = torch.arange(B - 1, -1, -1, device=dev)
    batches = [(qd, qtd, ntd), (qd[perm].contiguous(), qtd[perm].contiguous(), ntd[perm].contiguous())] * 3
    tokens, outs = [], []
    for bq, bt, bn in batches:
        tokens.append(two.begin(bq, bt, bn, fusion, mode=rr._lib.RR_DENSE_TENSOR))
        if len(tokens) == 2:
            outs.append(tokens.pop(0).result())
    outs += [t.result() for t in tokens]
    torch.cuda.synchronize(dev)
    for i, (r_, f_) in enumerate(outs):
        want_r, want_f = (ref_rows, ref_final) if i % 2 == 0 else (ref_rows[perm], ref_final[perm])
        assert torch.equal(r_, want_r) and torch.equal(f_, want_f), ("lanes", i)
    print(f"rank {rank} two lanes ok", flush=True)

    # ---- gate and best-review columns ride in the tuples (run_search :285-310): a sharded run_search with snippets
    # and gates on must return what the single-GPU SearchEngine.run_search returns ---------------------------------
    import pandas as pd
    from tests.golden_worlds import make_gate_texts, GATE_QUERIES
    N2, D2, V2 = 20_000, 64, 1500
    w2 = syn.make_corpus(N2, D2, V2)
    texts = make_gate_texts(N2, w2.doc_offsets, w2.token_ids)
    skus2 = syn.skus(N2)
    rngr = np.random.default_rng(21)
    M = 3 * N2
    rev_prod = rngr.integers(0, N2, size=M)
    rev_emb = rngr.standard_normal((M, D2)).astype(np.float32)
    rev_skus = [skus2[i] for i in rev_prod]
    Bx = 8 * world
    qx = syn.queries(Bx, D2)
    qt2 = syn.query_terms(Bx, 3, w2.doc_offsets, w2.token_ids, V2)
    query_strs = [" ".join(f"t{int(t) + 1}" for t in qt2[i]) + " " + GATE_QUERIES[i % len(GATE_QUERIES)] for i in range(Bx)]
    toks = [rr.drop_in.tokenize_query(s) for s in query_strs]
    ids = [[(int(t[1:]) - 1 if t[0] == "t" and t[1:].isdigit() and int(t[1:]) <= V2 else -1) for t in tk] for tk in toks]
    tid, ntm = rr.engine.HybridIndex.pack_terms(ids)
    groups = [rr.drop_in.build_gate_groups(s) for s in query_strs]
    kw = dict(k=20, rerank_k=0, w_dense=0.45, w_bm25=0.15, w_rerank=0.0, w_prior=0.10, w_best=0.30, prior_C=20.0, min_reviews=8)
    fx = rr.engine.Fusion(driver="streamlit", best_is_raw=True, **kw)
    a0, a1 = N2 * rank // world, N2 * (rank + 1) // world
    o2 = w2.doc_offsets[a0:a1 + 1] - w2.doc_offsets[a0]
    t2 = w2.token_ids[w2.doc_offsets[a0]:w2.doc_offsets[a1]]
    st2 = rr.engine.BM25Stats.local(o2, t2, V2, token_pos0=int(w2.doc_offsets[a0]))
    rr.dist.all_reduce_stats(st2, device=dev)
    st2.finalize()
    ixx = rr.engine.HybridIndex(w2.emb[a0:a1], o2, t2, V2, w2.n_reviews[a0:a1], w2.avg_stars[a0:a1], device=dev,
                                row_offset=a0, stats=st2)
    g_loc = rr.engine.GateIndex(texts[a0:a1], rr.drop_in.GATE_FIXED_GROUPS, device=dev)
    r_loc = rr.engine.ReviewIndex(rev_emb, rev_skus, skus2[a0:a1], device=dev)
    for r1x in (None, 8):                      # 8: too few tuples in round 1 -> the second round must carry the extras too
        sx = rr.dist.ShardedSearcher(ixx, round1_pool=r1x, extras=rr.dist.make_extras(g_loc, r_loc, groups, 0.5))
        rows_x, final_x = sx.search(torch.from_numpy(qx).to(dev), torch.from_numpy(tid).to(dev), torch.from_numpy(ntm).to(dev),
                                    fx, mode=rr._lib.RR_DENSE_EXACT)
        if r1x == 8:
            assert sx.last_repeated > 0
        if rank == 0:
            meta = pd.DataFrame({"sku": skus2, "n_reviews": w2.n_reviews.astype(np.float64), "avg_stars": w2.avg_stars,
                                 "agg_text": texts})
            reviews = pd.DataFrame({"sku": rev_skus, "text": [f"review {j}" for j in range(M)],
                                    "stars": (rev_prod % 5 + 1).astype(np.float64), "embedding": list(rev_emb)})
            table = dict(zip(query_strs, qx))
            se = rr.drop_in.SearchEngine(meta, w2.emb, syn.corpus_as_lists(w2.doc_offsets, w2.token_ids), skus2,
                                         encode=lambda s_: table[s_], reviews=reviews, device=str(dev))
            res = se.run_search_batch(query_strs, use_snips=True, max_scan=10**9, gate_penalty=0.5, **kw)
            rows_h, final_h = rows_x.cpu().numpy(), final_x.cpu().numpy()
            n_gated = n_best = 0
            for i, (top, snips, dbg) in enumerate(res):
                assert top["sku"].tolist() == [skus2[r] for r in rows_h[i]], ("extras", r1x, i)
                assert np.array_equal(top["_final"].values.astype(np.float32), final_h[i]), ("extras final", r1x, i)
                n_gated += int((top["_gate"].values < 1.0).sum())
                n_best += int((top["_best"].values > 0.0).sum())
            assert n_gated > 0 and n_best > 0, "the extras must actually matter in this case"
            se.ix.close()
    ixx.close()
    print(f"rank {rank} extras ok", flush=True)

    # ---- row shards x query groups: same answers for every layout ------------------------------------------
    for Q in [x for x in (2, 4) if world % x == 0]:
        grid = rr.dist.GridSearcher(None, Q)
        g, s, R = rr.dist.GridSearcher.layout(rank, world, Q)
        r0, r1 = N * s // R, N * (s + 1) // R
        o2 = full.doc_offsets[r0:r1 + 1] - full.doc_offsets[r0]
        t2 = full.token_ids[full.doc_offsets[r0]:full.doc_offsets[r1]]
        st2 = rr.engine.BM25Stats.local(o2, t2, V, token_pos0=int(full.doc_offsets[r0]))
        if R > 1:
            rr.dist.all_reduce_stats(st2, group=grid.row_group, device=dev)
        st2.finalize()
        ix2 = rr.engine.HybridIndex(full.emb[r0:r1], o2, t2, V, full.n_reviews[r0:r1], full.avg_stars[r0:r1],
                                    device=dev, row_offset=r0, stats=st2)
        grid.ix = ix2
        if grid.inner is not None:
            grid.inner.ix = ix2
        rows, final = grid.search(qd, qtd, ntd, fusion, mode=rr._lib.RR_DENSE_TENSOR)
        assert torch.equal(rows, ref_rows) and torch.equal(final, ref_final), ("grid", Q)
        print(f"rank {rank} grid Q={Q} R={R} ok", flush=True)
        ix2.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok", flush=True)
''')


def test_sharded_equals_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _run_world(tmp_path, 2 if n < 4 else 4)


def test_world_of_one_runs_the_same_protocol(tmp_path):
    """One rank over NCCL: rr_shard_tuples (no host synchronisation), the poisoned-tuple path for uncertified
    queries and the single all-gather buffer, on a single-GPU box."""
    _run_world(tmp_path, 1)


def _run_world(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, RR_REPO=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    for k in range(world):
        assert f"rank {k} ok" in r.stdout and f"rank {k} two lanes ok" in r.stdout and f"rank {k} extras ok" in r.stdout
