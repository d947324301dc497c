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

    # ---- gate and best-review columns ride in the tuples (run_search :285-310) ----------------------------------
    from tests.golden_worlds import make_gate_texts, GATE_QUERIES
    texts = make_gate_texts(N, full.doc_offsets, full.token_ids)
    skus = syn.skus(N)
    rngr = np.random.default_rng(21)
    M = 3 * N // 10
    rev_prod = rngr.integers(0, N, size=M)
    rev_emb = rngr.standard_normal((M, D)).astype(np.float32)
    rev_skus = [skus[i] for i in rev_prod]
    Bx = 16 * max(world, 1)
    qx = q[:Bx]
    qtx, ntx = qt[:Bx], nt[:Bx]
    groups = [rr.drop_in.build_gate_groups(GATE_QUERIES[i % len(GATE_QUERIES)]) for i in range(Bx)]
    fx = rr.engine.Fusion(k=20, rerank_k=0, w_dense=0.45, w_bm25=0.15, w_rerank=0.0, w_prior=0.10, w_best=0.30, best_is_raw=True)
    g_loc = rr.engine.GateIndex(texts[row0:row1], rr.drop_in.GATE_FIXED_GROUPS, device=dev)
    r_loc = rr.engine.ReviewIndex(rev_emb, rev_skus, skus[row0:row1], device=dev)
    sx = rr.dist.ShardedSearcher(ix, extras=rr.dist.make_extras(g_loc, r_loc, groups, 0.5))
    rows_x, final_x = sx.search(torch.from_numpy(qx).to(dev), torch.from_numpy(qtx).to(dev), torch.from_numpy(ntx).to(dev),
                                fx, mode=rr._lib.RR_DENSE_EXACT)
    if rank == 0:
        whole = rr.engine.HybridIndex(full.emb, full.doc_offsets, full.token_ids, V, full.n_reviews, full.avg_stars, device=dev)
        g_all = rr.engine.GateIndex(texts, rr.drop_in.GATE_FIXED_GROUPS, device=dev)
        r_all = rr.engine.ReviewIndex(rev_emb, rev_skus, skus, device=dev)
        cand, dense, cnt = whole.dense_topk(qx, fx.pool, rr._lib.RR_DENSE_EXACT)
        bm25, n_, avg_, grow_ = whole.candidate_tuples(qtx, ntx, cand)
        gate = g_all.factors(groups, cand, 0.5)
        best, _ = r_all.best(qx, cand, max_rows=None, as_numpy=False)
        w_rows, w_final, _, _ = whole.fuse(fx, dense, bm25, n_, avg_, grow_, count=cnt, best=best, gate=gate)
        assert torch.equal(rows_x, w_rows), int((rows_x != w_rows).sum())
        assert torch.equal(final_x, w_final)
        assert float(gate.min()) < 1.0 and float(best.max()) > 0.0, "the extras must actually matter in this case"
        whole.close()
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
