"""Multi-rank plumbing of the row-sharded search, on CPU with gloo (world size 2): BM25 statistics
all-reduce, tuple pack -> all-to-all -> unpack, and the property the merge relies on (the global
top-pool by dense score is contained in the union of the per-shard top-pools)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["RR_REPO"])
    import review_recommender_b200 as rr
    from oracle.primitives import cosine_topk_canonical

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N, D, V, B, pool = 3000, 32, 200, 8, 40
    syn = rr.synth
    row0 = N * rank // world
    n_local = N * (rank + 1) // world - row0
    full_offs, full_toks = syn.corpus_tokens(N, V)
    offs, toks = syn.corpus_tokens(n_local, V, row0)

    # ---- global BM25 statistics from per-shard statistics --------------------------------------
    st = rr.engine.BM25Stats.local(offs, toks, V, token_pos0=int(full_offs[row0]))
    rr.dist.all_reduce_stats(st)
    st.finalize()
    whole = rr.engine.BM25Stats.local(full_offs, full_toks, V).finalize()
    assert np.array_equal(st.idf, whole.idf) and st.avgdl == whole.avgdl and st.n_docs == N

    # ---- local top-pool tuples (NumPy stands in for the GPU kernels), exchange, merge -------------
    emb = syn.embeddings(n_local, D, row0)
    q = syn.queries(B, D)
    grow = np.empty((B, pool), dtype=np.int64); dense = np.empty((B, pool), dtype=np.float32)
    for b in range(B):
        idx, sims = cosine_topk_canonical(q[b], emb, pool)
        grow[b], dense[b] = idx + row0, sims
    n = (grow % 17).astype(np.float64); avg = (grow % 5).astype(np.float64); bm25 = (grow % 3).astype(np.float32)
    t = torch.from_numpy
    send = rr.dist.pack_tuples(world, t(grow), t(n), t(avg), t(dense), t(bm25))
    recv = rr.dist.exchange(send)
    Bg = B // world
    views, stride = rr.dist.field_views(recv, Bg, pool)
    assert stride == Bg * pool * rr.dist.TUPLE_BYTES
    full_emb = syn.embeddings(N, D)
    for b in range(Bg):
        gq = rank * Bg + b                                        # the global query this rank owns
        rows, scores = [], []
        for s in range(world):
            g_, n_, a_, d_, b_ = rr.dist.unpack_shard(recv, s, Bg, pool)
            assert np.array_equal(n_[b].numpy(), (g_[b].numpy() % 17).astype(np.float64))
            assert np.array_equal(b_[b].numpy(), (g_[b].numpy() % 3).astype(np.float32))
            lo, hi = N * s // world, N * (s + 1) // world
            assert np.all((g_[b].numpy() >= lo) & (g_[b].numpy() < hi))
            rows.append(g_[b].numpy()); scores.append(d_[b].numpy())
        rows, scores = np.concatenate(rows), np.concatenate(scores)
        order = np.lexsort((rows, -scores.astype(np.float64)))[:pool]
        ref_idx, ref_sims = cosine_topk_canonical(q[gq], full_emb, pool)
        assert np.array_equal(rows[order], ref_idx), (rank, b)
        assert np.allclose(scores[order], ref_sims, atol=1e-6)
    # ---- 40-byte tuples: gate factor and raw best-review similarity ride along (run_search :285-310) ------------
    gate = ((grow % 7) / 7.0).astype(np.float32); best = ((grow % 11) / 11.0).astype(np.float32)
    send2 = rr.dist.pack_tuples(world, t(grow), t(n), t(avg), t(dense), t(bm25), t(gate), t(best))
    assert send2.shape == (world, Bg * pool * rr.dist.TUPLE_BYTES_EXTRAS)
    recv2 = rr.dist.exchange(send2)
    views2, stride2 = rr.dist.field_views(recv2, Bg, pool, with_extras=True)
    assert stride2 == Bg * pool * rr.dist.TUPLE_BYTES_EXTRAS
    for s_ in range(world):
        g_ = views2["grow"][s_ * stride2: s_ * stride2 + Bg * pool * 8].view(torch.int64).numpy()
        ga = views2["gate"][s_ * stride2: s_ * stride2 + Bg * pool * 4].view(torch.float32).numpy()
        be = views2["best"][s_ * stride2: s_ * stride2 + Bg * pool * 4].view(torch.float32).numpy()
        assert np.array_equal(ga, ((g_ % 7) / 7.0).astype(np.float32)) and np.array_equal(be, ((g_ % 11) / 11.0).astype(np.float32))
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok", flush=True)
''')


def test_two_rank_exchange_and_stats(tmp_path):
    from __graft_entry__ import build
    build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RR_REPO=str(REPO), OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:                     # a free port, so reruns never collide
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


GRID_WORKER = textwrap.dedent('''
    import os, sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["RR_REPO"])
    import review_recommender_b200 as rr

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    class FakeIndex:
        """Stands in for the CUDA index: the 'answer' of query i is (i*10 + j, i + j/100)."""
        def hybrid_search(self, q, term_ids, n_terms, fusion, mode=0):
            ids = q[:, 0].to(torch.int64)
            k = fusion.k
            rows = ids[:, None] * 10 + torch.arange(k)[None, :]
            final = ids[:, None].to(torch.float32) + torch.arange(k)[None, :].to(torch.float32) / 100
            return rows, final

    class F:
        k, pool = 3, 5

    B = 8
    q = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, 4)
    for Q in (1, 2):
        g, s, R = rr.dist.GridSearcher.layout(rank, world, Q)
        assert (g, s, R) == (divmod(rank, world // Q) + (world // Q,))
        if Q == 1:
            continue            # R = 2 needs the CUDA kernels (tests/test_gpu_dist.py); here: groups only
        grid = rr.dist.GridSearcher(FakeIndex(), Q)
        assert grid.inner is None and grid.col_group is not None
        rows, final = grid.search(q, None, None, F())
        want_rows = torch.arange(B)[:, None] * 10 + torch.arange(3)[None, :]
        assert torch.equal(rows, want_rows), rows
        assert torch.allclose(final, torch.arange(B)[:, None].float() + torch.arange(3)[None, :].float() / 100)
        # odd (B/Q)*k: the gathered byte rows must stay 8-byte aligned (ADVICE r01: stride not divisible by 8)
        class F5:
            k, pool = 5, 5
        q2 = torch.arange(2, dtype=torch.float32)[:, None].repeat(1, 4)
        rows, final = grid.search(q2, None, None, F5())
        assert torch.equal(rows, torch.arange(2)[:, None] * 10 + torch.arange(5)[None, :]), rows
        assert torch.allclose(final, torch.arange(2)[:, None].float() + torch.arange(5)[None, :].float() / 100)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} grid ok", flush=True)
''')


def test_query_group_grid_on_two_ranks(tmp_path):
    """GridSearcher plumbing (sub-groups, batch slicing, the column all-gather) with two gloo ranks and a stub index."""
    script = tmp_path / "grid_worker.py"
    script.write_text(GRID_WORKER)
    env = dict(os.environ, RR_REPO=str(REPO), OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "rank 0 grid ok" in r.stdout and "rank 1 grid ok" in r.stdout
