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


PROTOCOL_WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["RR_REPO"])
    import review_recommender_b200 as rr

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N, D, B, POOL, K = 4000, 24, 8, 40, 7
    syn = rr.synth
    emb_all = syn.embeddings(N, D)
    # rows 0..49 all look like query 3: the global pool of that query sits on shard 0 -> round 1 (m < pool) cannot prove it
    q = syn.queries(B, D)
    emb_all[:50] = q[3] + 1e-3 * np.random.default_rng(1).standard_normal((50, D)).astype(np.float32)

    class Fusion:
        k, pool = K, POOL

    def meta(grow, qsum):
        bm25 = (((grow * 7 + qsum) % 13) / 13.0).to(torch.float32)
        return bm25, (grow % 17).to(torch.float64), (grow % 5 + 1).to(torch.float64)

    class StubIndex:
        """NumPy / torch-CPU stand-in for engine.HybridIndex: the same call surface ShardedSearcher uses."""
        def __init__(self, emb, row0):
            self.emb, self.row0, self.device = torch.from_numpy(np.ascontiguousarray(emb)), row0, torch.device("cpu")
        def dense_topk(self, q, m, mode=0, want_uncertified=False):
            sims = q @ self.emb.T
            n = sims.shape[1]
            key = np.lexsort((np.broadcast_to(np.arange(n), sims.shape), -sims.numpy().astype(np.float64)), axis=1)[:, :m]
            idx = torch.from_numpy(np.ascontiguousarray(key)).to(torch.int64)
            out = (idx, torch.gather(sims, 1, idx).contiguous(), torch.full((q.shape[0],), min(m, n), dtype=torch.int32))
            return out + (torch.zeros(q.shape[0], dtype=torch.int32),) if want_uncertified else out
        def candidate_tuples(self, term_ids, n_terms, cand):
            grow = cand + self.row0
            bm25, n, avg = meta(grow, term_ids.sum(dim=1, keepdim=True))
            return bm25, n, avg, grow
        def shard_tuples(self, q, term_ids, n_terms, m, G, mode=0):
            cand, dense, _ = self.dense_topk(q, m)
            bm25, n, avg, grow = self.candidate_tuples(term_ids, n_terms, cand)
            return rr.dist.pack_tuples(G, grow, n, avg, dense, bm25)
        def fuse_sharded(self, fusion, G, m, stride, Bg, dense, bm25, n, avg, grow, out=None, gate=None, best=None):
            def blk(v, s, dt, esz):
                return v[s * stride: s * stride + Bg * m * esz].view(dt).view(Bg, m)
            rows, final, flags = out
            for b in range(Bg):
                tup, incomplete, per_shard = [], 0, []
                for s in range(G):
                    g = blk(grow, s, torch.int64, 8)[b]; d = blk(dense, s, torch.float32, 4)[b]
                    bm = blk(bm25, s, torch.float32, 4)[b]; nn = blk(n, s, torch.float64, 8)[b]
                    ga = blk(gate, s, torch.float32, 4)[b] if gate is not None else torch.ones(m)
                    be = blk(best, s, torch.float32, 4)[b] if best is not None else torch.zeros(m)
                    if int(g[0]) == -2:
                        incomplete = 1
                    ok = g >= 0
                    per_shard.append((int(ok.sum()), float(d[ok].min()) if ok.any() else np.inf))
                    tup += [(float(d[j]), int(g[j]), float(bm[j]), float(nn[j]), float(ga[j]), float(be[j])) for j in range(m) if ok[j]]
                tup.sort(key=lambda t: (-t[0], t[1]))
                poolt = tup[:fusion.pool]
                cut = poolt[-1][0] if len(tup) >= fusion.pool else -np.inf
                if m < fusion.pool and any(c == m and w >= cut for c, w in per_shard):
                    incomplete = 1
                fin = [((np.float32(t[0]) + np.float32(0.1) * np.float32(t[2]) + np.float32(0.01 * t[3]) + np.float32(t[5])) * np.float32(t[4]), i)
                       for i, t in enumerate(poolt)]
                fin.sort(key=lambda x: (-x[0], x[1]))
                for i in range(fusion.k):
                    rows[b, i] = poolt[fin[i][1]][1] if i < len(fin) else -1
                    final[b, i] = float(fin[i][0]) if i < len(fin) else float("nan")
                flags[b] = incomplete
            return rows, final, flags

    r0, r1 = N * rank // world, N * (rank + 1) // world
    shard = StubIndex(emb_all[r0:r1], r0)
    whole = StubIndex(emb_all, 0)
    qd = torch.from_numpy(q)
    terms = torch.arange(B * 3, dtype=torch.int32).view(B, 3) % 11
    nts = torch.full((B,), 3, dtype=torch.int32)

    def make_extras(ix):
        def f(cand, qq, which=None):
            grow = cand + ix.row0
            pos = torch.arange(qq.shape[0]) if which is None else which
            gate = torch.where((grow + pos[:, None]) % 3 == 0, torch.tensor(0.5), torch.tensor(1.0)).to(torch.float32)
            best = ((grow % 7).to(torch.float32) / 7.0)
            return gate, best
        return f

    for with_extras in (False, True):
        # reference: one shard holding everything, m = pool (always exact)
        cand, dense, _ = whole.dense_topk(qd, POOL)
        bm25, n_, avg_, grow_ = whole.candidate_tuples(terms, nts, cand)
        g_, b_ = make_extras(whole)(cand, qd) if with_extras else (None, None)
        send = rr.dist.pack_tuples(1, grow_, n_, avg_, dense, bm25, g_, b_)
        views, stride = rr.dist.field_views(send, B, POOL, with_extras)
        want_rows = torch.empty((B, K), dtype=torch.int64); want_final = torch.empty((B, K), dtype=torch.float32)
        whole.fuse_sharded(Fusion, 1, POOL, stride, B, views["dense"], views["bm25"], views["n"], views["avg"], views["grow"],
                           out=(want_rows, want_final, torch.zeros(B, dtype=torch.int32)), gate=views.get("gate"), best=views.get("best"))
        for r1pool in (None, 24, 6):
            s_ = rr.dist.ShardedSearcher(shard, round1_pool=r1pool, lanes=2, extras=make_extras(shard) if with_extras else None)
            t1 = s_.begin(qd, terms, nts, Fusion)
            t2 = s_.begin(qd.flip(0).contiguous(), terms.flip(0).contiguous(), nts, Fusion)
            rows, final = t1.result()
            assert torch.equal(rows, want_rows), (with_extras, r1pool, rows, want_rows)
            assert torch.equal(final, want_final), (with_extras, r1pool)
            if r1pool is None:
                assert t1.repeated == 0, "m = pool here: round 1 is always exact"
            else:
                assert t1.repeated >= 1, "query 3's pool sits on one shard: a round 1 with m < pool cannot prove it"
            if r1pool == 6:
                assert t1.repeated > 1
            rows2, final2 = t2.result()
            if not with_extras:            # (the stub's gate depends on the batch position, so only compare without)
                assert torch.equal(rows2, want_rows.flip(0)) and torch.equal(final2, want_final.flip(0))
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} protocol ok", flush=True)
''')


def test_sharded_protocol_with_a_stub_index_on_two_ranks(tmp_path):
    """ShardedSearcher end to end on CPU (gloo, two ranks, NumPy stub index): round-1 pool m < pool, the proof that fails
    for a query whose pool sits on one shard, the collective second round (with `which`), 40-byte tuples with the gate /
    best-review columns, begin() / result() tokens -- results equal the one-shard answer."""
    script = tmp_path / "protocol_worker.py"
    script.write_text(PROTOCOL_WORKER)
    env = dict(os.environ, RR_REPO=str(REPO), OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "rank 0 protocol ok" in r.stdout and "rank 1 protocol ok" in r.stdout
