"""Row-sharded hybrid search: one process per GPU, `torch.distributed` (NCCL over NVLink) for the
single exchange step the path has.

The reference has no distributed code; the sharding follows BASELINE.json's
north_star: the corpus is cut row-wise, every rank holds rows [row_offset, row_offset+n) with the
postings of those rows and GLOBAL BM25 statistics, and per batch

  1. every rank computes, for ALL B queries, its local exact top-m by dense similarity (m = pool, or
     fewer with the two-round proof of `ShardedSearcher`) and the candidate tuples
     (dense, bm25, n_reviews, avg_stars, global row) -- K2/K3 + K1 candidates;
  2. ONE all-to-all ships, to rank r, the tuples of query slice r from every shard
     (B*pool*32 bytes leave each rank; an all-gather would move G times more);
  3. rank r merges its G*pool tuples per query by (dense desc, global row asc), keeps `pool`
     (exact: the global top-pool is a subset of the union of local top-pools) and runs the fusion
     kernel (K4) on the merged pool -- min-max and nanmean are pool-global, so they must run after
     the merge;
  4. one small all-gather returns the [B, k] results to every rank.

Round 1 runs without any host synchronisation (`rr_shard_tuples` writes the packed send buffer, K4 writes into
the all-gather buffer); the only read-back of a search is the list of queries that need the second round.

The pack / exchange / unpack helpers are device-agnostic so that the plumbing is tested on CPU
with gloo (tests/test_dist_cpu.py); the compute calls need the CUDA library.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist

TUPLE_BYTES = 32          # int64 row + float64 n + float64 avg + float32 dense + float32 bm25
TUPLE_BYTES_EXTRAS = 40   # + float32 gate factor + float32 raw best-review similarity


def all_reduce_stats(stats, group=None, device="cpu"):
    """Sum / min the per-shard BM25 statistics (engine.BM25Stats.local) over the ranks, in place."""
    import numpy as np
    df = torch.from_numpy(stats.df).to(device)
    fp = torch.from_numpy(stats.first_pos).to(device)
    tot = torch.tensor([stats.total_tokens, stats.n_docs], dtype=torch.int64, device=device)
    dist.all_reduce(df, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(fp, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    stats.df = np.ascontiguousarray(df.cpu().numpy())
    stats.first_pos = np.ascontiguousarray(fp.cpu().numpy())
    stats.total_tokens, stats.n_docs = int(tot[0].item()), int(tot[1].item())
    return stats


def pack_tuples(world: int, grow: torch.Tensor, n: torch.Tensor, avg: torch.Tensor, dense: torch.Tensor,
                bm25: torch.Tensor, gate: Optional[torch.Tensor] = None, best: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, pool] field tensors -> uint8 [world, (B/world)*pool*32]: block g holds, for the query
    slice owned by rank g, the five fields back to back (rows, n, avg, dense, bm25); with the optional per-candidate
    columns of run_search (gate factor, raw best-review similarity; both or neither) the tuple is 40 bytes."""
    B = grow.shape[0]
    assert B % world == 0, "pad the batch to a multiple of the world size"
    assert (gate is None) == (best is None)

    def blk(t):
        return t.contiguous().view(world, -1).view(torch.uint8)
    fields = [blk(grow), blk(n), blk(avg), blk(dense), blk(bm25)]
    if gate is not None:
        fields += [blk(gate), blk(best)]
    return torch.cat(fields, dim=1).contiguous()


def exchange(send: torch.Tensor, group=None) -> torch.Tensor:
    """all-to-all of the packed blocks: recv[s] = block that shard s addressed to this rank."""
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    return recv


def field_views(recv: torch.Tensor, per_rank_queries: int, pool: int, with_extras: bool = False):
    """Views of the received buffer: for each field the [bytes] tensor starting at shard 0's block of
    that field; shard s's block of the same field starts `stride` bytes later."""
    bp = per_rank_queries * pool
    flat = recv.view(-1)
    stride = bp * (TUPLE_BYTES_EXTRAS if with_extras else TUPLE_BYTES)
    off = {"grow": 0, "n": bp * 8, "avg": bp * 16, "dense": bp * 24, "bm25": bp * 28}
    if with_extras:
        off.update(gate=bp * 32, best=bp * 36)
    return {k: flat[v:] for k, v in off.items()}, stride


def unpack_shard(recv: torch.Tensor, shard: int, per_rank_queries: int, pool: int):
    """(grow i64, n f64, avg f64, dense f32, bm25 f32) [per_rank_queries, pool] of one source shard
    (used by tests and debugging; the fusion kernel reads the packed buffer directly)."""
    views, stride = field_views(recv, per_rank_queries, pool)
    bp = per_rank_queries * pool

    def take(name, dtype, esz):
        b = views[name][shard * stride: shard * stride + bp * esz]
        return b.view(dtype).view(per_rank_queries, pool)
    return (take("grow", torch.int64, 8), take("n", torch.float64, 8), take("avg", torch.float64, 8),
            take("dense", torch.float32, 4), take("bm25", torch.float32, 4))


def local_pool(pool: int, world: int) -> int:
    """Tuples each shard sends per query in round 1.  With rows spread evenly, a shard holds
    Binomial(pool, 1/world) of the global pool: mean + 6 sigma, rounded up to 16, capped at pool."""
    if world <= 1:
        return pool
    mean = pool / world
    m = int(math.ceil((mean + 6.0 * math.sqrt(mean)) / 16.0) * 16)
    return min(pool, max(16, m))


class _NoStream:
    """Stands in for a CUDA stream / event when the index lives on the CPU (protocol tests with gloo and a stub index)."""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass


def _on_cuda(ix) -> bool:
    dev = getattr(ix, "device", None)
    return dev is not None and torch.device(dev).type == "cuda"


class _Lane:
    """One batch in flight: its own handle over the shared index tensors (own scratch), its own stream, buffers."""

    def __init__(self, ix, stream):
        self.ix, self.stream = ix, stream
        self.send = None
        self.flags_host = None


class PendingShardedSearch:
    """A sharded search in flight (ShardedSearcher.begin).  `result()` must be called on every rank, in the order
    the searches were begun (its second round, when needed, is collective)."""

    def __init__(self, searcher, lane, args, rows, final, flags_host, done):
        self.s, self.lane, self.args = searcher, lane, args
        self.rows, self.final, self.flags_host, self.done = rows, final, flags_host, done
        self.repeated = 0

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.done is None:
            return self.rows, self.final
        self.done.synchronize()             # the ONE host wait of a search: which queries need the second round
        self.done = None
        q, term_ids, n_terms, fusion, mode = self.args
        idx = torch.nonzero(self.flags_host, as_tuple=False).view(-1)
        nf = int(idx.numel())
        self.repeated = self.s.last_repeated = nf
        lane, G = self.lane, self.s.world
        cuda = not isinstance(lane.stream, _NoStream)
        with (torch.cuda.stream(lane.stream) if cuda else lane.stream):
            if nf > 0:
                # a shard may hold more pool members than it sent, or could not certify its tensor-path result:
                # repeat those queries with m = pool through the synchronous entry points (always exact)
                idx = idx.to(self.rows.device)
                pad = (-nf) % G
                if pad:
                    idx = torch.cat([idx, idx[:1].expand(pad)])
                r2, f2, _ = self.s._round(lane, q[idx].contiguous(),
                                          None if term_ids is None else term_ids[idx].contiguous(),
                                          None if n_terms is None else n_terms[idx].contiguous(), fusion, mode,
                                          fusion.pool, deferred=False, which=idx)
                self.rows = self.rows.clone()
                self.final = self.final.clone()
                self.rows[idx[:nf]] = r2[:nf]
                self.final[idx[:nf]] = f2[:nf]
                lane.stream.synchronize()
        if cuda:
            cur = torch.cuda.current_stream()
            if cur != lane.stream:
                self.rows.record_stream(cur)
                self.final.record_stream(cur)
        return self.rows, self.final


class ShardedSearcher:
    """Hybrid search over a row-sharded corpus; call `search` (or `begin` / `result`) collectively on every rank.

    Two-round distributed top-pool.  Round 1: every shard sends its local exact top-m (m = local_pool,
    e.g. 48 of pool 150 on 8 shards), so the per-query work on a shard (shortlist selection, exact
    rescoring, candidate BM25) shrinks with the shard.  The owner merges G*m tuples and PROVES the
    result: a shard whose m-th similarity is below the merged pool's cut-off cannot hold another pool
    member.  Queries that fail the proof (rows clustered on one shard) are repeated with m = pool,
    which is always exact.  Results are therefore identical to the single-GPU search.

    Batches in flight.  `begin()` enqueues round 1 of a batch -- shortlist GEMM, selection, rescoring, candidate BM25,
    all-to-all, merge + fusion, result all-gather, flag read-back into pinned memory -- on one of `lanes` CUDA streams
    (each with its own scratch handle over the same index tensors) WITHOUT any host wait; `token.result()` waits for
    that batch only.  With two lanes the tail of batch i (latency-bound selection / exchange / host wait) runs under
    the GEMM of batch i+1; a serving loop is `t1 = begin(b1); t2 = begin(b2); r1 = t1.result(); t3 = begin(b3); ...`.
    Optional per-candidate columns of run_search (app/app_product_search.py:285-310): `extras(cand_local_rows, q,
    which) -> (gate f32[b, m] | None, best_raw f32[b, m] | None)` is evaluated on every shard for its own candidates
    and rides in the tuples (gate factor and raw best-review similarity are per-candidate quantities); `which` is None
    for the whole batch or the batch positions (LongTensor[b]) of the queries a second round repeats."""

    def __init__(self, index, group=None, round1_pool: Optional[int] = None, lanes: int = 1, extras=None):
        self.ix = index
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.round1_pool = round1_pool
        self.last_repeated = 0
        self.extras = extras
        self.n_lanes = max(1, int(lanes))
        self._lanes = None
        self._next = 0

    def _lane_list(self):
        if self._lanes is None or self._lanes[0].ix is not self.ix:
            if self.n_lanes == 1 or not _on_cuda(self.ix):
                lanes = [_Lane(self.ix, None)]                  # runs on the caller's current stream
            else:
                # every lane has its own stream, so the caller's stream stays free for the next batch's ingest
                lanes = [_Lane(self.ix if i == 0 else self.ix.view(), torch.cuda.Stream(device=self.ix.device))
                         for i in range(self.n_lanes)]
            self._lanes = lanes
        return self._lanes

    def _merge(self, lane, recv: torch.Tensor, fusion, m: int, B: int, with_extras: bool):
        """Owner side of a round: K4 over the G*m received tuples of this rank's B/G queries, then ONE all-gather
        of a byte buffer [rows int64[Bg,k] | final f32[Bg,k] | flags int32[Bg]] (the kernel writes into views of it)."""
        ix, G = lane.ix, self.world
        k, Bg = fusion.k, B // G
        views, stride = field_views(recv, Bg, m, with_extras)
        o_final, o_flags = Bg * k * 8, Bg * k * 12
        nbytes = (Bg * k * 12 + Bg * 4 + 7) // 8 * 8                       # rows of the gathered buffer stay 8-byte aligned
        mine = torch.empty(nbytes, dtype=torch.uint8, device=recv.device)
        rows = mine[:o_final].view(torch.int64).view(Bg, k)
        final = mine[o_final:o_flags].view(torch.float32).view(Bg, k)
        flags = mine[o_flags:o_flags + Bg * 4].view(torch.int32)
        ix.fuse_sharded(fusion, G, m, stride, Bg, views["dense"], views["bm25"], views["n"], views["avg"], views["grow"],
                        out=(rows, final, flags), gate=views.get("gate"), best=views.get("best"))
        out = torch.empty((G, mine.numel()), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(out.view(-1), mine, group=self.group)
        all_rows = out[:, :o_final].view(torch.int64).reshape(B, k)
        all_final = out[:, o_final:o_flags].view(torch.float32).reshape(B, k)
        all_flags = out[:, o_flags:o_flags + Bg * 4].view(torch.int32).reshape(B)
        return all_rows, all_final, all_flags

    def _round(self, lane, q, term_ids, n_terms, fusion, mode, m, deferred: bool, which=None):
        """deferred=True: no host synchronisation anywhere (queries the tensor path cannot certify come back
        flagged); deferred=False: rr_dense_topk redoes uncertified queries exactly before the exchange."""
        ix, G = lane.ix, self.world
        B = int(q.shape[0])
        with_extras = self.extras is not None
        if deferred and not with_extras:
            send = ix.shard_tuples(q, term_ids, n_terms, m, G, mode)
        else:
            if deferred:
                cand, dense, _cnt, unc = ix.dense_topk(q, m, mode, want_uncertified=True)
            else:
                cand, dense, _cnt = ix.dense_topk(q, m, mode)
                unc = None
            bm25, n, avg, grow = ix.candidate_tuples(term_ids, n_terms, cand)
            if unc is not None:
                grow = torch.where(unc[:, None] != 0, torch.full_like(grow, -2), grow)
            gate = best = None
            if with_extras:
                gate, best = self.extras(cand, q, which)
                gate = torch.ones_like(dense) if gate is None else gate.to(torch.float32)
                best = torch.zeros_like(dense) if best is None else best.to(torch.float32)
            send = pack_tuples(G, grow, n, avg, dense, bm25, gate, best)
        return self._merge(lane, exchange(send, self.group), fusion, m, B, with_extras)

    def begin(self, q: torch.Tensor, term_ids: Optional[torch.Tensor], n_terms: Optional[torch.Tensor], fusion,
              mode: int = 0) -> PendingShardedSearch:
        """Enqueue round 1 of a batch (q float32[B, D] identical on every rank, B % world == 0) without a host wait."""
        G = self.world
        B, pool = int(q.shape[0]), fusion.pool
        if B % G:
            raise ValueError("batch size must be a multiple of the world size")
        m = min(self.round1_pool or local_pool(pool, G), pool)
        lanes = self._lane_list()
        lane = lanes[self._next % len(lanes)]
        self._next += 1
        if not _on_cuda(lane.ix):
            # host-side protocol run (stub index): same rounds, nothing asynchronous
            run = _Lane(lane.ix, _NoStream())
            rows, final, flags = self._round(run, q, term_ids, n_terms, fusion, mode, m, deferred=True)
            lane.flags_host = flags.clone()
            return PendingShardedSearch(self, run, (q, term_ids, n_terms, fusion, mode), rows, final, lane.flags_host, _NoStream())
        cur = torch.cuda.current_stream()
        if lane.stream is None:
            lane_stream = cur
        else:
            lane_stream = lane.stream
            lane_stream.wait_stream(cur)                       # the inputs were produced on the caller's stream
        run = _Lane(lane.ix, lane_stream)
        with torch.cuda.stream(lane_stream):
            rows, final, flags = self._round(run, q, term_ids, n_terms, fusion, mode, m, deferred=True)
            if lane.flags_host is None or lane.flags_host.numel() != B:
                lane.flags_host = torch.empty(B, dtype=torch.int32).pin_memory()
            lane.flags_host.copy_(flags, non_blocking=True)
            done = torch.cuda.Event()
            done.record(lane_stream)
        return PendingShardedSearch(self, run, (q, term_ids, n_terms, fusion, mode), rows, final, lane.flags_host, done)

    def search(self, q: torch.Tensor, term_ids: Optional[torch.Tensor], n_terms: Optional[torch.Tensor], fusion,
               mode: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """q float32[B, D] (identical on every rank, B % world == 0) -> (global rows int64[B, k],
        final float32[B, k]) on every rank."""
        return self.begin(q, term_ids, n_terms, fusion, mode).result()


def make_extras(gate_ix=None, review_ix=None, groups_per_query=None, gate_penalty: float = 0.5):
    """`extras` callback of ShardedSearcher for run_search's per-candidate columns (app/app_product_search.py:285-310),
    evaluated by every shard on ITS OWN candidates (local rows):
      gate  `engine.GateIndex` over the shard's product texts + one gate-group list per query (drop_in.build_gate_groups)
      best  `engine.ReviewIndex` over the reviews of the shard's products: raw best-review similarity, min-max
            normalised after the cross-shard merge by K4 (Fusion.best_is_raw = True)
    The reference's `max_rows` cap on scanned reviews (:343-346) is defined over the reviews of the WHOLE pool and is
    not applied here: use it only where the cap does not bind (its default, 300 000 rows, rarely does)."""
    def extras(cand: torch.Tensor, q: torch.Tensor, which=None):
        gate = best = None
        if gate_ix is not None and groups_per_query is not None:
            groups = groups_per_query if which is None else [groups_per_query[int(i)] for i in which.tolist()]
            gate = gate_ix.factors(groups, cand, gate_penalty)
        if review_ix is not None:
            best, _ = review_ix.best(q, cand, max_rows=None, as_numpy=False)
        return gate, best
    return extras


class GridSearcher:
    """Row shards x query groups.  The `world` ranks form Q query groups of R = world/Q ranks (rank = g*R + s);
    every query group holds the WHOLE corpus cut into R row shards and answers its own slice of the batch with the
    ShardedSearcher protocol inside the group, so the per-query work a shard does (shortlist selection, exact
    rescoring, candidate BM25 -- proportional to the batch, not to the shard) is divided by Q as well.  Q = 1 is
    the plain row-sharded search; Q = world is pure query sharding (corpus replicated, no exchange).  Results are
    identical for every (R, Q) because each query is answered by an exact search over the whole corpus.

    Collective construction: every rank of `dist.group.WORLD` must create the searcher (it creates the sub-groups)."""

    def __init__(self, index, query_groups: int = 1, round1_pool: Optional[int] = None, lanes: int = 1, extras=None):
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.Q = int(query_groups)
        if self.Q < 1 or self.world % self.Q:
            raise ValueError("query_groups must divide the world size")
        self.R = self.world // self.Q
        self.g, self.s = divmod(self.rank, self.R)
        self.row_group = self.col_group = None
        for g in range(self.Q):                         # new_group is collective: every rank creates every group
            grp = dist.new_group([g * self.R + s for s in range(self.R)]) if self.R > 1 else None
            if g == self.g:
                self.row_group = grp
        for s in range(self.R):
            grp = dist.new_group([g * self.R + s for g in range(self.Q)]) if self.Q > 1 else None
            if s == self.s:
                self.col_group = grp
        self.ix = index
        self.inner = (ShardedSearcher(index, group=self.row_group, round1_pool=round1_pool, lanes=lanes, extras=extras)
                      if self.R > 1 else None)
        self.last_repeated = 0

    def begin(self, q, term_ids, n_terms, fusion, mode: int = 0):
        """Batches in flight (plain row sharding only, Q = 1): see ShardedSearcher.begin."""
        if self.Q != 1 or self.inner is None:
            raise ValueError("begin() needs plain row sharding (query_groups = 1, world > 1)")
        return self.inner.begin(q, term_ids, n_terms, fusion, mode)

    @staticmethod
    def layout(rank: int, world: int, query_groups: int) -> Tuple[int, int, int]:
        """(query group, row shard, rows shards per group) of a rank."""
        R = world // query_groups
        g, s = divmod(rank, R)
        return g, s, R

    def search(self, q: torch.Tensor, term_ids: Optional[torch.Tensor], n_terms: Optional[torch.Tensor], fusion,
               mode: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        B, k = int(q.shape[0]), fusion.k
        if B % self.world:
            raise ValueError("batch size must be a multiple of the world size")
        Bq = B // self.Q
        lo = self.g * Bq
        qs = q[lo:lo + Bq]
        ts = None if term_ids is None else term_ids[lo:lo + Bq]
        ns = None if n_terms is None else n_terms[lo:lo + Bq]
        if self.inner is not None:
            rows, final = self.inner.search(qs, ts, ns, fusion, mode)
            self.last_repeated = self.inner.last_repeated
        else:
            rows, final = self.ix.hybrid_search(qs, ts, ns, fusion, mode=mode)
        if self.Q == 1:
            return rows, final
        # one byte buffer [rows int64[Bq,k] | final f32[Bq,k] | pad]: every rank's block starts 8-byte aligned whatever
        # the parity of Bq*k (rows of the gathered buffer are viewed as int64)
        nb_r, nb_f = Bq * k * 8, Bq * k * 4
        nbytes = (nb_r + nb_f + 7) // 8 * 8
        mine = torch.zeros(nbytes, dtype=torch.uint8, device=rows.device)
        mine[:nb_r] = rows.reshape(-1).view(torch.uint8)
        mine[nb_r:nb_r + nb_f] = final.reshape(-1).view(torch.uint8)
        out = torch.empty((self.Q, nbytes), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(out.view(-1), mine, group=self.col_group)
        all_rows = out[:, :nb_r].view(torch.int64).reshape(B, k)
        all_final = out[:, nb_r:nb_r + nb_f].view(torch.float32).reshape(B, k)
        return all_rows, all_final
