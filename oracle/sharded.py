"""The reference's run_search on a corpus that does not fit one host array.  TEST INFRASTRUCTURE ONLY.

At BASELINE.json's full sizes (10 M x 384, 20 M documents) the reference cannot be run as written: its `Vn` would
be a 15 GB array and rank_bm25 keeps one dict per document.  This module answers a SAMPLE of queries exactly as the
reference would, by restricting the corpus to a subset that provably contains everything the reference reads:

  * `run_search` only ever looks at the dense top-`pool` rows (app/app_product_search.py:253-255) -- and the top-`pool`
    of the whole corpus is contained in the union of the top-`pool` of its row chunks.  Per chunk we keep the
    top-(pool + margin) rows by `chunk @ q` (the margin absorbs BLAS summation-order differences at the cut);
    `oracle.hybrid.run_search_core` then recomputes `mat @ qvec` over that union and selects, sorts, min-maxes, fuses
    with the reference's own arithmetic (oracle.primitives / oracle.hybrid, both pinned against the reference).
  * BM25: `get_scores` (app/app_product_search.py:206) is evaluated for the kept rows only.  A document's score
    depends on the rest of the corpus only through corpus_size, avgdl and the idf table, which are accumulated over
    ALL chunks (df, first-occurrence positions for the library's dict-order idf mean, token counts) and finalised by
    `oracle.bm25_okapi.idf_from_stats` -- the same code path BM25OkapiCSR uses.  The per-document arithmetic is the
    expression of BM25Okapi.get_scores, term by term in query order, in float64.

`tests/test_oracle_sharded.py` checks on CPU that this equals `run_search_core` over the whole arrays with
`BM25OkapiCSR` (ids and BM25 bit-for-bit, fused scores to the BLAS-shape tolerance).

Chunks may be added from several threads (`add_chunk` is thread-safe) and from several processes: `partial()` is a
picklable summary, `merge()` combines the summaries of all ranks on the rank that runs the comparison.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from .bm25_okapi import B_DEFAULT, EPSILON_DEFAULT, K1_DEFAULT, idf_from_stats
from .hybrid import cli_search_core, run_search_core

INT64_MAX = np.iinfo(np.int64).max


def chunk_doc_stats(lens: np.ndarray, token_ids: np.ndarray, vocab_size: int):
    """(df int64[V], first_pos int64[V] relative to the chunk's first token) of one row chunk."""
    v = int(vocab_size)
    tok = np.asarray(token_ids, dtype=np.int64)
    n = len(lens)
    doc = np.repeat(np.arange(n, dtype=np.int64), lens)
    keys = np.sort(doc * v + tok)
    head = np.empty(len(keys), dtype=bool)
    if len(keys):
        head[0] = True
        np.not_equal(keys[1:], keys[:-1], out=head[1:])
    df = np.bincount(keys[head] % v, minlength=v).astype(np.int64)
    first = np.full(v, INT64_MAX, dtype=np.int64)
    # fancy assignment keeps the LAST value written for a repeated index: walk the tokens backwards
    first[tok[::-1]] = np.arange(len(tok) - 1, -1, -1, dtype=np.int64)
    return df, first


class _ReducedBM25:
    """`get_scores(tokens)` of BM25Okapi over the kept documents of one sample query."""

    def __init__(self, tf: np.ndarray, doc_len: np.ndarray, term_ids: Sequence[int], tokens: Sequence[str], idf, df,
                 avgdl, k1, b):
        self.tf, self.doc_len = tf, doc_len
        self.pos_of: Dict[str, List[int]] = {}
        for j, t in enumerate(tokens):
            self.pos_of.setdefault(t, []).append(j)
        self.term_ids, self.idf, self.df, self.avgdl, self.k1, self.b = list(term_ids), idf, df, avgdl, k1, b

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(len(self.doc_len))
        doc_len = np.array(self.doc_len)
        for q in query:
            if q not in self.pos_of:
                raise KeyError(f"term {q!r} was not registered with the sharded oracle")
            j = self.pos_of[q][0]
            t = self.term_ids[j]
            if t < 0 or t >= len(self.df) or self.df[t] == 0:
                continue                      # unknown term: idf.get(q) is None -> contributes 0
            q_freq = self.tf[:, j]
            score += (self.idf[t] or 0) * (q_freq * (self.k1 + 1) /
                                           (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


class ShardedOracle:
    def __init__(self, q_sample: np.ndarray, qt_sample: np.ndarray, vocab_size: int, pool: int, margin: int = 32,
                 k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT):
        """q_sample float32[S, D] (the sampled queries), qt_sample int[S, L] their term ids (-1 = padding)."""
        self.q = np.ascontiguousarray(q_sample, dtype=np.float32)
        self.qt = np.asarray(qt_sample, dtype=np.int64)
        self.v, self.pool, self.margin = int(vocab_size), int(pool), int(margin)
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self._lock = threading.Lock()
        self.chunks: List[dict] = []
        self.final = None

    # ---- accumulation ------------------------------------------------------------------------------------
    def add_chunk(self, row0: int, emb: np.ndarray, lens: np.ndarray, token_ids: np.ndarray, n_reviews, avg_stars,
                  want_scores_for: Optional[np.ndarray] = None) -> None:
        """One row chunk: global rows row0 .. row0+len(emb); `emb` are the rows of the reference's Vn."""
        S = self.q.shape[0]
        n = emb.shape[0]
        pc = min(n, self.pool + self.margin)
        sims = emb @ self.q.T                                     # float32 [n, S]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        tok = np.asarray(token_ids)
        rows = np.empty((S, pc), dtype=np.int64)
        tf = np.zeros((S, pc, self.qt.shape[1]), dtype=np.int64)
        for i in range(S):
            idx = np.argpartition(-sims[:, i], pc - 1)[:pc] if pc < n else np.arange(n)
            idx = np.sort(idx)
            rows[i] = idx
            terms = self.qt[i]
            for c, d in enumerate(idx):
                seg = tok[offs[d]:offs[d + 1]]
                tf[i, c] = (seg[:, None] == terms[None, :]).sum(axis=0)
        df, first = chunk_doc_stats(lens, tok, self.v)
        rec = dict(row0=int(row0), n=int(n), n_tokens=int(offs[-1]), rows=rows + int(row0), emb=emb[rows],
                   tf=tf, doc_len=np.asarray(lens, dtype=np.int64)[rows], n_reviews=np.asarray(n_reviews)[rows],
                   avg_stars=np.asarray(avg_stars)[rows], df=df, first=first)
        with self._lock:
            self.chunks.append(rec)

    def partial(self) -> List[dict]:
        return self.chunks

    def merge(self, partials: Sequence[List[dict]]) -> "ShardedOracle":
        self.chunks = [c for p in partials for c in p]
        return self

    # ---- global statistics -----------------------------------------------------------------------------------
    def finalize(self) -> "ShardedOracle":
        ch = sorted(self.chunks, key=lambda c: c["row0"])
        r = ch[0]["row0"]
        for c in ch:
            if c["row0"] != r:
                raise ValueError(f"row chunks do not tile the corpus (gap or overlap at row {r})")
            r += c["n"]
        self.n_docs = sum(c["n"] for c in ch)
        total_tokens = sum(c["n_tokens"] for c in ch)
        df = np.zeros(self.v, dtype=np.int64)
        first = np.full(self.v, INT64_MAX, dtype=np.int64)
        pos0 = 0
        for c in ch:
            df += c["df"]
            loc = c["first"]
            first = np.minimum(first, np.where(loc == INT64_MAX, INT64_MAX, loc + pos0))
            pos0 += c["n_tokens"]
        self.df = df
        self.avgdl = total_tokens / self.n_docs                 # python int / int, as the library
        self.idf, self.average_idf, self.n_terms = idf_from_stats(df, first, self.n_docs, self.epsilon)
        self.chunks = ch
        self.final = True
        return self

    # ---- one sampled query, answered like the reference ------------------------------------------------------
    def reduced_world(self, i: int):
        """(V float32[R, D], meta DataFrame incl. `_grow`, bm25 object, skus, tokens) of sample query i."""
        if not self.final:
            self.finalize()
        grow = np.concatenate([c["rows"][i] for c in self.chunks])
        V = np.concatenate([c["emb"][i] for c in self.chunks]).astype(np.float32, copy=False)
        tf = np.concatenate([c["tf"][i] for c in self.chunks])
        doc_len = np.concatenate([c["doc_len"][i] for c in self.chunks])
        skus = [f"SKU{int(g):09d}" for g in grow]
        meta = pd.DataFrame({"sku": skus,
                             "n_reviews": np.concatenate([c["n_reviews"][i] for c in self.chunks]),
                             "avg_stars": np.concatenate([c["avg_stars"][i] for c in self.chunks]),
                             "_grow": grow})
        terms = [int(t) for t in self.qt[i] if t >= 0]
        tokens = [f"t{t + 1}" for t in terms]
        bm25 = _ReducedBM25(tf, doc_len, terms, tokens, self.idf, self.df, self.avgdl, self.k1, self.b)
        return V, meta, bm25, skus, tokens

    def run(self, i: int, k: int, driver: str = "streamlit", **fusion_kw):
        """(top-k DataFrame in rank order with `_grow` = global row, pool DataFrame) of sample query i."""
        V, meta, bm25, skus, tokens = self.reduced_world(i)
        fn = run_search_core if driver == "streamlit" else cli_search_core
        top, pool = fn(self.q[i], V, meta, bm25, skus, tokens, k=k, **fusion_kw)
        return top, pool


class ChunkedBM25Scores:
    """`BM25Okapi.get_scores` over a corpus held as row chunks of flat token ids (configs[3]: 20 M documents):
    global statistics first (`add_chunk` per chunk, then `finalize`), then `get_scores(term_ids)` chunk by chunk.
    Same float64 expression as BM25OkapiCSR.impacts / BM25Okapi.get_scores."""

    def __init__(self, vocab_size: int, k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT):
        self.v, self.k1, self.b, self.epsilon = int(vocab_size), k1, b, epsilon
        self._lock = threading.Lock()
        self.chunks: List[dict] = []

    def add_chunk(self, row0: int, lens: np.ndarray, token_ids: np.ndarray) -> None:
        df, first = chunk_doc_stats(lens, token_ids, self.v)
        rec = dict(row0=int(row0), n=len(lens), lens=np.asarray(lens, dtype=np.int64),
                   tok=np.asarray(token_ids, dtype=np.int32), df=df, first=first)
        with self._lock:
            self.chunks.append(rec)

    def finalize(self) -> "ChunkedBM25Scores":
        ch = sorted(self.chunks, key=lambda c: c["row0"])
        self.n_docs = sum(c["n"] for c in ch)
        df = np.zeros(self.v, dtype=np.int64)
        first = np.full(self.v, INT64_MAX, dtype=np.int64)
        pos0 = 0
        for c in ch:
            df += c["df"]
            first = np.minimum(first, np.where(c["first"] == INT64_MAX, INT64_MAX, c["first"] + pos0))
            pos0 += len(c["tok"])
        self.df = df
        self.avgdl = pos0 / self.n_docs
        self.idf, self.average_idf, _ = idf_from_stats(df, first, self.n_docs, self.epsilon)
        self.chunks = ch
        return self

    def _chunk_scores(self, c: dict, queries: Sequence[Sequence[int]], out: np.ndarray) -> None:
        doc_len = c["lens"]
        doc_of_tok = np.repeat(np.arange(c["n"], dtype=np.int64), doc_len)
        lo = c["row0"] - self.chunks[0]["row0"]
        contrib: Dict[int, np.ndarray] = {}
        for i, term_ids in enumerate(queries):
            score = out[i, lo:lo + c["n"]]
            for q in term_ids:
                q = int(q)
                if q < 0 or q >= self.v or self.df[q] == 0:
                    continue              # unknown term: idf.get(q) is None -> contributes 0
                if q not in contrib:
                    q_freq = np.bincount(doc_of_tok[c["tok"] == q], minlength=c["n"])
                    contrib[q] = (self.idf[q] or 0) * (q_freq * (self.k1 + 1) /
                                                       (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
                score += contrib[q]

    def get_scores_many(self, queries: Sequence[Sequence[int]], workers: int = 0) -> np.ndarray:
        """float64[len(queries), n_docs]; the chunks are scored by a pool of threads."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        out = np.zeros((len(queries), self.n_docs))
        workers = max(1, min(workers or (os.cpu_count() or 1), len(self.chunks)))
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda c: self._chunk_scores(c, queries, out), self.chunks))
        return out

    def get_scores(self, term_ids: Sequence[int]) -> np.ndarray:
        return self.get_scores_many([list(term_ids)], workers=1)[0]


def compare_with_oracle(so: ShardedOracle, sample: Sequence[int], rows: np.ndarray, final: np.ndarray, k: int,
                        driver: str = "streamlit", rtol: float = 1e-5, **fusion_kw) -> dict:
    """Position-by-position comparison of the library's top-k (`rows` int64[B, k] global rows, `final` float32[B, k],
    indexed by the batch positions in `sample`) with the oracle's answer for each sampled query.
    bit_exact_rate = identical id at the same rank / all ranks; ids that differ are `tie_explained` when the
    oracle's fused scores at the two ranks involved agree within `rtol` (the reference's sort_values is unstable,
    so such positions are interchangeable), otherwise they count as `unexplained`."""
    n_pos = n_same = n_tie = n_bad = 0
    max_rel = 0.0
    set_overlap = []
    for j, b in enumerate(sample):
        top, _ = so.run(j, k, driver, **fusion_kw)
        ref_rows = top["_grow"].values.astype(np.int64)
        ref_final = top["_final"].values.astype(np.float64)
        kk = len(ref_rows)
        got_rows, got_final = np.asarray(rows[b][:kk]), np.asarray(final[b][:kk], dtype=np.float64)
        scale = max(1e-3, float(np.max(np.abs(ref_final)))) if kk else 1.0
        max_rel = max(max_rel, float(np.max(np.abs(got_final - ref_final) / np.maximum(np.abs(ref_final), 1e-3 * scale))) if kk else 0.0)
        same = got_rows == ref_rows
        n_pos += kk
        n_same += int(same.sum())
        for p in np.nonzero(~same)[0]:
            tied = abs(got_final[p] - ref_final[p]) <= rtol * scale + 1e-7
            # the id the library put here must sit at a rank whose oracle score is tied with this rank's
            where = np.nonzero(ref_rows == got_rows[p])[0]
            tied = tied and (len(where) == 0 or abs(ref_final[where[0]] - ref_final[p]) <= rtol * scale + 1e-7)
            n_tie += int(tied)
            n_bad += int(not tied)
        set_overlap.append(len(set(got_rows.tolist()) & set(ref_rows.tolist())) / max(kk, 1))
    return {"queries": len(sample), "positions": n_pos, "bit_exact_rate": n_same / max(n_pos, 1),
            "tie_explained": n_tie, "unexplained": n_bad, "max_rel_fused": max_rel,
            "min_set_overlap": min(set_overlap) if set_overlap else 1.0}
