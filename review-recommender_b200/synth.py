"""Deterministic synthetic corpora of the shapes BASELINE.json names.

Everything is generated per 1 M-row chunk from a chunk-indexed seed, so any row shard can be
produced independently of the others (rank r of a row-sharded run generates only its rows).

  embeddings  default_rng(1000+chunk).standard_normal((rows, D), f32), rows L2-normalised the
              way the reference does at load time (app/app_product_search.py:110)
  queries     default_rng(2000), same recipe, B x D
  tokens      Zipf(s=1.07) over ranks 1..V, doc length clip(round(lognormal(ln 48, .5)), 4, 256),
              default_rng([3000+chunk, stream]); term id = rank-1, token string = f"t{rank}"
  query terms default_rng(4000): L distinct terms of a random document, topped up by Zipf draws
  metadata    default_rng([5000+chunk, stream]): n_reviews = clip(round(lognormal(ln 12, 1.2)), 1, 5000),
              avg_stars = round(clip(normal(4.1, .6), 1, 5), 3)   (nlp/10_product_prep.py:82)

NumPy versions here are the parity-test inputs (identical arrays go to the CPU oracle and to the
GPU).  `device_*` variants generate the same *distributions* directly in HBM with torch
generators for the full-size benchmark, where a host-side draw of 10 M x 384 would dominate
the run; they are not bit-identical to the NumPy ones and are only used for throughput.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

CHUNK = 1_000_000
ZIPF_S = 1.07


def _l2n(x: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return x / np.maximum(n, 1e-12)


def embeddings(n: int, d: int, row0: int = 0) -> np.ndarray:
    """float32[n, d], unit rows; rows are global rows row0 .. row0+n."""
    out = np.empty((n, d), dtype=np.float32)
    r = row0
    while r < row0 + n:
        chunk, within = divmod(r, CHUNK)
        take = min(CHUNK - within, row0 + n - r)
        rng = np.random.default_rng(1000 + chunk)
        block = rng.standard_normal((within + take, d), dtype=np.float32)[within:]
        out[r - row0:r - row0 + take] = _l2n(block)
        r += take
    return out


def queries(b: int, d: int) -> np.ndarray:
    rng = np.random.default_rng(2000)
    return _l2n(rng.standard_normal((b, d), dtype=np.float32))


def zipf_cdf(v: int) -> np.ndarray:
    p = np.arange(1, v + 1, dtype=np.float64) ** (-ZIPF_S)
    c = np.cumsum(p)
    return c / c[-1]


def _chunk_tokens(chunk: int, v: int, cdf: np.ndarray, rows: int = CHUNK) -> Tuple[np.ndarray, np.ndarray]:
    # separate streams for lengths and tokens so that a row prefix of a chunk is a prefix of both
    rng_len = np.random.default_rng([3000 + chunk, 0])
    rng_tok = np.random.default_rng([3000 + chunk, 1])
    lens = np.clip(np.rint(rng_len.lognormal(np.log(48.0), 0.5, size=rows)), 4, 256).astype(np.int64)
    u = rng_tok.random(int(lens.sum()))
    tok = np.searchsorted(cdf, u, side="left").astype(np.int32)
    np.minimum(tok, v - 1, out=tok)
    return lens, tok


def corpus_tokens(n: int, v: int, row0: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(doc_offsets int64[n+1], token_ids int32[total]) for global rows row0 .. row0+n."""
    cdf = zipf_cdf(v)
    lens_all: List[np.ndarray] = []
    toks_all: List[np.ndarray] = []
    r = row0
    while r < row0 + n:
        chunk, within = divmod(r, CHUNK)
        take = min(CHUNK - within, row0 + n - r)
        lens, tok = _chunk_tokens(chunk, v, cdf, rows=within + take)
        start = int(lens[:within].sum())
        lens_all.append(lens[within:])
        toks_all.append(tok[start:])
        r += take
    lens = np.concatenate(lens_all)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    return offs, np.concatenate(toks_all)


def query_terms(b: int, l: int, doc_offsets: np.ndarray, token_ids: np.ndarray, v: int) -> np.ndarray:
    """int32[b, l]: l distinct terms taken from a random document (topped up with Zipf draws)."""
    rng = np.random.default_rng(4000)
    cdf = zipf_cdf(v)
    n = doc_offsets.shape[0] - 1
    out = np.empty((b, l), dtype=np.int32)
    for i in range(b):
        d = int(rng.integers(0, n))
        toks = token_ids[doc_offsets[d]:doc_offsets[d + 1]]
        uniq = list(dict.fromkeys(toks.tolist()))           # first-appearance order
        rng.shuffle(uniq)
        chosen = uniq[:l]
        while len(chosen) < l:
            t = int(min(np.searchsorted(cdf, rng.random()), v - 1))
            if t not in chosen:
                chosen.append(t)
        out[i] = chosen
    return out


def metadata(n: int, row0: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(n_reviews int64[n], avg_stars float64[n])."""
    nr = np.empty(n, dtype=np.int64)
    av = np.empty(n, dtype=np.float64)
    r = row0
    while r < row0 + n:
        chunk, within = divmod(r, CHUNK)
        take = min(CHUNK - within, row0 + n - r)
        rng_n = np.random.default_rng([5000 + chunk, 0])
        rng_s = np.random.default_rng([5000 + chunk, 1])
        a = np.clip(np.rint(rng_n.lognormal(np.log(12.0), 1.2, size=within + take)), 1, 5000).astype(np.int64)
        s = np.round(np.clip(rng_s.normal(4.1, 0.6, size=within + take), 1.0, 5.0), 3)
        nr[r - row0:r - row0 + take] = a[within:]
        av[r - row0:r - row0 + take] = s[within:]
        r += take
    return nr, av


def token_strings(token_ids: np.ndarray) -> List[str]:
    return [f"t{int(t) + 1}" for t in token_ids]


def corpus_as_lists(doc_offsets: np.ndarray, token_ids: np.ndarray) -> List[List[str]]:
    """The `blob["corpus"]` form of product_bm25.pkl (nlp/12_product_prep.py:85-88)."""
    strs = token_strings(token_ids)
    return [strs[doc_offsets[i]:doc_offsets[i + 1]] for i in range(doc_offsets.shape[0] - 1)]


def skus(n: int, row0: int = 0) -> List[str]:
    return [f"SKU{row0 + i:09d}" for i in range(n)]


@dataclass
class SynthCorpus:
    emb: np.ndarray           # f32[N, D] unit rows
    doc_offsets: np.ndarray   # i64[N+1]
    token_ids: np.ndarray     # i32[total]
    n_reviews: np.ndarray     # i64[N]
    avg_stars: np.ndarray     # f64[N]
    vocab_size: int


def make_corpus(n: int, d: int, v: int, row0: int = 0) -> SynthCorpus:
    offs, toks = corpus_tokens(n, v, row0)
    nr, av = metadata(n, row0)
    return SynthCorpus(embeddings(n, d, row0), offs, toks, nr, av, v)


def query_terms_global(b: int, l: int, n_total: int, v: int) -> np.ndarray:
    """query_terms() with the source documents taken from the first min(n_total, CHUNK) rows of the corpus, so that
    the query set does not depend on how the corpus is sharded (every rank / every GPU count derives the same
    int32[b, l] from the recipe alone)."""
    offs, toks = corpus_tokens(min(int(n_total), CHUNK), v, 0)
    return query_terms(b, l, offs, toks, v)


@dataclass
class SynthChunk:
    row0: int                 # global row of the first row of this piece
    emb: np.ndarray           # f32[n, D] unit rows (reference l2_normalize)
    lens: np.ndarray          # i64[n] document lengths
    token_ids: np.ndarray     # i32[sum(lens)]
    n_reviews: np.ndarray     # i64[n]
    avg_stars: np.ndarray     # f64[n]
    extra: object = None      # whatever `per_chunk` returned (computed in the worker thread)


def chunk_stream(row0: int, n: int, d: int, v: int, workers: int = 0, piece_rows: int = CHUNK, per_chunk=None):
    """Yields the rows [row0, row0+n) of the section-8d recipe as SynthChunk pieces in row order, generated by a pool of
    threads (NumPy's generators, sort and searchsorted release the GIL).  A piece never crosses a 1 M-row recipe
    chunk; at most `workers` pieces are in flight, so host memory stays bounded (~2.2 GB per piece at 384-d).
    `per_chunk(piece) -> extra` runs in the worker thread (used by the tests / bench to fold every piece into the
    sharded CPU oracle while the next ones are generated)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    pieces = []
    r = row0
    while r < row0 + n:
        within = r % CHUNK
        take = min(CHUNK - within, piece_rows, row0 + n - r)
        pieces.append((r, take))
        r += take
    workers = max(1, min(workers or (os.cpu_count() or 1), len(pieces)))

    def make(r0, take):
        offs, toks = corpus_tokens(take, v, r0)
        nr, av = metadata(take, r0)
        piece = SynthChunk(r0, embeddings(take, d, r0) if d > 0 else None, np.diff(offs), toks, nr, av)   # d = 0: BM25-only corpora
        if per_chunk is not None:
            piece.extra = per_chunk(piece)
        return piece

    with ThreadPoolExecutor(workers) as ex:
        pending = []
        it = iter(pieces)
        for _ in range(workers):
            p = next(it, None)
            if p is not None:
                pending.append(ex.submit(make, *p))
        while pending:
            piece = pending.pop(0).result()
            p = next(it, None)
            if p is not None:
                pending.append(ex.submit(make, *p))
            yield piece
