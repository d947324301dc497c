"""Drop-in replacements for the reference's search functions (same names, argument meaning and
error behaviour), backed by librr_b200.so.  NumPy arrays / DataFrames in and out, exactly like the
functions they replace; the GPU index behind them is built once per artifact and cached, the way
the reference caches its own loads (st.cache_resource / st.cache_data).

    reference seam                                                   replacement here
    ---------------------------------------------------------------  ------------------------------
    utils.cosine_similarity_search(q, mat, top_k)      utils.py:111   cosine_similarity_search
    cosine_search(qvec, mat, topk)              app/test.py:125      cosine_search
    _cosine_pool(qvec, mat, pool)   app/app_product_search.py:192    _cosine_pool
    rank_bm25.BM25Okapi(corpus).get_scores(tokens)                   BM25Okapi (also importable as the
        (app/test.py:156,170; app/app_product_search.py:142,206)      module `rank_bm25`, see shims/)
    bm25_scores(bm25, toks, order_idx, top_idx) app/test.py:168      bm25_scores
    _bm25_for_candidates(blob, query, cand_skus)          :201       _bm25_for_candidates
    run_search(query, k, rerank_k, w_*, prior_C, use_snips,
               max_scan, min_reviews, gate_penalty)       :245       SearchEngine.run_search
    search(args)                                app/test.py:228      SearchEngine.search
    _best_snippets(qvec, cand_skus, max_rows)             :320       SearchEngine._best_snippets
    best_review_snippets(qvec, cand_skus, max_rows)  app/test.py:181 SearchEngine.best_review_snippets

There is no CPU fallback: without the CUDA library or a GPU these raise RRError.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, engine
from ._lib import RRError

# utils.py:11-12 (query-side tokenizer; the index-side tokenizer lives in nlp/12_product_prep.py)
TOKEN_RE = re.compile(r"[a-z0-9]+(?:'[a-z0-9]+)?")
STOP_WORDS = {"a", "an", "the", "and", "or", "of", "for", "to", "in", "on", "with", "is", "are", "it", "this", "that"}


def tokenize_query(query: str) -> List[str]:
    """utils.py:57-60."""
    return [t for t in TOKEN_RE.findall(query.lower()) if t not in STOP_WORDS]


# Gate vocabularies: the data tables of utils.py:15-38 (= SYN / COLORS in app/app_product_search.py and
# app/test.py).  Keys are what a query token must equal (synonym sets) / any member must occur in the
# query (colour sets); insertion order matters because the first 6 distinct groups are kept.
GATE_COLORS: Dict[str, frozenset] = {name: frozenset(words.split("|")) for name, words in (
    ("yellow", "yellow|mustard|lemon|gold|golden"), ("red", "red|scarlet|crimson|maroon"),
    ("blue", "blue|navy|cobalt|azure"), ("green", "green|emerald|olive"), ("black", "black"),
    ("white", "white|ivory"), ("pink", "pink|rose"), ("purple", "purple|violet|lavender"),
    ("orange", "orange|amber"), ("brown", "brown|tan|beige|khaki"), ("gray", "gray|grey|charcoal|slate"))}
GATE_SYNONYMS: Dict[str, frozenset] = {name: frozenset(words.split("|")) for name, words in (
    ("sock", "sock|socks"), ("headphone", "headphone|headphones|earphone|earphones|earbud|earbuds|headset"),
    ("keyboard", "keyboard|keyboards"), ("wireless", "wireless|bluetooth"),
    ("noise", "noise cancelling|noise-canceling|noise canceling|anc"), ("cat", "cat|cats|kitten|kittens|kitty"),
    ("dog", "dog|dogs|puppy|puppies"), ("design", "design|pattern|print|graphic|artwork|motif|theme"))}
GATE_FIXED_GROUPS: List[frozenset] = list(GATE_COLORS.values()) + list(GATE_SYNONYMS.values())
GATE_MAX_GROUPS = 6


def build_gate_groups(query: str) -> List[frozenset]:
    """utils.py:62-86 (= _build_gate_groups app/app_product_search.py:211-226, app/test.py:62-78): colour sets
    with a member occurring in the lower-cased query (substring test), then per query token its synonym
    set or, for other tokens of >= 4 characters, the token itself; duplicates dropped, first 6 kept."""
    lowered = query.lower()
    picked: List[frozenset] = [words for words in GATE_COLORS.values() if any(w in lowered for w in words)]
    for tok in tokenize_query(query):
        if tok in GATE_SYNONYMS:
            picked.append(GATE_SYNONYMS[tok])
        elif len(tok) >= 4:
            picked.append(frozenset((tok,)))
    distinct: List[frozenset] = []
    for g in picked:
        if g not in distinct:
            distinct.append(g)
    return distinct[:GATE_MAX_GROUPS]


# --------------------------------------------------------------------------------------------
# dense
# --------------------------------------------------------------------------------------------
class _DenseCacheEntry:
    __slots__ = ("mat", "fingerprint", "ix")

    def __init__(self, mat, fingerprint, ix):
        self.mat, self.fingerprint, self.ix = mat, fingerprint, ix


_dense_cache: Dict[Tuple, _DenseCacheEntry] = {}
_DENSE_CACHE_MAX = 4


def _fingerprint(mat: np.ndarray) -> int:
    """Cheap content check of a cached matrix: CRC of up to 64 evenly spaced rows (catches in-place edits of the
    matrix between calls without re-reading all of it; the reference function is pure)."""
    import zlib
    n = mat.shape[0]
    if n == 0:
        return 0
    rows = np.unique(np.linspace(0, n - 1, min(n, 64)).astype(np.int64))
    return zlib.crc32(np.ascontiguousarray(mat[rows]).tobytes())


def _dense_index_for(mat: np.ndarray, device: str = "cuda:0") -> "engine.HybridIndex":
    """One GPU copy per embeddings matrix OBJECT (the reference passes the same cached `Vn` on every call).  The entry
    holds a reference to the array, so its address cannot be handed to another array while the entry lives, a hit
    requires `entry.mat is mat`, and a row-sample fingerprint catches in-place edits."""
    if not isinstance(mat, np.ndarray) or mat.ndim != 2:
        raise RRError("embeddings_matrix must be a 2-D NumPy array")
    key = (id(mat), device)
    fp = _fingerprint(mat)
    ent = _dense_cache.get(key)
    if ent is not None and ent.mat is mat and ent.fingerprint == fp and ent.mat.shape == (ent.ix.n_docs, ent.ix.dim):
        _dense_cache[key] = _dense_cache.pop(key)                 # most recently used last
        return ent.ix
    if ent is not None:
        _dense_cache.pop(key).ix.close()
    while len(_dense_cache) >= _DENSE_CACHE_MAX:
        _dense_cache.pop(next(iter(_dense_cache))).ix.close()
    ix = engine.HybridIndex(np.ascontiguousarray(mat, dtype=np.float32), device=device,
                            make_bf16=mat.shape[0] >= 65536 and mat.shape[1] <= 384)
    _dense_cache[key] = _DenseCacheEntry(mat, fp, ix)
    return ix


def cosine_similarity_search(query_vector: np.ndarray, embeddings_matrix: np.ndarray, top_k: int
                             ) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:111-124: (indices int64[k], similarities float32[k]) by descending similarity, k clamped
    to the number of rows.  Ties are ordered by ascending row (the reference leaves them unordered)."""
    ix = _dense_index_for(embeddings_matrix)
    k = int(min(int(top_k), ix.n_docs))
    if k <= 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
    idx, sims, _ = ix.dense_topk(np.asarray(query_vector, dtype=np.float32)[None, :], k, _lib.RR_DENSE_EXACT)
    return idx[0].cpu().numpy(), sims[0].cpu().numpy()


def cosine_search(qvec: np.ndarray, mat: np.ndarray, topk: int):
    """app/test.py:125-132."""
    return cosine_similarity_search(qvec, mat, topk)


def _cosine_pool(qvec: np.ndarray, mat: np.ndarray, pool: int):
    """app/app_product_search.py:192-195."""
    return cosine_similarity_search(qvec, mat, pool)


# --------------------------------------------------------------------------------------------
# sparse
# --------------------------------------------------------------------------------------------
def flatten_corpus(corpus: Sequence[Sequence[str]]):
    """list of token lists -> (doc_offsets int64[N+1], token_ids int32[total], {token: id}) with ids assigned in
    first-appearance order (the insertion order of rank_bm25's dicts).  One C-level pass (itertools.chain +
    pandas.factorize) instead of a Python loop per token."""
    import itertools
    import pandas as pd
    n = len(corpus)
    lens = np.fromiter((len(d) for d in corpus), dtype=np.int64, count=n)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    flat = list(itertools.chain.from_iterable(corpus))
    if not flat:
        return offs, np.zeros(0, dtype=np.int32), {}
    codes, uniques = pd.factorize(np.asarray(flat, dtype=object))
    return offs, codes.astype(np.int32), {w: i for i, w in enumerate(uniques.tolist())}


def _on_device(a: np.ndarray, device: str):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class BM25Okapi:
    """rank_bm25.BM25Okapi drop-in: same constructor and `get_scores(tokens) -> float64[N]`.

    The corpus (list of token lists, `blob["corpus"]` of product_bm25.pkl) is turned into integer
    term ids (first-appearance order, like the library's dict), the statistics and the tile-blocked
    postings are built by the host builder (bm25_build.cpp) and scoring runs in K1 on the GPU.
    Scores are accumulated in fp32 (impacts are the correctly rounded fp32 of the library's float64
    terms) and returned as float64, so they agree with the library to ~1e-7 relative; both call
    sites cast to float32 immediately (app/test.py:170, app/app_product_search.py:206).
    """

    def __init__(self, corpus: Sequence[Sequence[str]], tokenizer: Optional[Callable] = None,
                 k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25, device: str = "cuda:0",
                 tile_docs: int = engine.DEFAULT_TILE_DOCS):
        if tokenizer:
            corpus = [tokenizer(doc) for doc in corpus]
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.tokenizer = tokenizer
        self._offs, self._ids, vocab = flatten_corpus(corpus)
        self.corpus_size = len(self._offs) - 1
        if self.corpus_size == 0:
            raise ZeroDivisionError("division by zero")          # what rank_bm25 raises for an empty corpus
        self.vocab = vocab
        self.doc_len = np.diff(self._offs).tolist()
        v = max(1, len(vocab))
        # statistics, forward index and postings are built on the GPU from the integer ids
        gb = engine.GpuIndexBuilder(_on_device(self._offs, device), _on_device(self._ids, device), v, tile_docs)
        stats = gb.local_stats().finalize(epsilon)
        self.avgdl = stats.avgdl
        self.average_idf = stats.average_idf
        self.idf = {w: float(stats.idf[i]) for w, i in vocab.items()}
        placeholder = np.zeros((self.corpus_size, 4), dtype=np.float32)
        self._ix = engine.HybridIndex(placeholder, None, None, v, device=device, stats=stats, k1=k1, b=b,
                                      postings=gb.finish(stats, k1, b), make_bf16=False)

    def term_ids(self, tokens: Sequence[str]) -> List[int]:
        return [self.vocab.get(t, -1) for t in tokens]

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        ids, n = engine.HybridIndex.pack_terms([self.term_ids(list(query))])
        out = self._ix.bm25_get_scores(ids, n)
        return out[0].cpu().numpy().astype(np.float64)

    def get_batch_scores(self, query: Sequence[str], doc_ids: Sequence[int]) -> List[float]:
        """rank_bm25's get_batch_scores: scores of the given documents only (K1 candidate mode)."""
        ids, n = engine.HybridIndex.pack_terms([self.term_ids(list(query))])
        cand = np.asarray(list(doc_ids), dtype=np.int64)[None, :]
        return self._ix.bm25_candidates(ids, n, cand)[0].cpu().numpy().astype(np.float64).tolist()


def ensure_same_order(meta, bm25_skus: Sequence[str]) -> Optional[List[int]]:
    """app/test.py:159-166: the permutation that brings the BM25 documents into meta order (duplicate SKUs in the
    blob: the last one wins), or None when ANY meta SKU is missing from the blob -- bm25_scores then reads the
    scores in blob order (identity)."""
    idx_map = {s: i for i, s in enumerate(bm25_skus)}
    try:
        return [idx_map[s] for s in meta["sku"].astype(str).tolist()]
    except KeyError:
        return None


def bm25_scores(bm25, query_tokens: List[str], order_idx: Optional[List[int]], top_idx: np.ndarray) -> np.ndarray:
    """app/test.py:168-173: full scores cast to float32, optional permutation into meta order, gather."""
    if isinstance(bm25, BM25Okapi):
        # same values without materialising N scores: K1 candidate mode at the documents the gather would read
        top_idx = np.asarray(top_idx, dtype=np.int64)
        docs = top_idx if order_idx is None else np.asarray(order_idx, dtype=np.int64)[top_idx]
        if docs.size and (docs.min() < -bm25.corpus_size or docs.max() >= bm25.corpus_size):
            raise IndexError(f"index {int(docs.max())} is out of bounds for axis 0 with size {bm25.corpus_size}")
        docs = np.where(docs < 0, docs + bm25.corpus_size, docs)
        return np.asarray(bm25.get_batch_scores(query_tokens, docs.tolist()), dtype=np.float32)
    scores_all = np.array(bm25.get_scores(query_tokens), dtype=np.float32)
    if order_idx is not None:
        scores_all = scores_all[np.array(order_idx)]
    return scores_all[top_idx]


def _bm25_for_candidates(bm25_blob, query: str, cand_skus: List[str]) -> np.ndarray:
    """app/app_product_search.py:201-208 (missing blob or empty token list -> zeros; SKU absent from
    the blob -> 0.0; duplicate SKUs: the last one wins)."""
    if not bm25_blob:
        return np.zeros(len(cand_skus), dtype=np.float32)
    toks = tokenize_query(query)
    if not toks:
        return np.zeros(len(cand_skus), dtype=np.float32)
    bm25, skus = bm25_blob["bm25"], bm25_blob["skus"]
    lut = bm25_blob.get("_sku_to_doc")
    if lut is None:
        lut = {skus[i]: i for i in range(len(skus))}
        bm25_blob["_sku_to_doc"] = lut
    docs = [lut.get(str(s), -1) for s in cand_skus]
    if isinstance(bm25, BM25Okapi):
        return np.asarray(bm25.get_batch_scores(toks, docs), dtype=np.float32)     # -1 -> 0.0 on the device
    scores_all = np.array(bm25.get_scores(toks), dtype=np.float32)
    return np.array([scores_all[d] if d >= 0 else 0.0 for d in docs], dtype=np.float32)


# --------------------------------------------------------------------------------------------
# artifact loaders
# --------------------------------------------------------------------------------------------
def load_product_index(emb_path, meta_path, normalize: bool = True):
    """load_product_index app/test.py:134-146 (= _product_index app/app_product_search.py:87-117): the meta frame
    and the row-normalised float32 embeddings; the same SystemExit messages for missing / inconsistent files."""
    import os
    import pandas as pd
    if not os.path.exists(str(emb_path)) or not os.path.exists(str(meta_path)):
        raise SystemExit("[ERR] product_emb.npy and/or product_emb_meta.parquet missing in data/processed/")
    meta = pd.read_parquet(meta_path)
    if "sku" not in meta.columns or "agg_text" not in meta.columns:
        raise SystemExit("[ERR] product_emb_meta.parquet must have 'sku' and 'agg_text'")
    V = np.load(emb_path, mmap_mode="r").astype(np.float32)
    if len(meta) != V.shape[0]:
        raise SystemExit(f"[ERR] length mismatch: meta={len(meta)} vs emb_rows={V.shape[0]}")
    x = np.array(V)
    if not normalize:
        return meta.reset_index(drop=True), x
    Vn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)          # l2_normalize utils.py:40-44
    return meta.reset_index(drop=True), Vn


def load_bm25_blob(bm25_pkl):
    """The pickle half of load_bm25 (app/test.py:148-157): None if the file does not exist."""
    import os
    import pickle
    if not os.path.exists(str(bm25_pkl)):
        return None
    with open(bm25_pkl, "rb") as f:
        return pickle.load(f)


# --------------------------------------------------------------------------------------------
# the two search drivers
# --------------------------------------------------------------------------------------------
class SearchEngine:
    """Holds what `_product_index()` + `_bm25_loader()` hold in the reference and answers queries
    with the reference's signatures.  `encode(query) -> float32[D]` stands in for the sentence
    transformer (`normalize_embeddings=True`), `rerank(query, texts) -> scores` for the cross-encoder
    (None = unavailable: zeros, app/app_product_search.py:275), `gate(text, query, penalty) -> factor`
    for the attribute gates (None = 1.0)."""

    def __init__(self, meta, Vn: np.ndarray, bm25_corpus: Optional[Sequence[Sequence[str]]] = None,
                 bm25_skus: Optional[Sequence[str]] = None, encode: Optional[Callable] = None,
                 rerank: Optional[Callable] = None, gate: Optional[Callable] = None, device: str = "cuda:0",
                 reviews=None, normalize: bool = False):
        """`normalize=True`: `Vn` holds the raw rows of product_emb.npy and is L2-normalised on the device
        (bit-identical to l2_normalize, engine.HybridIndex).  `reviews`: the reviews_with_embeddings.parquet frame (columns sku, text, stars, embedding) or
        None when the file does not exist (snippets are then skipped, like `REV_EMB.exists()` :285)."""
        import pandas as pd
        self.meta = meta.reset_index(drop=True)
        self.encode, self.rerank, self.gate = encode, rerank, gate
        self.reviews = None
        self.review_ix = None
        self.gate_ix = None
        if gate is None and "agg_text" in self.meta.columns:
            self.gate_ix = engine.GateIndex(self.meta["agg_text"].astype(str).tolist(), GATE_FIXED_GROUPS, device=device)
        if reviews is not None and "sku" in reviews.columns and len(reviews):
            self.reviews = reviews.reset_index(drop=True)
            E = np.stack(self.reviews["embedding"].values).astype(np.float32)
            self.review_ix = engine.ReviewIndex(E, self.reviews["sku"].astype(str).tolist(),
                                                self.meta["sku"].astype(str).tolist(), device=device)
        n = len(self.meta)
        if n != Vn.shape[0]:
            raise SystemExit(f"[ERR] length mismatch: meta={n} vs emb_rows={Vn.shape[0]}")     # app/test.py:142
        nrev = pd.to_numeric(self.meta.get("n_reviews", pd.Series([np.nan] * n)), errors="coerce").fillna(0).values
        avg = pd.to_numeric(self.meta.get("avg_stars", pd.Series([np.nan] * n)), errors="coerce").values
        self.vocab: Dict[str, int] = {}
        self.bm25_active = bm25_corpus is not None
        self._ix_cli = None                 # second BM25 alignment, only when the CLI's identity rule applies
        self._cli_identity = False
        self._device = device
        if bm25_corpus is not None:
            # term ids in first-appearance order OF THE BLOB (the library's dict order decides the idf mean);
            # document lengths / statistics are those of the blob, whatever the alignment to meta rows
            self._flat_offs, self._flat_ids, self.vocab = flatten_corpus(bm25_corpus)
            self._n_blob = len(self._flat_offs) - 1
            v = max(1, len(self.vocab))
            self._stats = engine.BM25Stats.local(self._flat_offs, self._flat_ids, v).finalize()
            # Streamlit rule (:207-208): align BM25 documents to meta rows by SKU -- the last duplicate wins, a meta
            # row without a document scores 0.  The CLI's ensure_same_order (app/test.py:159-166) gives the same
            # alignment when every meta SKU is in the blob, and IDENTITY order (row r <-> blob document r) otherwise.
            sku_to_doc = {str(s): i for i, s in enumerate(bm25_skus)}
            doc_of_row = np.fromiter((sku_to_doc.get(s, -1) for s in self.meta["sku"].astype(str).tolist()),
                                     dtype=np.int64, count=n)
            self._cli_identity = bool(np.any(doc_of_row < 0))
            self.ix = self._build_index(Vn, doc_of_row, nrev, avg, normalize, None)
        else:
            self.ix = engine.HybridIndex(Vn, n_reviews=nrev, avg_stars=avg, device=device, normalize=normalize)
        self._nrev, self._avg = nrev, avg

    def _build_index(self, Vn, doc_of_row: np.ndarray, nrev, avg, normalize: bool, share):
        """HybridIndex whose row r holds the tokens of BM25 document doc_of_row[r] (-1: no document, no postings);
        the index itself is built on the GPU from the integer ids."""
        n = len(doc_of_row)
        lens = np.where(doc_of_row >= 0, np.diff(self._flat_offs)[np.maximum(doc_of_row, 0)], 0)
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        src = np.repeat(self._flat_offs[np.maximum(doc_of_row, 0)] - offs[:-1], lens) + np.arange(int(offs[-1]), dtype=np.int64)
        toks = self._flat_ids[src] if len(src) else np.zeros(0, dtype=np.int32)
        v = max(1, len(self.vocab))
        return engine.HybridIndex(Vn, _on_device(offs, self._device), _on_device(toks, self._device), v, nrev, avg,
                                  device=self._device, stats=self._stats, normalize=normalize, share=share)

    def _index_for(self, driver: str) -> "engine.HybridIndex":
        """The index whose BM25 alignment the driver's rule prescribes (see __init__)."""
        if driver == "streamlit" or not (self.bm25_active and self._cli_identity):
            return self.ix
        if self._ix_cli is None:
            n = len(self.meta)
            ident = np.arange(n, dtype=np.int64)
            ident[ident >= self._n_blob] = -1            # the reference raises IndexError when it reads such a row
            self._ix_cli = self._build_index(None, ident, self._nrev, self._avg, False, self.ix)
        return self._ix_cli

    def _terms(self, query: str):
        toks = tokenize_query(query)
        ids = [self.vocab.get(t, -1) for t in toks]
        return toks, engine.HybridIndex.pack_terms([ids])

    # ---- best-review snippets ---------------------------------------------------------------------
    def _snippets_for_rows(self, qvec: np.ndarray, rows: np.ndarray, max_rows: int, text_cap: int):
        """(raw best similarity float32[len(rows)], {sku: {"score", "text", "stars"}}) for product rows."""
        score, file_pos = self.review_ix.best(qvec[None, :], rows[None, :].astype(np.int64), max_rows=max_rows)
        score, file_pos = score[0], file_pos[0]
        snips: Dict[str, Dict] = {}
        skus = self.meta["sku"].astype(str).values
        has_stars = "stars" in self.reviews.columns
        for i, r in enumerate(rows):
            f = int(file_pos[i])
            if f < 0:
                continue
            rev = self.reviews.iloc[f]
            snips[str(skus[r])] = {"score": float(score[i]), "text": str(rev["text"])[:text_cap],
                                   "stars": float(rev["stars"]) if has_stars else float("nan")}
        return score, snips

    def _rows_of_skus(self, cand_skus: Sequence[str]) -> np.ndarray:
        lut = getattr(self, "_sku_row", None)
        if lut is None:
            lut = {}
            for r, s in enumerate(self.meta["sku"].astype(str).tolist()):
                lut.setdefault(s, r)
            self._sku_row = lut
        return np.asarray([lut.get(str(s), -1) for s in cand_skus], dtype=np.int64)

    def _best_snippets(self, qvec: np.ndarray, cand_skus: List[str], max_rows: int = 300_000) -> Dict[str, Dict]:
        """app/app_product_search.py:320-370 (any failure -> {}, like its try/except)."""
        if self.review_ix is None:
            return {}
        if max_rows <= 0:
            return {}            # the reference's np.stack of zero rows raises and is swallowed (:348, :366)
        rows = self._rows_of_skus(cand_skus)
        return self._snippets_for_rows(np.asarray(qvec, dtype=np.float32), rows, max_rows, 600)[1]

    def best_review_snippets(self, qvec: np.ndarray, cand_skus: List[str], max_rows: int = 1_000_000) -> Dict[str, Dict]:
        """app/test.py:181-215."""
        if self.review_ix is None:
            return {}
        rows = self._rows_of_skus(cand_skus)
        if max_rows <= 0 and self.review_ix.cap_limits(rows[None, :], 0) is not None:
            raise ValueError("need at least one array to stack")       # np.stack([]) at app/test.py:206
        return self._snippets_for_rows(np.asarray(qvec, dtype=np.float32), rows, max_rows, 400)[1]

    def _search(self, query: str, fusion: "engine.Fusion", gate_penalty: float, use_snips: bool = False,
                max_scan: int = 300_000):
        return self._search_batch([query], fusion, gate_penalty, use_snips, max_scan)[0]

    def _search_batch(self, queries: Sequence[str], fusion: "engine.Fusion", gate_penalty: float,
                      use_snips: bool = False, max_scan: int = 300_000):
        """The numeric core of run_search / search for a batch of queries in ONE pass over the GPU kernels
        (dense top-pool, candidate BM25, gates, best reviews, fusion); per query the result is what the
        single-query call returns.  -> list of (top-k DataFrame, tokens, snippets)."""
        B = len(queries)
        if B == 0:
            return []
        qmat = np.stack([np.asarray(self.encode(q), dtype=np.float32) for q in queries])
        toks = [tokenize_query(q) for q in queries]
        pool = fusion.pool
        ix = self._index_for(fusion.driver)
        cand, dense, cnt = ix.dense_topk(qmat, pool)
        counts = cnt.cpu().numpy()
        cand_h = cand.cpu().numpy()
        if ix is self._ix_cli and len(self.meta) > self._n_blob and int(cand_h.max()) >= self._n_blob:
            # identity order with a blob shorter than meta: `scores_all[top_idx]` (app/test.py:173) raises
            raise IndexError(f"index {int(cand_h.max())} is out of bounds for axis 0 with size {self._n_blob}")
        if self.bm25_active and any(toks):
            # a query without tokens scores zeros (:203-204) -- an empty term list does exactly that
            tid, nt = engine.HybridIndex.pack_terms([[self.vocab.get(t, -1) for t in tk] for tk in toks])
            bm25, n, avg, grow = ix.candidate_tuples(tid, nt, cand)
        else:
            bm25, n, avg, grow = ix.candidate_tuples(None, None, cand)
        frames = [self.meta.iloc[cand_h[b, :counts[b]]].reset_index(drop=True) for b in range(B)]

        rerank = gate = best = None
        if fusion.rerank_k > 0:
            z = np.zeros((B, pool), dtype=np.float32)
            if self.rerank is not None:
                for b in range(B):
                    rr_k = min(fusion.rerank_k, int(counts[b]))
                    texts = frames[b]["agg_text"].astype(str).str.slice(0, 2000).tolist()[:rr_k]
                    rr = np.array(self.rerank(queries[b], texts), dtype=np.float32)
                    lo, hi = float(np.min(rr)), float(np.max(rr))
                    if np.isfinite(lo) and np.isfinite(hi) and hi - lo >= 1e-12:
                        z[b, :rr_k] = ((rr - lo) / (hi - lo + 1e-12)).astype(np.float32)      # _minmax, :182-187
            rerank = z
        if self.gate_ix is not None:
            # calculate_gate_factor over agg_text[:6000] of the pool on the GPU (:297-302, app/test.py:291-297)
            gate = self.gate_ix.factors([build_gate_groups(q) for q in queries], cand, gate_penalty)
        elif self.gate is not None and "agg_text" in self.meta.columns:
            g = np.ones((B, pool), dtype=np.float32)
            for b in range(B):
                texts = frames[b]["agg_text"].astype(str).str.slice(0, 6000).tolist()
                g[b, :counts[b]] = np.array([self.gate(t, queries[b], gate_penalty) for t in texts], dtype=np.float32)
            gate = g
        snips_all: List[Dict[str, Dict]] = [{} for _ in range(B)]
        if use_snips and self.review_ix is not None:
            text_cap = 600 if fusion.driver == "streamlit" else 400
            if max_scan <= 0 and fusion.driver != "streamlit":
                for b in range(B):      # raises like the CLI when some candidate has reviews
                    self.best_review_snippets(qmat[b], frames[b]["sku"].astype(str).tolist(), max_scan)
            if max_scan > 0:
                raw, file_pos = self.review_ix.best(qmat, cand, max_rows=max_scan)
                skus = self.meta["sku"].astype(str).values
                has_stars = "stars" in self.reviews.columns
                for b in range(B):
                    for i in range(int(counts[b])):
                        f = int(file_pos[b, i])
                        if f >= 0:
                            rev = self.reviews.iloc[f]
                            snips_all[b][str(skus[cand_h[b, i]])] = {
                                "score": float(raw[b, i]), "text": str(rev["text"])[:text_cap],
                                "stars": float(rev["stars"]) if has_stars else float("nan")}
                if any(snips_all):
                    # queries without any snippet keep an all-zero column, whose min-max is zeros as well (:288-294)
                    best = raw
                    fusion.best_is_raw = True
        top_rows, final, pos, comp = ix.fuse(fusion, dense, bm25, n, avg, grow, count=cnt, rerank=rerank,
                                                  best=best, gate=gate, want_components=True)
        pos_h, comp_h = pos.cpu().numpy(), comp.cpu().numpy()
        gate_h = None if gate is None else (gate.cpu().numpy() if hasattr(gate, "cpu") else gate)
        results = []
        for b in range(B):
            p_ = pos_h[b]
            p_ = p_[p_ >= 0]
            c = comp_h[b]
            out = frames[b].iloc[p_].reset_index(drop=True).copy()
            out["_dense"], out["_bm25"], out["_prior"] = c[p_, 0], c[p_, 1], c[p_, 2]
            out["_trust"], out["_final"] = c[p_, 3], c[p_, 4]
            out["_rerank"] = rerank[b][p_] if rerank is not None else 0.0
            out["_best"] = c[p_, 7]
            out["_gate"] = gate_h[b][p_] if gate_h is not None else np.ones(len(p_), dtype=np.float32)
            results.append((out, toks[b], snips_all[b]))
        return results

    def run_search(self, query: str, k: int, rerank_k: int, w_dense: float, w_bm25: float, w_rerank: float,
                   w_prior: float, w_best: float, prior_C: float, use_snips: bool, max_scan: int,
                   min_reviews: int, gate_penalty: float):
        """Signature and return shape of run_search (app/app_product_search.py:245-317):
        (DataFrame of the top k rows in rank order, snippets dict, debug dict)."""
        fusion = engine.Fusion(k=k, rerank_k=rerank_k, w_dense=w_dense, w_bm25=w_bm25, w_rerank=w_rerank,
                               w_prior=w_prior, w_best=w_best, prior_C=prior_C, min_reviews=min_reviews,
                               driver="streamlit", bm25_absent=not self.bm25_active)
        out, toks, snips = self._search(query, fusion, gate_penalty, use_snips=use_snips, max_scan=max_scan)
        return out, snips, {"bm25_active": self.bm25_active, "tokens": toks,
                             "groups": [list(g) for g in build_gate_groups(query)], "pool": fusion.pool}

    def run_search_batch(self, queries: Sequence[str], k: int = 10, rerank_k: int = 0, w_dense: float = 0.55,
                         w_bm25: float = 0.20, w_rerank: float = 0.20, w_prior: float = 0.20, w_best: float = 0.10,
                         prior_C: float = 20.0, use_snips: bool = False, max_scan: int = 300_000, min_reviews: int = 8,
                         gate_penalty: float = 0.5):
        """run_search for a list of queries as one GPU batch: [(DataFrame, snippets, debug)] in query order,
        each element equal to run_search(query, ...).  This is what a batched eval driver calls instead of one
        run_search per (method, query) (evaluate_ranking_methods evals/performance_metrics.py:238-281)."""
        fusion = engine.Fusion(k=k, rerank_k=rerank_k, w_dense=w_dense, w_bm25=w_bm25, w_rerank=w_rerank,
                               w_prior=w_prior, w_best=w_best, prior_C=prior_C, min_reviews=min_reviews,
                               driver="streamlit", bm25_absent=not self.bm25_active)
        res = self._search_batch(list(queries), fusion, gate_penalty, use_snips=use_snips, max_scan=max_scan)
        return [(out, snips, {"bm25_active": self.bm25_active, "tokens": toks,
                              "groups": [list(g) for g in build_gate_groups(q)], "pool": fusion.pool})
                for q, (out, toks, snips) in zip(queries, res)]

    def batched_search_function(self, queries: Sequence[str]):
        """A `search_function(query, **config)` for the reference's evaluate_ranking_methods
        (evals/performance_metrics.py:238-281) that answers every query of `queries` from ONE GPU batch per
        distinct config instead of one run_search call per (method, query)."""
        cache: Dict[Tuple, Dict[str, Tuple]] = {}
        qlist = list(dict.fromkeys(queries))

        def search_function(query, **config):
            key = tuple(sorted(config.items()))
            if key not in cache:
                cache[key] = dict(zip(qlist, self.run_search_batch(qlist, **config)))
            hit = cache[key].get(query)
            return hit if hit is not None else self.run_search_batch([query], **config)[0]
        return search_function

    @classmethod
    def from_artifacts(cls, emb_path, meta_path, bm25_pkl=None, reviews_path=None, **kw) -> "SearchEngine":
        """Builds the engine from the reference's artifact files: product_emb.npy, product_emb_meta.parquet,
        product_bm25.pkl ({"skus", "corpus", "tokenizer"}, nlp/12_product_prep.py:85-88) and, optionally,
        reviews_with_embeddings.parquet -- load_product_index + load_bm25 (app/test.py:134-157)."""
        meta, V = load_product_index(emb_path, meta_path, normalize=False)     # rows are normalised on the device
        blob = load_bm25_blob(bm25_pkl) if bm25_pkl is not None else None
        reviews = None
        if reviews_path is not None:
            import os
            import pandas as pd
            if os.path.exists(str(reviews_path)):
                reviews = pd.read_parquet(reviews_path)
        return cls(meta, V, blob["corpus"] if blob else None, [str(x) for x in blob["skus"]] if blob else None,
                   reviews=reviews, normalize=True, **kw)

    def search(self, args):
        """The numeric core of search(args) (app/test.py:228-309); returns the top-k DataFrame."""
        fusion = engine.Fusion(k=args.k, rerank_k=args.rerank_k, w_dense=args.w_dense, w_bm25=args.w_bm25,
                               w_rerank=args.w_rerank, w_prior=args.w_prior, w_best=args.w_best,
                               prior_C=args.prior_C, driver="cli", bm25_absent=not self.bm25_active)
        out, _, snips = self._search(args.query, fusion, getattr(args, "gate_penalty", 0.5),
                                     use_snips=not getattr(args, "no_snippets", True),
                                     max_scan=getattr(args, "max_reviews_scan", 1_000_000))
        out = out.rename(columns={"_best": "_bestrev"})
        self.last_snippets = snips
        return out
