"""Host side of the hybrid-retrieval engine: index construction and batched search.

PyTorch is used for device memory and streams only; every computation on the search path is a
call into librr_b200.so (`_lib`).  Nothing here imports the oracle and nothing falls back to
NumPy: without the library (or without a CUDA device) construction raises.

Reference seams (paths relative to the reference root):
  * index inputs     product_emb.npy rows L2-normalised at load (app/app_product_search.py:98-110),
                     product_bm25.pkl corpus (nlp/12_product_prep.py:85-88), n_reviews / avg_stars
                     of product_emb_meta.parquet (nlp/11_build_product_embeddings.py:86-90)
  * HybridIndex.hybrid_search   the numeric core of run_search (app/app_product_search.py:253-312)
                                and of search (app/test.py:238-309), batched
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FusionParams, IndexDesc, RRError, check

K1_DEFAULT, B_DEFAULT, EPSILON_DEFAULT = 1.5, 0.75, 0.25      # rank_bm25.BM25Okapi defaults
DEFAULT_TILE_DOCS = int(__import__("os").environ.get("RR_TILE_DOCS", "16384"))
INT64_MAX = np.iinfo(np.int64).max


def _ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# --------------------------------------------------------------------------------------------
# BM25 statistics and postings (host builder in bm25_build.cpp)
# --------------------------------------------------------------------------------------------
@dataclass
class BM25Stats:
    """Corpus-global statistics.  For a row-sharded corpus every rank computes `local`, the
    fields are all-reduced (sum / min / sum / sum) and `finalize` is called on the result."""
    vocab_size: int
    df: np.ndarray                      # int64[V]
    first_pos: np.ndarray               # int64[V], global flat token position of first occurrence
    total_tokens: int = 0
    n_docs: int = 0
    idf: Optional[np.ndarray] = None    # float64[V]
    average_idf: float = 0.0
    avgdl: float = 0.0
    epsilon: float = EPSILON_DEFAULT

    @staticmethod
    def local(doc_offsets: np.ndarray, token_ids: np.ndarray, vocab_size: int, token_pos0: int = 0) -> "BM25Stats":
        lib = _lib.load()
        doc_offsets = np.ascontiguousarray(doc_offsets, dtype=np.int64)
        token_ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        n = doc_offsets.shape[0] - 1
        df = np.zeros(vocab_size, dtype=np.int64)
        fp = np.full(vocab_size, INT64_MAX, dtype=np.int64)
        tot = C.c_int64(0)
        check(lib.rr_bm25_local_stats(_ptr(doc_offsets), _ptr(token_ids), n, vocab_size, token_pos0,
                                      _ptr(df), _ptr(fp), C.byref(tot)))
        return BM25Stats(vocab_size, df, fp, int(tot.value), int(n))

    def finalize(self, epsilon: float = EPSILON_DEFAULT) -> "BM25Stats":
        lib = _lib.load()
        if self.n_docs <= 0:
            raise RRError("BM25 statistics over an empty corpus")
        self.idf = np.zeros(self.vocab_size, dtype=np.float64)
        avg = C.c_double(0.0)
        check(lib.rr_bm25_idf(_ptr(self.df), _ptr(self.first_pos), self.vocab_size, self.n_docs, epsilon,
                              _ptr(self.idf), C.byref(avg)))
        self.average_idf = float(avg.value)
        self.avgdl = self.total_tokens / self.n_docs       # python int / int, as rank_bm25
        self.epsilon = epsilon
        return self


@dataclass
class HostPostings:
    data: np.ndarray        # uint64[nnz]   {u32 doc, f32 impact}
    tile_base: np.ndarray   # uint64[n_tiles+1]
    blk_off: np.ndarray     # uint32[n_tiles*(V+1)]
    n_tiles: int
    tile_docs: int
    vocab_size: int
    fwd_off: Optional[np.ndarray] = None    # uint64[N+1]
    fwd_data: Optional[np.ndarray] = None   # uint64[fwd_off[N]]  {u32 term, f32 impact}


def build_postings(doc_offsets: np.ndarray, token_ids: np.ndarray, stats: BM25Stats,
                   k1: float = K1_DEFAULT, b: float = B_DEFAULT, tile_docs: int = DEFAULT_TILE_DOCS,
                   n_threads: int = 0) -> HostPostings:
    lib = _lib.load()
    if stats.idf is None:
        raise RRError("BM25Stats.finalize() has not been called")
    doc_offsets = np.ascontiguousarray(doc_offsets, dtype=np.int64)
    token_ids = np.ascontiguousarray(token_ids, dtype=np.int32)
    n = doc_offsets.shape[0] - 1
    h = C.c_void_p(0)
    check(lib.rr_bm25_build_postings(_ptr(doc_offsets), _ptr(token_ids), n, stats.vocab_size, _ptr(stats.idf),
                                     stats.avgdl, k1, b, tile_docs, n_threads, C.byref(h)))
    try:
        nnz = int(lib.rr_postings_nnz(h))
        n_tiles = int(lib.rr_postings_n_tiles(h))

        def view(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()
        data = view(lib.rr_postings_data(h), nnz, np.uint64)
        tile_base = view(lib.rr_postings_tile_base(h), n_tiles + 1, np.uint64)
        blk_off = view(lib.rr_postings_blk_off(h), n_tiles * (stats.vocab_size + 1), np.uint32)
        fwd_off = view(lib.rr_postings_fwd_off(h), n + 1, np.uint64)
        fwd_data = view(lib.rr_postings_fwd_data(h), int(fwd_off[-1]) if n > 0 else 0, np.uint64)
    finally:
        lib.rr_postings_free(h)
    return HostPostings(data, tile_base, blk_off, n_tiles, tile_docs, stats.vocab_size, fwd_off, fwd_data)


# --------------------------------------------------------------------------------------------
# fusion parameters
# --------------------------------------------------------------------------------------------
@dataclass
class Fusion:
    """Arguments of run_search (app/app_product_search.py:245-248) / search (app/test.py:345-358)
    that reach the numeric core.  `driver` picks the pool floor (150 Streamlit :253, 100 CLI :238)
    and whether the trust factor is applied (:303,309 vs app/test.py:308)."""
    k: int = 10
    rerank_k: int = 0
    w_dense: float = 0.55
    w_bm25: float = 0.20
    w_rerank: float = 0.20
    w_prior: float = 0.20
    w_best: float = 0.10
    prior_C: float = 20.0
    min_reviews: int = 8
    saturation: int = 80
    driver: str = "streamlit"
    bm25_absent: bool = False

    @property
    def pool(self) -> int:
        return max(self.k, self.rerank_k, 150 if self.driver == "streamlit" else 100)

    def to_c(self) -> FusionParams:
        return FusionParams(self.w_dense, self.w_bm25, self.w_rerank, self.w_prior, self.w_best, self.prior_C,
                            int(self.min_reviews), int(self.saturation), 1 if self.driver == "streamlit" else 0,
                            1 if self.rerank_k > 0 else 0,
                            1 if (self.bm25_absent and self.driver != "streamlit") else 0,
                            int(self.k), int(self.pool))


# --------------------------------------------------------------------------------------------
# the index
# --------------------------------------------------------------------------------------------
class HybridIndex:
    """One row shard of the corpus resident in HBM: fp32 + bf16 embeddings, tile-blocked CSR
    postings, per-doc metadata.  All tensors are owned here; the C handle borrows them."""

    def __init__(self, emb, doc_offsets: Optional[np.ndarray] = None, token_ids: Optional[np.ndarray] = None,
                 vocab_size: int = 0, n_reviews=None, avg_stars=None, device: str | torch.device = "cuda:0",
                 row_offset: int = 0, stats: Optional[BM25Stats] = None, k1: float = K1_DEFAULT,
                 b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT, tile_docs: int = DEFAULT_TILE_DOCS,
                 make_bf16: bool = True, postings: Optional[HostPostings] = None, forward_index: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RRError("HybridIndex needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        if isinstance(emb, np.ndarray):
            emb = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32))
        self.emb = emb.to(self.device, dtype=torch.float32).contiguous()
        self.n_docs, self.dim = int(self.emb.shape[0]), int(self.emb.shape[1])
        self.row_offset = int(row_offset)
        self.dim_pad = 0
        self.emb_bf16 = None
        self.max_row_norm = float(torch.linalg.vector_norm(self.emb, dim=1).max().item()) if self.n_docs else 0.0
        if make_bf16:
            self.dim_pad = (self.dim + 63) // 64 * 64
            bf = torch.zeros((self.n_docs, self.dim_pad), dtype=torch.bfloat16, device=self.device)
            step = 1 << 20
            for r in range(0, self.n_docs, step):
                bf[r:r + step, :self.dim] = self.emb[r:r + step].to(torch.bfloat16)   # round-to-nearest-even
            self.emb_bf16 = bf

        self.vocab_size = 0
        self.stats = stats
        self.k1, self.b = k1, b
        self.post = self.tile_base = self.blk_off = self.fwd_off = self.fwd_data = None
        self.tile_docs, self.n_tiles = 0, 0
        if postings is None and doc_offsets is not None and vocab_size > 0:
            if stats is None:
                stats = BM25Stats.local(doc_offsets, token_ids, vocab_size).finalize(epsilon)
                self.stats = stats
            postings = build_postings(doc_offsets, token_ids, stats, k1, b, tile_docs)
        if postings is not None:
            self.vocab_size = postings.vocab_size
            self.tile_docs, self.n_tiles = postings.tile_docs, postings.n_tiles
            self.post = torch.from_numpy(postings.data.view(np.int64)).to(self.device)
            self.tile_base = torch.from_numpy(postings.tile_base.view(np.int64)).to(self.device)
            self.blk_off = torch.from_numpy(postings.blk_off.view(np.int32)).to(self.device)
            if self.post.numel() == 0:
                self.post = torch.zeros(2, dtype=torch.int64, device=self.device)
            if forward_index and postings.fwd_off is not None:
                self.fwd_off = torch.from_numpy(postings.fwd_off.view(np.int64)).to(self.device)
                fd = postings.fwd_data if postings.fwd_data.size else np.zeros(1, dtype=np.uint64)
                self.fwd_data = torch.from_numpy(fd.view(np.int64)).to(self.device)

        def meta(x, nan_to_zero):
            if x is None:
                return None
            a = np.asarray(x, dtype=np.float64)
            if nan_to_zero:
                a = np.where(np.isnan(a), 0.0, a)          # .fillna(0), app/app_product_search.py:264
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.n_reviews = meta(n_reviews, True)
        self.avg_stars = meta(avg_stars, False)

        desc = IndexDesc(self.n_docs, self.row_offset, self.dim, self.dim_pad,
                         self.emb.data_ptr(), self.emb_bf16.data_ptr() if self.emb_bf16 is not None else None,
                         self.max_row_norm, self.vocab_size, self.tile_docs, self.n_tiles,
                         self.post.data_ptr() if self.post is not None else None,
                         self.tile_base.data_ptr() if self.tile_base is not None else None,
                         self.blk_off.data_ptr() if self.blk_off is not None else None,
                         self.fwd_off.data_ptr() if self.fwd_off is not None else None,
                         self.fwd_data.data_ptr() if self.fwd_data is not None else None,
                         self.n_reviews.data_ptr() if self.n_reviews is not None else None,
                         self.avg_stars.data_ptr() if self.avg_stars is not None else None)
        h = C.c_void_p(0)
        check(self.lib.rr_index_create(C.byref(h), C.byref(desc), self.device.index or 0))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.rr_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------------
    def _dev(self, x, dtype) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to(self.device, dtype=dtype).contiguous()

    @staticmethod
    def pack_terms(term_lists: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
        """ragged term-id lists -> (int32[B, Lmax] padded with -1, int32[B])."""
        b = len(term_lists)
        lmax = max(1, max((len(t) for t in term_lists), default=1))
        ids = np.full((b, lmax), -1, dtype=np.int32)
        n = np.zeros(b, dtype=np.int32)
        for i, t in enumerate(term_lists):
            n[i] = len(t)
            if len(t):
                ids[i, :len(t)] = np.asarray(t, dtype=np.int32)
        return ids, n

    # ---- sparse ----------------------------------------------------------------------------
    def bm25_get_scores(self, term_ids, n_terms) -> torch.Tensor:
        """float32[B, n_docs] (a view of a [B, ld] buffer): batched BM25Okapi.get_scores."""
        term_ids = self._dev(term_ids, torch.int32)
        n_terms = self._dev(n_terms, torch.int32)
        B, lmax = int(term_ids.shape[0]), int(term_ids.shape[1])
        ld = (self.n_docs + 3) // 4 * 4
        out = torch.empty((B, ld), dtype=torch.float32, device=self.device)
        check(self.lib.rr_bm25_get_scores(self._h, _ptr(term_ids), _ptr(n_terms), B, lmax, _ptr(out), ld, _stream()))
        return out[:, :self.n_docs]

    def bm25_candidates(self, term_ids, n_terms, cand) -> torch.Tensor:
        term_ids = self._dev(term_ids, torch.int32)
        n_terms = self._dev(n_terms, torch.int32)
        cand = self._dev(cand, torch.int64)
        B, pool = int(cand.shape[0]), int(cand.shape[1])
        out = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        check(self.lib.rr_bm25_candidates(self._h, _ptr(term_ids), _ptr(n_terms), B, int(term_ids.shape[1]),
                                          _ptr(cand), pool, _ptr(out), _stream()))
        return out

    # ---- dense -----------------------------------------------------------------------------
    def dense_topk(self, q, pool: int, mode: int = _lib.RR_DENSE_AUTO):
        """(idx int64[B, pool] local rows, sims float32[B, pool], count int32[B])."""
        q = self._dev(q, torch.float32)
        if q.dim() == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise RRError(f"query dim {q.shape[1]} != index dim {self.dim}")
        B = int(q.shape[0])
        idx = torch.empty((B, pool), dtype=torch.int64, device=self.device)
        sims = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        cnt = torch.empty((B,), dtype=torch.int32, device=self.device)
        check(self.lib.rr_dense_topk(self._h, _ptr(q), B, pool, mode, _ptr(idx), _ptr(sims), _ptr(cnt), _stream()))
        return idx, sims, cnt

    def dense_stats(self) -> dict:
        st = _lib.DenseStats()
        check(self.lib.rr_dense_last_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    # ---- candidates / fusion ------------------------------------------------------------------
    def candidate_tuples(self, term_ids, n_terms, cand):
        """BM25 + metadata + global rows of candidate local rows: (bm25 f32, n f64, avg f64, grow i64)."""
        cand = self._dev(cand, torch.int64)
        B, pool = int(cand.shape[0]), int(cand.shape[1])
        lmax = 0
        if term_ids is not None:
            term_ids = self._dev(term_ids, torch.int32)
            n_terms = self._dev(n_terms, torch.int32)
            lmax = int(term_ids.shape[1])
        bm25 = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        n = torch.empty((B, pool), dtype=torch.float64, device=self.device)
        avg = torch.empty((B, pool), dtype=torch.float64, device=self.device)
        grow = torch.empty((B, pool), dtype=torch.int64, device=self.device)
        check(self.lib.rr_candidate_tuples(self._h, _ptr(term_ids), _ptr(n_terms), B, lmax, _ptr(cand), pool,
                                           _ptr(bm25), _ptr(n), _ptr(avg), _ptr(grow), _stream()))
        return bm25, n, avg, grow

    def fuse(self, fusion: Fusion, dense, bm25, n, avg, grow, count=None, rerank=None, best=None, gate=None,
             want_components: bool = False):
        """K4 on candidate tuples [B, n_in] -> (rows int64[B,k], final f32[B,k], pos int32[B,k], components)."""
        B, n_in = int(dense.shape[0]), int(dense.shape[1])
        p = fusion.to_c()
        rows = torch.empty((B, fusion.k), dtype=torch.int64, device=self.device)
        final = torch.empty((B, fusion.k), dtype=torch.float32, device=self.device)
        pos = torch.empty((B, fusion.k), dtype=torch.int32, device=self.device)
        comp = torch.empty((B, fusion.pool, 8), dtype=torch.float32, device=self.device) if want_components else None

        def opt(x, dt):
            return None if x is None else self._dev(x, dt)
        rerank, best, gate = opt(rerank, torch.float32), opt(best, torch.float32), opt(gate, torch.float32)
        count = opt(count, torch.int32)
        check(self.lib.rr_fuse_topk(C.byref(p), B, n_in, _ptr(count), _ptr(dense), _ptr(bm25), _ptr(n), _ptr(avg),
                                    _ptr(grow), _ptr(rerank), _ptr(best), _ptr(gate), _ptr(rows), _ptr(final),
                                    _ptr(pos), _ptr(comp), self.device.index or 0, _stream()))
        return rows, final, pos, comp

    def fuse_sharded(self, fusion: Fusion, n_shards: int, per_shard: int, shard_stride_bytes: int, B: int,
                     dense, bm25, n, avg, grow):
        """K4 over tuples received from `n_shards` row shards (cross-shard merge + fusion).  The field
        tensors are views into one exchange buffer; shard s's [B, per_shard] block of a field starts
        s*shard_stride_bytes after the field's base."""
        p = fusion.to_c()
        rows = torch.empty((B, fusion.k), dtype=torch.int64, device=self.device)
        final = torch.empty((B, fusion.k), dtype=torch.float32, device=self.device)
        flags = torch.zeros((B,), dtype=torch.int32, device=self.device)
        check(self.lib.rr_fuse_topk_sharded(C.byref(p), B, n_shards, per_shard, shard_stride_bytes, _ptr(dense),
                                            _ptr(bm25), _ptr(n), _ptr(avg), _ptr(grow), _ptr(rows), _ptr(final),
                                            _ptr(flags), self.device.index or 0, _stream()))
        return rows, final, flags

    # ---- one-shot ------------------------------------------------------------------------------
    def hybrid_search(self, q, term_ids, n_terms, fusion: Fusion, mode: int = _lib.RR_DENSE_AUTO):
        """Device tensors in, device tensors out: (global rows int64[B,k], final f32[B,k])."""
        q = self._dev(q, torch.float32)
        B = int(q.shape[0])
        lmax = 0
        if term_ids is not None:
            term_ids = self._dev(term_ids, torch.int32)
            n_terms = self._dev(n_terms, torch.int32)
            lmax = int(term_ids.shape[1])
        p = fusion.to_c()
        rows = torch.empty((B, fusion.k), dtype=torch.int64, device=self.device)
        final = torch.empty((B, fusion.k), dtype=torch.float32, device=self.device)
        check(self.lib.rr_hybrid_search(self._h, _ptr(q), _ptr(term_ids), _ptr(n_terms), B, lmax, C.byref(p), mode,
                                        _ptr(rows), _ptr(final), _stream()))
        return rows, final

    def hybrid_search_host(self, q: np.ndarray, term_ids: Optional[np.ndarray], n_terms: Optional[np.ndarray],
                           fusion: Fusion, mode: int = _lib.RR_DENSE_AUTO, out_rows: Optional[np.ndarray] = None,
                           out_final: Optional[np.ndarray] = None):
        """Host (NumPy / pinned) buffers in and out through rr_hybrid_search_host; copies happen
        inside the call.  This is what the reference-facing drop-ins call."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        B = int(q.shape[0])
        lmax = 0
        if term_ids is not None:
            term_ids = np.ascontiguousarray(term_ids, dtype=np.int32)
            n_terms = np.ascontiguousarray(n_terms, dtype=np.int32)
            lmax = int(term_ids.shape[1])
        p = fusion.to_c()
        rows = out_rows if out_rows is not None else np.empty((B, fusion.k), dtype=np.int64)
        final = out_final if out_final is not None else np.empty((B, fusion.k), dtype=np.float32)
        check(self.lib.rr_hybrid_search_host(self._h, _ptr(q), _ptr(term_ids), _ptr(n_terms), B, lmax, C.byref(p),
                                             mode, _ptr(rows), _ptr(final), _stream()))
        return rows, final


def profile_enable(on: bool) -> None:
    check(_lib.load().rr_profile_enable(1 if on else 0))


def profile_collect() -> Dict[str, Tuple[float, int]]:
    """{kernel class: (summed device milliseconds, launches)} since the last collect."""
    n = len(_lib.PROF_CLASSES)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(_lib.load().rr_profile_collect(ms, cnt, n))
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.PROF_CLASSES)}


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().rr_launch_count(1 if reset else 0))
