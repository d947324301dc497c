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
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FusionParams, IndexDesc, RRError, check

K1_DEFAULT, B_DEFAULT, EPSILON_DEFAULT = 1.5, 0.75, 0.25      # rank_bm25.BM25Okapi defaults
# 12288 docs = 48 KB of fp32 accumulators per CTA -> 4 resident CTAs/SM (r02 sweep at 8 M docs, B = 64:
# 16384 -> 62 %, 12288 -> 69 %, 8192 -> 60 % of measured HBM peak)
DEFAULT_TILE_DOCS = int(__import__("os").environ.get("RR_TILE_DOCS", "12288"))
INT64_MAX = np.iinfo(np.int64).max


def _ptr(t) -> Optional[int]:
    """Address of a tensor / array for a c_void_p parameter or struct field (None = NULL).  This sits on the
    single-query latency path (five arrays per host call): `ndarray.ctypes.data` costs 1.2-2 us, the buffer-protocol
    route 0.5 us; read-only, empty or non-contiguous arrays take the slow route."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    if isinstance(t, np.ndarray):
        try:
            return C.addressof(C.c_char.from_buffer(t))
        except (TypeError, ValueError, BufferError):
            return t.ctypes.data
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
    data: np.ndarray        # uint64[nnz]   {u32 doc, f32 impact}: tile-blocked frequent region, then the rare lists
    tile_base: np.ndarray   # uint64[n_tiles+1]
    dir: np.ndarray         # uint32[n_tiles*(n_freq+1)]
    term_slot: np.ndarray   # int32[V]   directory slot of a frequent term, -1 = rare
    rare_off: np.ndarray    # uint64[V+1]
    n_freq: int
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
        n_freq = int(lib.rr_postings_n_freq(h))

        def view(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()
        data = view(lib.rr_postings_data(h), nnz, np.uint64)
        tile_base = view(lib.rr_postings_tile_base(h), n_tiles + 1, np.uint64)
        dir_ = view(lib.rr_postings_dir(h), n_tiles * (n_freq + 1), np.uint32)
        term_slot = view(lib.rr_postings_term_slot(h), stats.vocab_size, np.int32)
        rare_off = view(lib.rr_postings_rare_off(h), stats.vocab_size + 1, np.uint64)
        fwd_off = view(lib.rr_postings_fwd_off(h), n + 1, np.uint64)
        fwd_data = view(lib.rr_postings_fwd_data(h), int(fwd_off[-1]) if n > 0 else 0, np.uint64)
    finally:
        lib.rr_postings_free(h)
    return HostPostings(data, tile_base, dir_, term_slot, rare_off, n_freq, n_tiles, tile_docs, stats.vocab_size, fwd_off,
                        fwd_data)


@dataclass
class DevicePostings:
    """The same index as HostPostings, built on the GPU (GpuIndexBuilder) and already resident in HBM."""
    data: torch.Tensor        # int64[nnz]   {u32 doc, f32 impact}
    tile_base: torch.Tensor   # int64[n_tiles+1]
    dir: torch.Tensor         # int32[n_tiles*(n_freq+1)]
    term_slot: torch.Tensor   # int32[V]
    rare_off: torch.Tensor    # int64[V+1]
    n_freq: int
    n_tiles: int
    tile_docs: int
    vocab_size: int
    fwd_off: torch.Tensor     # int64[N+1]
    fwd_data: torch.Tensor    # int64[fwd_off[N]]  {u32 term, f32 impact}


class GpuIndexBuilder:
    """BM25Okapi.__init__ (app/test.py:156, app/app_product_search.py:142) for a tokenised corpus that already lives in
    device memory: statistics, forward index and tile-blocked postings are built by rr_bm25_gpu_build_* and are
    bit-identical to the host builder's.  Usage (row-sharded: all-reduce the statistics between the two calls):

        b = GpuIndexBuilder(doc_offsets_dev, token_ids_dev, V)
        stats = b.local_stats(token_pos0); [dist.all_reduce_stats(stats)]; stats.finalize()
        postings = b.finish(stats)          # DevicePostings, pass as HybridIndex(postings=...)
    """

    def __init__(self, doc_offsets: torch.Tensor, token_ids: torch.Tensor, vocab_size: int,
                 tile_docs: int = DEFAULT_TILE_DOCS):
        self.lib = _lib.load()
        if not (doc_offsets.is_cuda and token_ids.is_cuda):
            raise RRError("GpuIndexBuilder needs device tensors (use build_postings for a host corpus)")
        self.device = doc_offsets.device
        self.doc_offsets = doc_offsets.to(torch.int64).contiguous()
        self.token_ids = token_ids.to(torch.int32).contiguous()
        self.n_docs = int(self.doc_offsets.numel()) - 1
        self.n_tokens = int(self.token_ids.numel())
        self.vocab_size, self.tile_docs = int(vocab_size), int(tile_docs)
        nu, npost, nt, nf = C.c_int64(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            check(self.lib.rr_bm25_gpu_build_begin(C.byref(h), _ptr(self.doc_offsets), _ptr(self.token_ids), self.n_docs,
                                                   self.n_tokens, self.vocab_size, self.tile_docs, C.byref(nu), C.byref(npost),
                                                   C.byref(nt), C.byref(nf), self.device.index or 0, _stream()))
        self._h = h
        self.n_unique, self.n_postings, self.n_tiles = int(nu.value), int(npost.value), int(nt.value)
        self.n_freq = int(nf.value)

    def local_stats(self, token_pos0: int = 0) -> BM25Stats:
        df = torch.zeros(self.vocab_size, dtype=torch.int64, device=self.device)
        fp = torch.full((self.vocab_size,), INT64_MAX, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rr_bm25_gpu_build_stats(self._h, int(token_pos0), _ptr(df), _ptr(fp), _stream()))
        return BM25Stats(self.vocab_size, df.cpu().numpy(), fp.cpu().numpy(), self.n_tokens, self.n_docs)

    def finish(self, stats: BM25Stats, k1: float = K1_DEFAULT, b: float = B_DEFAULT) -> DevicePostings:
        if stats.idf is None:
            raise RRError("BM25Stats.finalize() has not been called")
        dev = self.device
        idf = torch.from_numpy(np.ascontiguousarray(stats.idf, dtype=np.float64)).to(dev)
        data = torch.empty(max(self.n_postings, 2), dtype=torch.int64, device=dev)
        tile_base = torch.empty(self.n_tiles + 1, dtype=torch.int64, device=dev)
        dir_ = torch.empty(max(self.n_tiles * (self.n_freq + 1), 1), dtype=torch.int32, device=dev)
        term_slot = torch.empty(self.vocab_size, dtype=torch.int32, device=dev)
        rare_off = torch.empty(self.vocab_size + 1, dtype=torch.int64, device=dev)
        fwd_off = torch.empty(self.n_docs + 1, dtype=torch.int64, device=dev)
        fwd_data = torch.empty(max(self.n_unique, 1), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.rr_bm25_gpu_build_finish(self._h, _ptr(idf), float(stats.avgdl), float(k1), float(b), _ptr(data),
                                                    _ptr(tile_base), _ptr(dir_), _ptr(term_slot), _ptr(rare_off), _ptr(fwd_off),
                                                    _ptr(fwd_data), _stream()))
            torch.cuda.current_stream().synchronize()            # `idf` and the builder scratch may be released now
        self.close()
        return DevicePostings(data[:max(self.n_postings, 2)], tile_base, dir_, term_slot, rare_off, self.n_freq, self.n_tiles,
                              self.tile_docs, self.vocab_size, fwd_off, fwd_data)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.rr_bm25_gpu_build_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
    best_is_raw: bool = False       # `best` handed to fuse() holds raw similarities, K4 min-max normalises them

    @property
    def pool(self) -> int:
        return max(self.k, self.rerank_k, 150 if self.driver == "streamlit" else 100)

    def to_c(self) -> FusionParams:
        return FusionParams(self.w_dense, self.w_bm25, self.w_rerank, self.w_prior, self.w_best, self.prior_C,
                            int(self.min_reviews), int(self.saturation), 1 if self.driver == "streamlit" else 0,
                            1 if self.rerank_k > 0 else 0,
                            1 if (self.bm25_absent and self.driver != "streamlit") else 0,
                            int(self.k), int(self.pool), 1 if self.best_is_raw else 0)


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
                 make_bf16: bool = True, postings: Optional[HostPostings] = None, forward_index: bool = True,
                 normalize: bool = False, share: Optional["HybridIndex"] = None):
        """`normalize=True`: `emb` holds the raw rows of product_emb.npy; they are L2-normalised on the device
        exactly like the reference does at load (rr_normalize_rows), fused with the bf16 copy.
        `share`: reuse the embedding tensors (fp32 + bf16) of another index over the same rows (`emb` is ignored)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RRError("HybridIndex needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        if share is not None:
            emb, normalize, make_bf16 = share.emb, False, False
        if isinstance(emb, np.ndarray):
            emb = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32))
        self.emb = emb.to(self.device, dtype=torch.float32).contiguous()
        self.n_docs, self.dim = int(self.emb.shape[0]), int(self.emb.shape[1])
        self.row_offset = int(row_offset)
        self.dim_pad = 0
        self.emb_bf16 = None
        if normalize and self.n_docs:
            if self.emb.data_ptr() == (emb.data_ptr() if isinstance(emb, torch.Tensor) else 0):
                self.emb = self.emb.clone()                      # never normalise the caller's tensor in place
            bf = None
            if make_bf16:
                self.dim_pad = (self.dim + 63) // 64 * 64
                bf = torch.empty((self.n_docs, self.dim_pad), dtype=torch.bfloat16, device=self.device)
            check(self.lib.rr_normalize_rows(_ptr(self.emb), self.n_docs, self.dim, _ptr(self.emb), _ptr(bf), self.dim_pad,
                                             _ptr(None), self.device.index or 0, _stream()))
            self.emb_bf16 = bf
        if share is not None:
            self.emb_bf16, self.dim_pad, self.max_row_norm = share.emb_bf16, share.dim_pad, share.max_row_norm
        else:
            self.max_row_norm = 0.0
            if self.n_docs:
                mx = torch.empty(1, dtype=torch.float32, device=self.device)
                check(self.lib.rr_max_row_norm(_ptr(self.emb), self.n_docs, self.dim, _ptr(mx), self.device.index or 0, _stream()))
                self.max_row_norm = float(mx.item()) * (1.0 + 1e-6)      # float accumulation slack: the bound must not be low
        if make_bf16 and self.emb_bf16 is None:
            self.dim_pad = (self.dim + 63) // 64 * 64
            bf = torch.empty((self.n_docs, self.dim_pad), dtype=torch.bfloat16, device=self.device)
            if self.n_docs:
                # bf16 copy only (rows already normalised by the caller): same kernel, norms not applied
                check(self.lib.rr_bf16_rows(_ptr(self.emb), self.n_docs, self.dim, _ptr(bf), self.dim_pad,
                                            self.device.index or 0, _stream()))
            self.emb_bf16 = bf

        self.vocab_size = 0
        self.stats = stats
        self.k1, self.b = k1, b
        self.post = self.tile_base = self.dir = self.term_slot = self.rare_off = self.fwd_off = self.fwd_data = None
        self.tile_docs, self.n_tiles, self.n_freq = 0, 0, 0
        if postings is None and isinstance(doc_offsets, torch.Tensor) and doc_offsets.is_cuda and vocab_size > 0:
            # corpus already in device memory: build the index on the GPU
            gb = GpuIndexBuilder(doc_offsets, token_ids, vocab_size, tile_docs)
            if stats is None:
                stats = gb.local_stats().finalize(epsilon)
                self.stats = stats
            postings = gb.finish(stats, k1, b)
        if postings is None and doc_offsets is not None and vocab_size > 0:
            if stats is None:
                stats = BM25Stats.local(doc_offsets, token_ids, vocab_size).finalize(epsilon)
                self.stats = stats
            postings = build_postings(doc_offsets, token_ids, stats, k1, b, tile_docs)
        if isinstance(postings, DevicePostings):
            self.vocab_size = postings.vocab_size
            self.tile_docs, self.n_tiles, self.n_freq = postings.tile_docs, postings.n_tiles, postings.n_freq
            self.post, self.tile_base, self.dir = postings.data, postings.tile_base, postings.dir
            self.term_slot, self.rare_off = postings.term_slot, postings.rare_off
            if forward_index:
                self.fwd_off, self.fwd_data = postings.fwd_off, postings.fwd_data
        elif postings is not None:
            self.vocab_size = postings.vocab_size
            self.tile_docs, self.n_tiles, self.n_freq = postings.tile_docs, postings.n_tiles, postings.n_freq
            self.post = torch.from_numpy(postings.data.view(np.int64)).to(self.device)
            self.tile_base = torch.from_numpy(postings.tile_base.view(np.int64)).to(self.device)
            d_ = postings.dir if postings.dir.size else np.zeros(1, dtype=np.uint32)
            self.dir = torch.from_numpy(d_.view(np.int32)).to(self.device)
            self.term_slot = torch.from_numpy(postings.term_slot).to(self.device)
            self.rare_off = torch.from_numpy(postings.rare_off.view(np.int64)).to(self.device)
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
                         self.max_row_norm, self.vocab_size, self.tile_docs, self.n_tiles, self.n_freq,
                         self.post.data_ptr() if self.post is not None else None,
                         self.tile_base.data_ptr() if self.tile_base is not None else None,
                         self.dir.data_ptr() if self.dir is not None else None,
                         self.term_slot.data_ptr() if self.term_slot is not None else None,
                         self.rare_off.data_ptr() if self.rare_off is not None else None,
                         self.fwd_off.data_ptr() if self.fwd_off is not None else None,
                         self.fwd_data.data_ptr() if self.fwd_data is not None else None,
                         self.n_reviews.data_ptr() if self.n_reviews is not None else None,
                         self.avg_stars.data_ptr() if self.avg_stars is not None else None)
        self._desc = desc
        h = C.c_void_p(0)
        check(self.lib.rr_index_create(C.byref(h), C.byref(desc), self.device.index or 0))
        self._h = h

    def index_bytes(self) -> Dict[str, int]:
        """Bytes of the BM25 index in HBM: postings (both regions), metadata that locates them (tile bases, tile
        directory of the frequent terms, term slots, rare-list offsets) and the forward index."""
        def nb(t):
            return 0 if t is None else int(t.numel()) * t.element_size()
        return {"postings": nb(self.post), "directory": nb(self.tile_base) + nb(self.dir) + nb(self.term_slot) + nb(self.rare_off),
                "forward": nb(self.fwd_off) + nb(self.fwd_data), "n_freq": int(self.n_freq), "n_tiles": int(self.n_tiles)}

    def view(self) -> "HybridIndex":
        """A second handle over the SAME device buffers with its own scratch memory, so that two batches can be in
        flight on two CUDA streams (the handle only borrows the index tensors; this object keeps them alive)."""
        import copy
        other = copy.copy(self)                      # shares every tensor attribute
        h = C.c_void_p(0)
        check(self.lib.rr_index_create(C.byref(h), C.byref(self._desc), self.device.index or 0))
        other._h = h
        return other

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
    def bm25_get_scores(self, term_ids, n_terms, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """float32[B, n_docs] (a view of a [B, ld] buffer): batched BM25Okapi.get_scores.  `out`: a float32 [B, ld] CUDA
        buffer to reuse (ld = n_docs rounded up to a multiple of 4), e.g. when timing -- a fresh multi-GB allocation
        per call costs more than the kernel."""
        term_ids = self._dev(term_ids, torch.int32)
        n_terms = self._dev(n_terms, torch.int32)
        B, lmax = int(term_ids.shape[0]), int(term_ids.shape[1])
        ld = (self.n_docs + 3) // 4 * 4
        if out is None:
            out = torch.empty((B, ld), dtype=torch.float32, device=self.device)
        elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, ld)):
            raise RRError(f"out must be a contiguous float32 CUDA tensor of shape ({B}, {ld})")
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
    def dense_topk(self, q, pool: int, mode: int = _lib.RR_DENSE_AUTO, want_uncertified: bool = False):
        """(idx int64[B, pool] local rows, sims float32[B, pool], count int32[B]).  want_uncertified=True: no host
        synchronisation, a fourth tensor int32[B] marks the queries whose pool is not proven exact (not redone)."""
        q = self._dev(q, torch.float32)
        if q.dim() == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise RRError(f"query dim {q.shape[1]} != index dim {self.dim}")
        B = int(q.shape[0])
        idx = torch.empty((B, pool), dtype=torch.int64, device=self.device)
        sims = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        cnt = torch.empty((B,), dtype=torch.int32, device=self.device)
        if want_uncertified:
            unc = torch.empty((B,), dtype=torch.int32, device=self.device)
            check(self.lib.rr_dense_topk_deferred(self._h, _ptr(q), B, pool, mode, _ptr(idx), _ptr(sims), _ptr(cnt),
                                                  _ptr(unc), _stream()))
            return idx, sims, cnt, unc
        check(self.lib.rr_dense_topk(self._h, _ptr(q), B, pool, mode, _ptr(idx), _ptr(sims), _ptr(cnt), _stream()))
        return idx, sims, cnt

    def debug_bf16_scores(self, q, row0: int, n_rows: int) -> torch.Tensor:
        """float32[B, n_rows]: the raw tensor-core (bf16 x bf16 -> fp32) scores of rows row0 .. row0+n_rows, B <= 128."""
        q = self._dev(q, torch.float32)
        if q.dim() == 1:
            q = q[None, :]
        out = torch.full((int(q.shape[0]), int(n_rows)), float("nan"), dtype=torch.float32, device=self.device)
        check(self.lib.rr_dense_debug_bf16_scores(self._h, _ptr(q), int(q.shape[0]), int(row0), int(n_rows), _ptr(out),
                                                  _stream()))
        return out

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
        # tuple fields are device pointers for the kernel: anything else (NumPy, CPU tensor, strided view, wrong dtype)
        # is converted here rather than handed over as a raw pointer
        dense, bm25 = self._dev(dense, torch.float32), self._dev(bm25, torch.float32)
        n, avg, grow = self._dev(n, torch.float64), self._dev(avg, torch.float64), self._dev(grow, torch.int64)
        check(self.lib.rr_fuse_topk(C.byref(p), B, n_in, _ptr(count), _ptr(dense), _ptr(bm25), _ptr(n), _ptr(avg),
                                    _ptr(grow), _ptr(rerank), _ptr(best), _ptr(gate), _ptr(rows), _ptr(final),
                                    _ptr(pos), _ptr(comp), self.device.index or 0, _stream()))
        return rows, final, pos, comp

    def shard_tuples(self, q, term_ids, n_terms, m: int, n_ranks: int, mode: int = _lib.RR_DENSE_AUTO,
                     send: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The shard's exact top-m + candidate tuples for all B queries, packed for the all-to-all
        (uint8 [n_ranks, (B/n_ranks)*m*32], layout of dist.pack_tuples); no host synchronisation."""
        q = self._dev(q, torch.float32)
        B = int(q.shape[0])
        lmax = 0
        if term_ids is not None:
            term_ids = self._dev(term_ids, torch.int32)
            n_terms = self._dev(n_terms, torch.int32)
            lmax = int(term_ids.shape[1])
        nbytes = (B // n_ranks) * m * 32
        if send is None or send.numel() != n_ranks * nbytes:
            send = torch.empty((n_ranks, nbytes), dtype=torch.uint8, device=self.device)
        check(self.lib.rr_shard_tuples(self._h, _ptr(q), _ptr(term_ids), _ptr(n_terms), B, lmax, m, mode, n_ranks,
                                       _ptr(send), _stream()))
        return send

    def fuse_sharded(self, fusion: Fusion, n_shards: int, per_shard: int, shard_stride_bytes: int, B: int,
                     dense, bm25, n, avg, grow, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                     gate=None, best=None):
        """K4 over tuples received from `n_shards` row shards (cross-shard merge + fusion).  The field
        tensors are views into one exchange buffer; shard s's [B, per_shard] block of a field starts
        s*shard_stride_bytes after the field's base."""
        for name, t, dt in (("dense", dense, torch.float32), ("bm25", bm25, torch.float32), ("n", n, torch.float64),
                            ("avg", avg, torch.float64), ("grow", grow, torch.int64)):
            # views into the exchange buffer: must already be device memory of this index (no silent host pointers)
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device and t.is_contiguous()):
                raise RRError(f"fuse_sharded: `{name}` must be a contiguous CUDA tensor on {self.device}")
        p = fusion.to_c()
        if out is not None:
            rows, final, flags = out
        else:
            rows = torch.empty((B, fusion.k), dtype=torch.int64, device=self.device)
            final = torch.empty((B, fusion.k), dtype=torch.float32, device=self.device)
            flags = torch.zeros((B,), dtype=torch.int32, device=self.device)
        check(self.lib.rr_fuse_topk_sharded(C.byref(p), B, n_shards, per_shard, shard_stride_bytes, _ptr(dense),
                                            _ptr(bm25), _ptr(n), _ptr(avg), _ptr(grow), _ptr(gate), _ptr(best), _ptr(rows),
                                            _ptr(final), _ptr(flags), self.device.index or 0, _stream()))
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

    def hybrid_search_begin(self, q, term_ids, n_terms, fusion: Fusion, mode: int = _lib.RR_DENSE_AUTO) -> "PendingSearch":
        """Enqueue a whole hybrid search on the current stream without any host synchronisation and return a token;
        `token.result()` waits for it and repeats the (rare) queries the tensor path could not certify.  Several
        tokens may be pending at once, each on its own stream and its own handle (`view()`)."""
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
        unc = torch.empty((B,), dtype=torch.int32, device=self.device)
        check(self.lib.rr_hybrid_search_deferred(self._h, _ptr(q), _ptr(term_ids), _ptr(n_terms), B, lmax, C.byref(p),
                                                 mode, _ptr(rows), _ptr(final), _ptr(unc), _stream()))
        n_unc = unc.sum()                                        # enqueued on the same stream, read in result()
        done = torch.cuda.Event()
        done.record()
        return PendingSearch(self, q, term_ids, n_terms, fusion, mode, rows, final, unc, n_unc, done,
                             torch.cuda.current_stream())

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


class PendingSearch:
    """A hybrid search in flight (HybridIndex.hybrid_search_begin)."""

    def __init__(self, ix, q, term_ids, n_terms, fusion, mode, rows, final, unc, n_unc, done, stream):
        self.ix, self.q, self.term_ids, self.n_terms, self.fusion, self.mode = ix, q, term_ids, n_terms, fusion, mode
        self.rows, self.final, self.unc, self.n_unc, self.done = rows, final, unc, n_unc, done
        self.stream = stream            # the stream the search was enqueued on: its outputs live (allocator-wise) there
        self.repeated = 0

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(global rows int64[B, k], final float32[B, k]) -- exact, like hybrid_search."""
        if self.done is not None:
            self.done.synchronize()
            self.done = None
            nf = int(self.n_unc.item())
            self.repeated = nf
            if nf > 0:
                # the repeat and the scatter run on the stream the token was created on, whatever is current now
                with torch.cuda.stream(self.stream):
                    idx = torch.nonzero(self.unc, as_tuple=False).view(-1)
                    r2, f2 = self.ix.hybrid_search(self.q[idx].contiguous(),
                                                   None if self.term_ids is None else self.term_ids[idx].contiguous(),
                                                   None if self.n_terms is None else self.n_terms[idx].contiguous(),
                                                   self.fusion, self.mode)
                    self.rows[idx] = r2
                    self.final[idx] = f2
                self.stream.synchronize()
            cur = torch.cuda.current_stream()
            if cur != self.stream:          # the caller reads the outputs on its own stream
                self.rows.record_stream(cur)
                self.final.record_stream(cur)
        return self.rows, self.final


class ReviewIndex:
    """Review embeddings grouped by product, resident in HBM, for best-review scoring (_best_snippets
    app/app_product_search.py:320-370, best_review_snippets app/test.py:181-215).

    `rev_emb` float32[M, D] and `rev_skus` (M strings) are the `embedding` / `sku` columns of
    reviews_with_embeddings.parquet in file order; `product_skus` are the `sku` column of
    product_emb_meta in row order.  Reviews are regrouped so that the reviews of one SKU are contiguous
    slots in file order (the order `groupby("sku")` + `argmax` sees, :355-356); rows are L2-normalised
    here once (the reference does it per call on the selected subset, :349) in float32 like _l2norm."""

    def __init__(self, rev_emb: np.ndarray, rev_skus: Sequence[str], product_skus: Sequence[str],
                 device: str | torch.device = "cuda:0"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RRError("ReviewIndex needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        rev_emb = np.asarray(rev_emb, dtype=np.float32)
        if rev_emb.ndim != 2 or rev_emb.shape[0] != len(rev_skus):
            raise RRError("rev_emb must be [M, D] with one row per review")
        group_of_sku: Dict[str, int] = {}
        row_group = np.empty(len(product_skus), dtype=np.int64)
        for r, sku in enumerate(product_skus):
            row_group[r] = group_of_sku.setdefault(str(sku), len(group_of_sku))
        n_groups = len(group_of_sku)
        rev_group = np.fromiter((group_of_sku.get(str(x), -1) for x in rev_skus), dtype=np.int64, count=len(rev_skus))
        keep = np.nonzero(rev_group >= 0)[0]
        order = keep[np.argsort(rev_group[keep], kind="stable")]          # file order inside a SKU
        counts = np.bincount(rev_group[order], minlength=n_groups).astype(np.int64)
        off = np.zeros(n_groups + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        self.slot_file = np.ascontiguousarray(order, dtype=np.int64)      # slot -> file position
        self.group_count = counts
        self.group_off = off
        self.row_group = row_group
        rng = np.stack([off[row_group], off[row_group + 1]], axis=1) if len(row_group) else np.zeros((0, 2), np.int64)
        e = rev_emb[order]
        nrm = np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-12)  # _l2norm, :179-180 (float32 throughout)
        e = np.ascontiguousarray(e / nrm, dtype=np.float32)
        self.n_products, self.dim, self.n_slots = len(row_group), int(rev_emb.shape[1]), int(len(order))
        self.emb = (torch.from_numpy(e) if self.n_slots else torch.zeros((1, max(self.dim, 1)))).to(self.device)
        self.range = torch.from_numpy(np.ascontiguousarray(rng)).to(self.device)
        self.slot_file_dev = torch.from_numpy(self.slot_file if self.n_slots else np.zeros(1, np.int64)).to(self.device)

    def cap_limits(self, cand: np.ndarray, max_rows: Optional[int]) -> Optional[np.ndarray]:
        """int64[B]: first file position the `max_rows` cap drops (`sub_meta.iloc[:max_rows]`, :343-346), or
        None when no query selects more than max_rows reviews.  The selection is `sku.isin(set(cand_skus))`."""
        if max_rows is None:
            return None
        lim = np.full(cand.shape[0], INT64_MAX, dtype=np.int64)
        hit = False
        for b in range(cand.shape[0]):
            rows = cand[b][(cand[b] >= 0) & (cand[b] < self.n_products)]
            groups = np.unique(self.row_group[rows])
            total = int(self.group_count[groups].sum())
            if total > max_rows:
                hit = True
                files = np.concatenate([self.slot_file[self.group_off[g]:self.group_off[g + 1]] for g in groups])
                lim[b] = np.sort(files)[max_rows] if max_rows > 0 else -1
        return lim if hit else None

    def best(self, q, cand, max_rows: Optional[int] = None, as_numpy: bool = True):
        """(best similarity float32[B, pool], file position of that review int64[B, pool]; 0 / -1 when the
        product has no review).  `cand` int64[B, pool] product rows."""
        q = q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        q = q.to(self.device, dtype=torch.float32).contiguous()
        if q.dim() == 1:
            q = q[None, :]
        if int(q.shape[1]) != self.dim:
            raise RRError(f"query dim {int(q.shape[1])} != review embedding dim {self.dim}")
        cand_t = cand if isinstance(cand, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cand, dtype=np.int64))
        cand_t = cand_t.to(self.device, dtype=torch.int64).contiguous()
        B, pool = int(cand_t.shape[0]), int(cand_t.shape[1])
        limit = self.cap_limits(cand_t.cpu().numpy(), max_rows) if max_rows is not None else None
        limit_t = torch.from_numpy(limit).to(self.device) if limit is not None else None
        score = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        slot = torch.empty((B, pool), dtype=torch.int64, device=self.device)
        check(self.lib.rr_best_review_scores(_ptr(self.emb), _ptr(self.range), self.n_products, self.dim, _ptr(q), B,
                                             _ptr(cand_t), pool,
                                             _ptr(self.slot_file_dev) if limit_t is not None else _ptr(None),
                                             _ptr(limit_t), _ptr(score), _ptr(slot), self.device.index or 0, _stream()))
        file_pos = torch.where(slot >= 0, self.slot_file_dev[slot.clamp_min(0)], slot) if self.n_slots else slot
        if as_numpy:
            return score.cpu().numpy(), file_pos.cpu().numpy()
        return score, file_pos


class GateIndex:
    """Product text resident in HBM for the attribute gates (calculate_gate_factor utils.py:88-101 over
    `agg_text[:6000]`, app/app_product_search.py:297-302).  `texts` is the agg_text column in row order
    (anything; `str()` is applied like `.astype(str)`).  `fixed_groups` are query-independent groups (the
    COLORS / SYNONYMS sets) whose per-row answers are precomputed into a bitmap at load."""

    def __init__(self, texts: Sequence, fixed_groups: Sequence[Sequence[str]] = (), device: str | torch.device = "cuda:0",
                 max_chars: int = 6000):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RRError("GateIndex needs a CUDA device (no CPU fallback)")
        if len(fixed_groups) > 32:
            raise RRError("at most 32 fixed gate groups")
        self.device = torch.device(device)
        enc = [str(t)[:max_chars].lower().encode("utf-8") for t in texts]
        n = len(enc)
        lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=n)
        if n and int(lens.max()) > np.iinfo(np.int32).max:
            raise RRError("text too long")
        padded = (lens + 15) // 16 * 16
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(padded, out=off[1:])
        blob = np.zeros(int(off[-1]) + 16, dtype=np.uint8)
        for i, b in enumerate(enc):
            if b:
                blob[off[i]:off[i] + len(b)] = np.frombuffer(b, dtype=np.uint8)
        self.n_docs = n
        self.text = torch.from_numpy(blob).to(self.device)
        self.text_off = torch.from_numpy(off[:-1].copy() if n else np.zeros(1, np.int64)).to(self.device)
        self.text_len = torch.from_numpy(lens.astype(np.int32) if n else np.zeros(1, np.int32)).to(self.device)
        self.fixed_groups = [frozenset(g) for g in fixed_groups]
        self.fixed_bits = None
        if self.fixed_groups and n:
            pat, pat_off, gp_off, _ = self._encode([[g for g in self.fixed_groups]])
            bits = torch.zeros(n, dtype=torch.int32, device=self.device)
            check(self.lib.rr_gate_fixed_bitmaps(_ptr(self.text), _ptr(self.text_off), _ptr(self.text_len), n,
                                                 _ptr(pat), _ptr(pat_off), _ptr(gp_off), len(self.fixed_groups),
                                                 _ptr(bits), self.device.index or 0, _stream()))
            self.fixed_bits = bits

    def _encode(self, groups_per_query: Sequence[Sequence[Sequence[str]]]):
        """-> device tensors (pattern bytes, pat_off, group_pat_off, query_group_off) + host group_fixed list."""
        pat = bytearray()
        pat_off, gp_off, qg_off, fixed = [0], [0], [0], []
        lut = {g: i for i, g in enumerate(self.fixed_groups)}
        for groups in groups_per_query:
            for g in groups:
                fixed.append(lut.get(frozenset(g), -1))
                for syn in sorted(g):
                    pat += str(syn).lower().encode("utf-8")
                    pat_off.append(len(pat))
                gp_off.append(len(pat_off) - 1)
            qg_off.append(len(gp_off) - 1)
        pat_arr = np.frombuffer(bytes(pat) + b"\0" * 16, dtype=np.uint8).copy()

        def dev(a, dt):
            return torch.from_numpy(np.asarray(a, dtype=dt)).to(self.device)
        self._qg_off = dev(qg_off, np.int32)
        return dev(pat_arr, np.uint8), dev(pat_off, np.int32), dev(gp_off, np.int32), fixed

    def factors(self, groups_per_query: Sequence[Sequence[Sequence[str]]], cand, penalty: float,
                want_hits: bool = False, use_bitmaps: bool = True):
        """gate float32[B, pool] (device) for candidate product rows int64[B, pool]."""
        cand = cand if isinstance(cand, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cand, dtype=np.int64))
        cand = cand.to(self.device, dtype=torch.int64).contiguous()
        B, pool = int(cand.shape[0]), int(cand.shape[1])
        if len(groups_per_query) != B:
            raise RRError("one group list per query")
        if any(len(g) > 32 for g in groups_per_query):
            raise RRError("at most 32 gate groups per query")
        pat, pat_off, gp_off, fixed = self._encode(groups_per_query)
        gate = torch.empty((B, pool), dtype=torch.float32, device=self.device)
        hits = torch.empty((B, pool), dtype=torch.int32, device=self.device) if want_hits else None
        bits = self.fixed_bits if use_bitmaps else None
        gfix = torch.from_numpy(np.asarray(fixed if fixed else [-1], dtype=np.int32)).to(self.device) if bits is not None else None
        check(self.lib.rr_gate_factors(_ptr(self.text), _ptr(self.text_off), _ptr(self.text_len), self.n_docs,
                                       _ptr(bits), _ptr(pat), _ptr(pat_off), _ptr(gp_off), _ptr(gfix),
                                       _ptr(self._qg_off), B, _ptr(cand), pool, float(penalty), _ptr(gate), _ptr(hits),
                                       self.device.index or 0, _stream()))
        return (gate, hits) if want_hits else gate


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
