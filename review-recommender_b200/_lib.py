"""ctypes binding of librr_b200.so (the C ABI declared in include/rr_b200.h).

There is no fallback: if the shared library is missing or fails to load, importing the
compute modules raises.  Build it with `python __graft_entry__.py` (or
`python review-recommender_b200/build.py`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "librr_b200.so"

RR_DENSE_AUTO, RR_DENSE_EXACT, RR_DENSE_TENSOR = 0, 1, 2
PROF_CLASSES = ["bm25_tile", "bm25_cand", "dense_gemv", "select_rows", "tc_filter", "tc_select", "rescore",
                "tc_finalize", "fuse", "misc"]


class RRError(RuntimeError):
    pass


class IndexDesc(C.Structure):
    _fields_ = [
        ("n_docs", C.c_int64), ("row_offset", C.c_int64), ("dim", C.c_int32), ("dim_pad", C.c_int32),
        ("d_emb_f32", C.c_void_p), ("d_emb_bf16", C.c_void_p), ("max_row_norm", C.c_float),
        ("vocab_size", C.c_int32), ("tile_docs", C.c_int32), ("n_tiles", C.c_int32), ("n_freq", C.c_int32),
        ("d_postings", C.c_void_p), ("d_tile_base", C.c_void_p), ("d_dir", C.c_void_p),
        ("d_term_slot", C.c_void_p), ("d_rare_off", C.c_void_p),
        ("d_fwd_off", C.c_void_p), ("d_fwd_data", C.c_void_p),
        ("d_n_reviews", C.c_void_p), ("d_avg_stars", C.c_void_p),
    ]


class FusionParams(C.Structure):
    _fields_ = [
        ("w_dense", C.c_double), ("w_bm25", C.c_double), ("w_rerank", C.c_double), ("w_prior", C.c_double),
        ("w_best", C.c_double), ("prior_C", C.c_double), ("min_reviews", C.c_int32), ("saturation", C.c_int32),
        ("use_trust", C.c_int32), ("rerank_is_f32", C.c_int32), ("bm25_is_f64_zero", C.c_int32),
        ("k", C.c_int32), ("pool", C.c_int32), ("best_is_raw", C.c_int32),
    ]


class DenseStats(C.Structure):
    _fields_ = [("path", C.c_int32), ("n_uncertified", C.c_int32), ("n_overflow", C.c_int32),
                ("shortlist", C.c_int32), ("n_segments", C.c_int32), ("eps", C.c_float)]


# name -> (restype, argtypes); every symbol include/rr_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "rr_last_error": (C.c_char_p, []),
    "rr_abi_version": (C.c_int, []),
    "rr_bm25_local_stats": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P]),
    "rr_bm25_idf": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_double, _P, _P]),
    "rr_bm25_build_postings": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, C.c_double, C.c_double, C.c_double,
                                         C.c_int32, C.c_int32, C.POINTER(_P)]),
    "rr_postings_nnz": (C.c_int64, [_P]),
    "rr_postings_n_tiles": (C.c_int32, [_P]),
    "rr_postings_data": (_P, [_P]),
    "rr_postings_tile_base": (_P, [_P]),
    "rr_postings_n_freq": (C.c_int32, [_P]),
    "rr_postings_dir": (_P, [_P]),
    "rr_postings_term_slot": (_P, [_P]),
    "rr_postings_rare_off": (_P, [_P]),
    "rr_bm25_dir_threshold": (C.c_int32, [C.c_int32]),
    "rr_postings_fwd_off": (_P, [_P]),
    "rr_postings_fwd_data": (_P, [_P]),
    "rr_postings_free": (None, [_P]),
    "rr_index_create": (C.c_int, [C.POINTER(_P), C.POINTER(IndexDesc), C.c_int]),
    "rr_index_destroy": (None, [_P]),
    "rr_bm25_get_scores": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "rr_bm25_candidates": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "rr_dense_topk": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "rr_dense_debug_bf16_scores": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P]),
    "rr_candidate_tuples": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P]),
    "rr_fuse_topk": (C.c_int, [C.POINTER(FusionParams), C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, _P, _P, _P, C.c_int, _P]),
    "rr_fuse_topk_sharded": (C.c_int, [C.POINTER(FusionParams), C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P,
                                       _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "rr_dense_topk_deferred": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "rr_struct_sizes": (None, [_P]),
    "rr_profile_enable": (C.c_int, [C.c_int]),
    "rr_profile_collect": (C.c_int, [_P, _P, C.c_int32]),
    "rr_best_review_scores": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P,
                                        C.c_int, _P]),
    "rr_bm25_gpu_build_begin": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int, _P]),
    "rr_bm25_gpu_build_stats": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "rr_bm25_gpu_build_finish": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rr_bm25_gpu_build_free": (None, [_P]),
    "rr_normalize_rows": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P, C.c_int, _P]),
    "rr_max_row_norm": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int, _P]),
    "rr_bf16_rows": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int, _P]),
    "rr_gate_factors": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_int32, C.c_double,
                                  _P, _P, C.c_int, _P]),
    "rr_gate_fixed_bitmaps": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, C.c_int32, _P, C.c_int, _P]),
    "rr_shard_tuples": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "rr_hybrid_search": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.POINTER(FusionParams), C.c_int32,
                                   _P, _P, _P]),
    "rr_hybrid_search_deferred": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.POINTER(FusionParams), C.c_int32,
                                            _P, _P, _P, _P]),
    "rr_hybrid_search_host": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.POINTER(FusionParams), C.c_int32,
                                        _P, _P, _P]),
    "rr_launch_count": (C.c_int64, [C.c_int]),
    "rr_dense_last_stats": (C.c_int, [_P, C.POINTER(DenseStats)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises RRError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RRError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_int32 * 3)()
    lib.rr_struct_sizes(sizes)
    mine = (C.sizeof(IndexDesc), C.sizeof(FusionParams), C.sizeof(DenseStats))
    if tuple(sizes) != mine:
        raise RRError(f"struct layout mismatch between _lib.py {mine} and {LIB_PATH.name} {tuple(sizes)}: rebuild the library")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().rr_last_error()
        raise RRError(f"librr_b200 error {rc}: {msg.decode() if msg else ''}")
