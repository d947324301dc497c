#!/usr/bin/env python3
"""Generate tests/golden/* by running the REFERENCE's own code on seeded inputs.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What is executed from the reference, unmodified and read in place:
  * /root/reference/utils.py                      (imported as a module)
  * /root/reference/app/test.py                   (imported; `search(args)` run end to end with the
                                                   artifact paths pointed at a temp dir and the three
                                                   lazy model loaders replaced)
  * /root/reference/app/app_product_search.py     (imported under a stub `streamlit`; `run_search`
                                                   run with `_product_index`, `_st_encoder`,
                                                   `_bm25_loader`, `_cross_encoder` replaced)
`rank_bm25` is not available anywhere (see oracle/bm25_okapi.py), so the BM25 object handed to the
reference drivers is oracle.bm25_okapi.BM25Okapi; the goldens therefore pin everything around
`get_scores` (gather, min-max, prior, trust, blend, sort), not rank_bm25's arithmetic itself.

Outputs (small, committed):
  primitives.npz       inputs + outputs of utils.py / app/test.py primitives
  search_cases.npz     shared inputs of the end-to-end cases (embeddings, queries, corpus, metadata)
  search_cases.json    per case: parameters, reference top-k skus and component scores
  bm25_fixture.json    hand-checked BM25 known answers on the reference's fixture corpus
"""
from __future__ import annotations

import importlib.util
import io
import json
import os
import pickle
import sys
import tempfile
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle.bm25_okapi import BM25Okapi  # noqa: E402
import review_recommender_b200 as rr  # noqa: E402

synth = rr.synth


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------------------------
# stub streamlit: enough for app/app_product_search.py to import without running a search
# ----------------------------------------------------------------------------------------------
class _Any:
    """Absorbs any UI call.  Decorator factories return the function unchanged."""

    def __init__(self, n_iter: int = 0):
        self._n = n_iter

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Any):
            return a[0]                      # used as a bare decorator
        return _Any()

    def __getattr__(self, name):
        return _Any()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __bool__(self):
        return False

    def __iter__(self):
        return iter([_Any() for _ in range(self._n)])

    def __getitem__(self, i):
        return _Any()

    def strip(self):
        return ""


def _make_streamlit_stub():
    st = types.ModuleType("streamlit")

    def _cache(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    def _multi(spec, *a, **k):
        n = spec if isinstance(spec, int) else len(spec)
        return [_Any() for _ in range(n)]

    st.cache_resource = _cache
    st.cache_data = _cache
    st.query_params = {}
    st.tabs = _multi
    st.columns = _multi
    st.button = lambda *a, **k: False
    st.file_uploader = lambda *a, **k: None
    st.stop = lambda: None

    def _getattr(name):
        return _Any()
    st.__getattr__ = _getattr
    return st


def import_reference_modules(tmp: Path):
    os.environ["LOG_FILE"] = str(tmp / "logs" / "app.log")   # config.setup_logging writes here, not in /root/reference
    sys.path.insert(0, str(REF))
    ref_utils = _load("ref_utils", REF / "utils.py")
    ref_cli = _load("ref_cli", REF / "app" / "test.py")
    sys.modules["streamlit"] = _make_streamlit_stub()
    ref_st = _load("ref_streamlit_app", REF / "app" / "app_product_search.py")
    return ref_utils, ref_cli, ref_st


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def golden_primitives(ref_utils, ref_cli, ref_st):
    rng = np.random.default_rng(77)
    out = {}
    x = rng.standard_normal((7, 16)).astype(np.float32)
    x[3] = 0.0
    out["l2_in"] = x
    out["l2_out"] = ref_utils.l2_normalize(x)

    mm_cases = {
        "f32": rng.standard_normal(150).astype(np.float32),
        "f64": rng.standard_normal(150),
        "const": np.full(9, 3.25, dtype=np.float32),
        "nan": np.array([1.0, np.nan, 2.0], dtype=np.float32),
        "inf": np.array([1.0, np.inf, 2.0], dtype=np.float32),
        "tiny": np.array([1.0, 1.0 + 1e-13], dtype=np.float64),
        "one": np.array([0.5], dtype=np.float32),
        "bm25like": np.abs(rng.standard_normal(150)).astype(np.float32) * (rng.random(150) < 0.3),
    }
    for name, arr in mm_cases.items():
        arr = np.asarray(arr)
        out[f"mm_in_{name}"] = arr
        a = ref_utils.minmax_normalize(arr)
        b = ref_st._minmax(arr)
        c = ref_cli.minmax(arr)
        assert a.dtype == np.float32 and np.array_equal(a, b, equal_nan=True) and np.array_equal(a, c, equal_nan=True)
        out[f"mm_out_{name}"] = a

    avg = np.round(np.clip(rng.normal(4.1, 0.6, 150), 1, 5), 3)
    avg[[5, 17]] = np.nan
    n = np.clip(np.rint(rng.lognormal(np.log(12), 1.2, 150)), 0, 5000).astype(np.int64)
    n[3] = 0
    out["prior_avg"], out["prior_n"] = avg, n
    out["prior_out"] = ref_utils.bayesian_prior(avg, n, 20.0)
    assert np.array_equal(out["prior_out"], ref_st._bayes_prior(avg, n, 20.0), equal_nan=True)
    assert np.array_equal(out["prior_out"], ref_cli.bayesian_prior(avg, n, 20.0), equal_nan=True)
    out["prior_small"] = ref_utils.bayesian_prior(np.array([4.0, 3.0, np.nan]), np.array([0, 5, 10]))

    out["trust_n"] = n
    out["trust_out_50"] = ref_utils.trust_score_from_reviews(n, min_reviews=8)
    out["trust_out_80"] = ref_st._trust_from_reviews(n, min_reviews=8, sat=80)
    out["trust_out_m0"] = ref_st._trust_from_reviews(n.astype(np.float64), min_reviews=0, sat=80)

    mat = ref_utils.l2_normalize(rng.standard_normal((500, 48)).astype(np.float32))
    q = ref_utils.l2_normalize(rng.standard_normal((1, 48)).astype(np.float32))[0]
    idx, sims = ref_utils.cosine_similarity_search(q, mat, 25)
    idx2, sims2 = ref_cli.cosine_search(q, mat, 25)
    idx3, sims3 = ref_st._cosine_pool(q, mat, 25)
    assert np.array_equal(idx, idx2) and np.array_equal(idx, idx3) and np.array_equal(sims, sims3)
    out["cos_mat"], out["cos_q"], out["cos_idx"], out["cos_sims"] = mat, q, idx, sims
    idx_all, sims_all = ref_utils.cosine_similarity_search(q, mat[:10], 50)
    out["cos_idx_clamp"], out["cos_sims_clamp"] = idx_all, sims_all

    queries = ["best wireless headphones for music", "noise-cancelling headphones, really good!",
               "The cat's pajamas & an 80's-style yellow SOCK", "", "a an the"]
    toks = [ref_utils.tokenize_query(s) for s in queries]
    assert toks == [ref_cli.tokenize_query(s) for s in queries] == [ref_st._tokenize(s) for s in queries]
    np.savez_compressed(HERE / "primitives.npz", **out)
    return {"tokenize": {"queries": queries, "tokens": toks}}


# ----------------------------------------------------------------------------------------------
# end-to-end search cases through the reference drivers
# ----------------------------------------------------------------------------------------------
class _FakeEncoder:
    def __init__(self, table):
        self.table = table

    def encode(self, texts, normalize_embeddings=True):
        return np.stack([self.table[t] for t in texts])


def golden_search(ref_cli, ref_st, tmp: Path):
    N, D, V, B, L = 1200, 64, 600, 6, 4
    c = synth.make_corpus(N, D, V)
    qv = synth.queries(B, D)
    qt = synth.query_terms(B, L, c.doc_offsets, c.token_ids, V)
    corpus = synth.corpus_as_lists(c.doc_offsets, c.token_ids)
    skus = synth.skus(N)
    avg = c.avg_stars.copy()
    avg[::97] = np.nan                       # products without ratings
    nrev = c.n_reviews.astype(np.float64)
    nrev[::131] = np.nan                     # products without a count -> fillna(0)
    meta = pd.DataFrame({"sku": skus, "n_reviews": nrev, "avg_stars": avg,
                         "last_ts": 0, "agg_text": ["" for _ in range(N)]})
    # make a few exact dense ties and duplicate rows
    emb = c.emb.copy()
    emb[10] = emb[4]
    emb[700] = emb[4]

    query_strs = [" ".join(f"t{int(t) + 1}" for t in row) for row in qt]
    query_strs[3] = query_strs[3] + " " + query_strs[3].split()[0]      # duplicated query token
    query_strs[4] = "zzzunknownzzz " + query_strs[4]                      # unknown token
    query_strs[5] = "the of and"                                           # tokenises to nothing
    table = {s: qv[i] for i, s in enumerate(query_strs)}

    data = tmp / "data" / "processed"
    data.mkdir(parents=True, exist_ok=True)
    np.save(data / "product_emb.npy", emb)
    meta.to_parquet(data / "product_emb_meta.parquet")
    # BM25 blob in a DIFFERENT order than meta, as ensure_same_order expects to handle
    perm = np.random.default_rng(5).permutation(N)
    blob = {"skus": [skus[i] for i in perm], "corpus": [corpus[i] for i in perm], "tokenizer": "simple_en_v1"}
    with open(data / "product_bm25.pkl", "wb") as f:
        pickle.dump(blob, f, protocol=4)

    np.savez_compressed(HERE / "search_cases.npz", emb=emb, queries=qv, doc_offsets=c.doc_offsets,
                        token_ids=c.token_ids, n_reviews=nrev, avg_stars=avg, bm25_perm=perm)

    cases = []
    # ---- CLI driver ----
    ref_cli.P_EMB = data / "product_emb.npy"
    ref_cli.P_META = data / "product_emb_meta.parquet"
    ref_cli.BM25_PKL = data / "product_bm25.pkl"
    ref_cli.REV_EMB = data / "none.parquet"
    ref_cli._load_st_encoder = lambda: _FakeEncoder(table)
    ref_cli._load_rankbm25 = lambda: BM25Okapi

    def _no_ce():
        raise RuntimeError("cross-encoder unavailable in golden generation")
    ref_cli._load_cross_encoder = _no_ce

    cli_param_sets = [
        dict(k=10, rerank_k=0, w_dense=0.55, w_bm25=0.15, w_rerank=0.15, w_prior=0.10, w_best=0.05, prior_C=20.0),
        dict(k=25, rerank_k=50, w_dense=0.55, w_bm25=0.20, w_rerank=0.0, w_prior=0.20, w_best=0.0, prior_C=20.0),
        dict(k=120, rerank_k=0, w_dense=0.3, w_bm25=0.5, w_rerank=0.0, w_prior=0.2, w_best=0.0, prior_C=5.0),
    ]
    for pi, ps in enumerate(cli_param_sets):
        for qi, qs in enumerate(query_strs):
            jpath = tmp / f"cli_{pi}_{qi}.json"
            args = types.SimpleNamespace(query=qs, k=ps["k"], rerank_k=ps["rerank_k"], no_snippets=True,
                                         max_reviews_scan=0, w_dense=ps["w_dense"], w_bm25=ps["w_bm25"],
                                         w_rerank=ps["w_rerank"], w_prior=ps["w_prior"], w_best=ps["w_best"],
                                         prior_C=ps["prior_C"], gate_penalty=1.0, json_out=str(jpath))
            # capture the pre-rounding frame by wrapping minmax? simpler: recompute from printed JSON + capture `cand`
            captured = {}
            orig_sort = pd.DataFrame.sort_values

            def spy(self, *a, **k):
                if a and a[0] == "_final" or k.get("by") == "_final":
                    captured["pool"] = self.copy()
                return orig_sort(self, *a, **k)
            pd.DataFrame.sort_values = spy
            try:
                with redirect_stdout(io.StringIO()):
                    ref_cli.search(args)
            finally:
                pd.DataFrame.sort_values = orig_sort
            res = json.loads(jpath.read_text())["results"]
            pool = captured["pool"]
            cases.append({
                "driver": "cli", "query_index": qi, "query": qs, "params": ps,
                "top_skus": [r["sku"] for r in res],
                "top_scores_rounded": [r["score"] for r in res],
                "pool_skus": pool["sku"].astype(str).tolist(),
                "pool_final": [float(np.float32(v)) for v in pool["_final"].values],
                "pool_dense": [float(np.float32(v)) for v in pool["_dense"].values],
                "pool_bm25": [float(v) for v in pool["_bm25"].values],
                "pool_prior": [float(v) for v in pool["_prior"].values],
            })

    # ---- Streamlit driver ----
    Vn = ref_st._l2norm(np.array(emb), axis=1)
    meta_r = meta.reset_index(drop=True)
    bm25_blob = {"bm25": BM25Okapi(blob["corpus"]), "skus": [str(s) for s in blob["skus"]]}
    ref_st._product_index = lambda: (meta_r, Vn)
    ref_st._st_encoder = lambda name: _FakeEncoder(table)
    ref_st._bm25_loader = lambda: bm25_blob
    ref_st._cross_encoder = lambda name: None
    st_param_sets = [
        dict(k=10, rerank_k=0, w_dense=0.55, w_bm25=0.20, w_rerank=0.20, w_prior=0.20, w_best=0.10, prior_C=20.0, min_reviews=8),
        dict(k=100, rerank_k=0, w_dense=0.55, w_bm25=0.20, w_rerank=0.0, w_prior=0.20, w_best=0.0, prior_C=20.0, min_reviews=8),
        dict(k=10, rerank_k=50, w_dense=0.55, w_bm25=0.20, w_rerank=0.20, w_prior=0.20, w_best=0.10, prior_C=20.0, min_reviews=8),
        dict(k=200, rerank_k=0, w_dense=0.8, w_bm25=0.1, w_rerank=0.0, w_prior=0.1, w_best=0.0, prior_C=20.0, min_reviews=0),
    ]
    for pi, ps in enumerate(st_param_sets):
        for qi, qs in enumerate(query_strs):
            captured = {}
            orig_sort = pd.DataFrame.sort_values

            def spy(self, *a, **k):
                if a and a[0] == "_final":
                    captured["pool"] = self.copy()
                return orig_sort(self, *a, **k)
            pd.DataFrame.sort_values = spy
            try:
                top, snips, dbg = ref_st.run_search(qs, ps["k"], ps["rerank_k"], ps["w_dense"], ps["w_bm25"],
                                                    ps["w_rerank"], ps["w_prior"], ps["w_best"], ps["prior_C"],
                                                    False, 0, ps["min_reviews"], 1.0)
            finally:
                pd.DataFrame.sort_values = orig_sort
            pool = captured["pool"]
            cases.append({
                "driver": "streamlit", "query_index": qi, "query": qs, "params": ps,
                "pool_size": dbg["pool"], "tokens": dbg["tokens"],
                "top_skus": top["sku"].astype(str).tolist(),
                "top_final": [float(v) for v in top["_final"].values],
                "pool_skus": pool["sku"].astype(str).tolist(),
                "pool_final": [float(v) for v in pool["_final"].values],
                "pool_dense": [float(v) for v in pool["_dense"].values],
                "pool_bm25": [float(v) for v in pool["_bm25"].values],
                "pool_prior": [float(v) for v in pool["_prior"].values],
                "pool_trust": [float(v) for v in pool["_trust"].values],
            })
    return {"N": N, "D": D, "V": V, "query_strs": query_strs, "cases": cases}


def golden_bm25_fixture():
    """Known answers of SURVEY.md 8c (hand-checked) on the reference's fixture corpus
    tests/conftest.py:94-99 and on a negative-idf corpus."""
    return {
        "fixture_corpus": [["wireless", "headphones", "bluetooth"], ["yellow", "cat", "socks", "soft"],
                           ["gaming", "keyboard", "mechanical"]],
        "fixture_avgdl": 10 / 3,
        "fixture_idf": 0.5108256237659907,
        "fixture_scores": {
            "wireless headphones": [1.06979188, 0.0, 0.0],
            "cat": [0.0, 0.46864736, 0.0],
            "missing": [0.0, 0.0, 0.0],
        },
        "neg_corpus": [["a", "b"], ["a", "c"], ["a", "d"], ["e"]],
        "neg_idf_a": 0.12709467905808056,
        "neg_mean_idf": 0.5083787162323222,
        "neg_scores": {"a e": [0.11941782, 0.11941782, 0.11941782, 1.04974956]},
    }


def main():
    if not REF.exists():
        raise SystemExit("needs /root/reference (build container only)")
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        cwd = os.getcwd()
        os.chdir(tmp)                         # any relative path the reference touches lands in the temp dir
        try:
            ref_utils, ref_cli, ref_st = import_reference_modules(tmp)
            prim_meta = golden_primitives(ref_utils, ref_cli, ref_st)
            search = golden_search(ref_cli, ref_st, tmp)
        finally:
            os.chdir(cwd)
    (HERE / "search_cases.json").write_text(json.dumps({**search, **prim_meta}, indent=0))
    (HERE / "bm25_fixture.json").write_text(json.dumps(golden_bm25_fixture(), indent=1))
    print("wrote", [p.name for p in HERE.iterdir()])


if __name__ == "__main__":
    main()
