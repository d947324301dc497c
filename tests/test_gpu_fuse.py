"""K4 parity: normalisation + prior + trust + blend + top-k against the reference's own outputs
(golden cases captured from run_search / search) and against the oracle on synthetic pools."""
import json

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

from oracle import primitives as P
from oracle.bm25_okapi import BM25Okapi
from oracle.hybrid import cli_search_core, run_search_core
from tests.parity import FUSED_RTOL, assert_ids_match_modulo_ties


def _rr():
    import review_recommender_b200 as rr
    return rr


@pytest.fixture(scope="module")
def world(golden_dir):
    rr = _rr()
    cases = json.loads((golden_dir / "search_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    n = z["emb"].shape[0]
    skus = rr.synth.skus(n)
    corpus = rr.synth.corpus_as_lists(z["doc_offsets"], z["token_ids"])
    perm = z["bm25_perm"]
    meta = pd.DataFrame({"sku": skus, "n_reviews": z["n_reviews"], "avg_stars": z["avg_stars"],
                         "agg_text": ["" for _ in range(n)]})
    bm25 = BM25Okapi([corpus[i] for i in perm])
    Vn = P.l2_normalize(np.array(z["emb"]), axis=1)
    return dict(cases=cases, z=z, meta=meta, bm25=bm25, bm25_skus=[skus[i] for i in perm], Vn=Vn, skus=skus)


def test_fuse_reproduces_reference_cases_from_their_raw_pools(world):
    """Feed K4 the raw pool tuples the reference saw (dense_raw, bm25_raw, n, avg) and require the
    fused scores and the ranking the reference produced."""
    rr = _rr()
    z = world["z"]
    ix = rr.engine.HybridIndex(world["Vn"], device="cuda:0", make_bf16=False)
    n_bit_exact = n_total = 0
    for c in world["cases"]["cases"]:
        ps = dict(c["params"])
        q = z["queries"][c["query_index"]]
        toks = P.tokenize_query(c["query"])
        if c["driver"] == "cli":
            top, pool = cli_search_core(q, world["Vn"], world["meta"], world["bm25"], world["bm25_skus"], toks, **ps)
            fusion = rr.engine.Fusion(driver="cli", **ps)
        else:
            top, pool = run_search_core(q, world["Vn"], world["meta"], world["bm25"], world["bm25_skus"], toks, **ps)
            fusion = rr.engine.Fusion(driver="streamlit", **ps)
        assert pool["sku"].tolist() == c["pool_skus"]
        Pn = len(pool)
        assert Pn == fusion.pool
        import torch
        dev = ix.device
        dense = torch.from_numpy(pool["_dense_raw"].values.astype(np.float32))[None].to(dev)
        bm25 = torch.from_numpy(pool["_bm25_raw"].values.astype(np.float32))[None].to(dev)
        nrev = torch.from_numpy(np.nan_to_num(world["meta"]["n_reviews"].values[pool["_row"].values], nan=0.0))[None].to(dev)
        avg = torch.from_numpy(world["meta"]["avg_stars"].values[pool["_row"].values].astype(np.float64))[None].to(dev)
        grow = torch.from_numpy(pool["_row"].values.astype(np.int64))[None].to(dev)
        rerank = torch.zeros((1, Pn), dtype=torch.float32, device=dev) if fusion.rerank_k > 0 else None
        rows, final, pos, comp = ix.fuse(fusion, dense, bm25, nrev, avg, grow, rerank=rerank, want_components=True)
        comp = comp.cpu().numpy()[0]
        ref_final = np.float32(c["pool_final"])
        np.testing.assert_allclose(comp[:, 4], ref_final, rtol=FUSED_RTOL, atol=1e-7)
        np.testing.assert_array_equal(comp[:, 0], np.float32(c["pool_dense"]))
        np.testing.assert_array_equal(comp[:, 1], np.float32(c["pool_bm25"]))
        np.testing.assert_allclose(comp[:, 2], np.float64(c["pool_prior"]), rtol=1e-6, atol=1e-7)
        if c["driver"] == "streamlit":
            np.testing.assert_array_equal(comp[:, 3], np.float32(c["pool_trust"]))
        n_bit_exact += int(np.sum(comp[:, 4] == ref_final))
        n_total += Pn
        k = fusion.k
        sku_of = world["skus"]
        got_skus = [sku_of[r] for r in rows.cpu().numpy()[0][:len(c["top_skus"])]]
        ref_rows = [int(s[3:]) for s in c["top_skus"]]
        ref_top_final = np.sort(ref_final)[::-1][:len(ref_rows)]
        assert_ids_match_modulo_ties(rows.cpu().numpy()[0][:len(ref_rows)], final.cpu().numpy()[0][:len(ref_rows)],
                                     ref_rows, ref_top_final, 1e-7, f"{c['driver']} q{c['query_index']}")
    print(f"fused scores bit-identical to the reference: {n_bit_exact}/{n_total}")
    assert n_bit_exact >= 0.999 * n_total
    ix.close()


@pytest.mark.parametrize("pool,k,n_in", [(150, 100, 150), (150, 10, 1200), (1000, 1000, 1000), (100, 100, 40)])
def test_fuse_random_pools_and_cross_shard_merge(pool, k, n_in):
    """Random tuples incl. NaN ratings; n_in > pool exercises the (dense desc, row asc) merge."""
    rr = _rr()
    import torch
    rng = np.random.default_rng(pool * 7 + n_in)
    B = 5
    ix = rr.engine.HybridIndex(np.eye(4, dtype=np.float32), device="cuda:0", make_bf16=False)
    dense = rng.standard_normal((B, n_in)).astype(np.float32) * 0.05
    dense[:, 3] = dense[:, 5]                                      # exact dense ties
    bm25 = (np.abs(rng.standard_normal((B, n_in))) * (rng.random((B, n_in)) < 0.4)).astype(np.float32)
    nrev = np.clip(np.rint(rng.lognormal(np.log(12), 1.2, (B, n_in))), 0, 5000)
    avg = np.round(np.clip(rng.normal(4.1, 0.6, (B, n_in)), 1, 5), 3)
    avg[rng.random((B, n_in)) < 0.05] = np.nan
    avg[4] = np.nan                                                # a pool without any rating
    bm25[3] = 0.0                                                  # a query without BM25 signal
    grow = np.stack([rng.permutation(10 * n_in)[:n_in] for _ in range(B)]).astype(np.int64)
    fusion = rr.engine.Fusion(k=k, rerank_k=0, driver="streamlit")
    fusion_pool = max(k, 150)
    assert fusion.pool == fusion_pool
    if pool != fusion_pool:
        fusion = rr.engine.Fusion(k=k, rerank_k=pool if pool > 150 else 0, driver="cli" if pool == 100 else "streamlit")
    dev = ix.device
    t = lambda a: torch.from_numpy(a).to(dev)
    rows, final, pos, _ = ix.fuse(fusion, t(dense), t(bm25), t(nrev), t(avg), t(grow))
    rows, final = rows.cpu().numpy(), final.cpu().numpy()
    for b in range(B):
        order = np.lexsort((grow[b], -dense[b].astype(np.float64)))[:fusion.pool]
        dn = P.minmax_normalize(dense[b][order])
        bn = P.minmax_normalize(bm25[b][order])
        n_, a_ = nrev[b][order], avg[b][order]
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pr = P.bayesian_prior(a_, n_, C=fusion.prior_C)
        vol = np.log1p(n_) / (np.log1p(n_).max() + 1e-9)
        prior = P.minmax_normalize(pr) * 0.7 + 0.3 * vol
        z = np.zeros(len(order), dtype=np.float32)
        rer = z if fusion.rerank_k > 0 else 0.0
        fin = (fusion.w_dense * dn + fusion.w_bm25 * bn + fusion.w_rerank * rer + fusion.w_prior * prior +
               fusion.w_best * z).astype(np.float32)
        if fusion.driver == "streamlit":
            fin = fin * P.trust_score_from_reviews(n_, fusion.min_reviews, 80)
        fin = fin * np.ones(len(order), dtype=np.float32)
        ref_order = np.lexsort((np.arange(len(order)), -fin.astype(np.float64)))[:k]
        kk = len(ref_order)
        np.testing.assert_allclose(final[b, :kk], fin[ref_order], rtol=FUSED_RTOL, atol=1e-7)
        assert_ids_match_modulo_ties(rows[b, :kk], final[b, :kk], grow[b][order][ref_order], fin[ref_order], 1e-6, f"b{b}")
        assert np.all(rows[b, kk:] == -1)
    ix.close()
