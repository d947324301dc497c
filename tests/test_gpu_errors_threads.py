"""Error behaviour and re-entrancy of the C-ABI: every call returns a code and a thread-local message
(the Python layer raises RRError), nothing aborts or falls back; one index handle is called concurrently from
several Python threads the way Streamlit sessions share a cached resource (app/app_product_search.py:53,71,119)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rr():
    import review_recommender_b200 as rr
    return rr


def test_bad_arguments_raise_with_a_message():
    import torch
    rr = _rr()
    c = rr.synth.make_corpus(3000, 64, 400)
    ix = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, 400, c.n_reviews, c.avg_stars, make_bf16=False)
    q = rr.synth.queries(4, 64)
    with pytest.raises(rr._lib.RRError, match="dim"):
        ix.dense_topk(rr.synth.queries(2, 32), 10)
    with pytest.raises(rr._lib.RRError, match="8192"):
        ix.dense_topk(q, 9000)
    with pytest.raises(rr._lib.RRError, match="bf16"):
        ix.dense_topk(q, 10, rr._lib.RR_DENSE_TENSOR)              # tensor path asked for, no bf16 copy uploaded
    with pytest.raises(rr._lib.RRError):
        rr.engine.GateIndex(["a"], [["x"]] * 33)                    # more than 32 fixed groups
    # the handle is still usable after errors
    idx, sims, cnt = ix.dense_topk(q, 10)
    assert int(cnt.min()) == 10 and torch.isfinite(sims).all()
    # unknown / out-of-range term ids contribute nothing; an empty batch is a no-op
    ids = np.array([[-1, 400, 10**6, 5]], dtype=np.int32)
    s = ix.bm25_get_scores(ids, np.array([4], dtype=np.int32))
    s5 = ix.bm25_get_scores(np.array([[5]], dtype=np.int32), np.array([1], dtype=np.int32))
    assert torch.equal(s, s5)
    assert ix.bm25_get_scores(np.zeros((0, 1), np.int32), np.zeros(0, np.int32)).shape[0] == 0
    ix.close()


def test_concurrent_calls_on_one_handle():
    import torch
    rr = _rr()
    n, d, v = 80_000, 128, 3000
    c = rr.synth.make_corpus(n, d, v)
    ix = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, v, c.n_reviews, c.avg_stars)
    fusion = rr.engine.Fusion(k=20, rerank_k=0, w_rerank=0.0, w_best=0.0)
    batches = []
    for t in range(6):
        b = 40 + 8 * t
        q = np.roll(rr.synth.queries(b, d), t, axis=0)
        qt = rr.synth.query_terms(b, 4, c.doc_offsets, c.token_ids, v).astype(np.int32)
        batches.append((q, qt, np.full(b, 4, dtype=np.int32)))
    serial = [ix.hybrid_search_host(*bt, fusion) for bt in batches]
    out, errs = [None] * len(batches), []

    def work(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(3):
                    out[i] = ix.hybrid_search_host(*batches[i], fusion)
        except Exception as e:                                       # pragma: no cover
            errs.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for (r0, f0), (r1, f1) in zip(serial, out):
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(f0, f1)

    # device-pointer entry point (no stream sync inside): calls from different threads on different streams share the
    # handle's scratch buffers and must still not interfere (cross-stream fence inside the library)
    dev_out = [None] * len(batches)

    def work_dev(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                q, qt, nt = (torch.from_numpy(x).cuda() for x in batches[i])
                for _ in range(5):
                    r, f = ix.hybrid_search(q, qt, nt, fusion, mode=rr._lib.RR_DENSE_EXACT if i % 2 else rr._lib.RR_DENSE_TENSOR)
                st.synchronize()
                dev_out[i] = (r.cpu().numpy(), f.cpu().numpy())
        except Exception as e:                                       # pragma: no cover
            errs.append(e)
    threads = [threading.Thread(target=work_dev, args=(i,)) for i in range(len(batches))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for (r0, f0), (r1, f1) in zip(serial, dev_out):
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(f0, f1)
    ix.close()


def test_batches_in_flight_on_two_handles_equal_the_synchronous_search():
    """hybrid_search_begin / PendingSearch.result on two handles (view()) and two streams over the same index
    tensors; a query the tensor path cannot certify (600 near-duplicates of its target) is repeated exactly."""
    import torch
    rr = _rr()
    n, d, v = 90_000, 128, 3000
    c = rr.synth.make_corpus(n, d, v)
    rng = np.random.default_rng(3)
    base = c.emb[17].copy()
    for r in range(1000, 1600):
        x = base + 1e-4 * rng.standard_normal(d).astype(np.float32)
        c.emb[r] = x / np.linalg.norm(x)
    ix = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, v, c.n_reviews, c.avg_stars)
    lanes = [(ix, torch.cuda.Stream()), (ix.view(), torch.cuda.Stream())]
    fusion = rr.engine.Fusion(k=50, rerank_k=0, w_rerank=0.0, w_best=0.0)
    batches = []
    for t in range(5):
        q = np.roll(rr.synth.queries(64, d), 3 * t, axis=0)
        if t == 2:
            q[5] = base
        qt = np.roll(rr.synth.query_terms(64, 4, c.doc_offsets, c.token_ids, v).astype(np.int32), t, axis=0)
        batches.append(tuple(torch.from_numpy(x).cuda() for x in (q, qt, np.full(64, 4, dtype=np.int32))))
    want = [ix.hybrid_search(*b, fusion, mode=rr._lib.RR_DENSE_EXACT) for b in batches]
    torch.cuda.synchronize()
    tokens = []
    for i, b in enumerate(batches):
        h, st = lanes[i % 2]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            tokens.append(h.hybrid_search_begin(*b, fusion, mode=rr._lib.RR_DENSE_TENSOR))
    got = [t.result() for t in tokens]
    assert tokens[2].repeated >= 1 and tokens[0].repeated == 0
    for (r0, f0), (r1, f1) in zip(want, got):
        assert torch.equal(r0, r1) and torch.equal(f0, f1)
    lanes[1][0].close()
    ix.close()


def test_small_host_batches_replay_a_graph_with_fresh_inputs():
    """rr_hybrid_search_host, B <= 8 on the exact path: the second call with a shape captures a CUDA graph and later calls
    replay it.  Every call must see ITS inputs (the graph reads the handle's staging buffers) and return what the plain
    path returns; interleaving shapes (B = 1, 3, 1) and fusion parameters keeps separate graphs."""
    import os
    rr = _rr()
    n, d, v = 6000, 64, 900
    c = rr.synth.make_corpus(n, d, v)
    q = rr.synth.queries(40, d)
    qt = rr.synth.query_terms(40, 4, c.doc_offsets, c.token_ids, v).astype(np.int32)
    fa = rr.engine.Fusion(k=10, rerank_k=0, w_rerank=0.0, w_best=0.0)
    fb = rr.engine.Fusion(k=7, rerank_k=0, w_dense=0.3, w_bm25=0.5, w_rerank=0.0, w_best=0.0, driver="cli")
    plan = [(i, 1, fa) for i in range(6)] + [(6, 3, fa), (9, 1, fb), (10, 3, fa), (13, 1, fb), (14, 1, fa), (15, 3, fb), (18, 3, fb)]
    results = {}
    for graphs in (True, False):
        if graphs:
            os.environ.pop("RR_NO_GRAPHS", None)
        else:
            os.environ["RR_NO_GRAPHS"] = "1"
        try:
            ix = rr.engine.HybridIndex(c.emb, c.doc_offsets, c.token_ids, v, c.n_reviews, c.avg_stars)
            out = []
            for i0, b, f in plan:
                rr.engine.launch_count(reset=True)
                r, s = ix.hybrid_search_host(q[i0:i0 + b], qt[i0:i0 + b], np.full(b, 4, dtype=np.int32), f)
                out.append((r.copy(), s.copy(), rr.engine.launch_count()))
            results[graphs] = out
            ix.close()
        finally:
            os.environ.pop("RR_NO_GRAPHS", None)
    for (r1, s1, l1), (r0, s0, l0) in zip(results[True], results[False]):
        np.testing.assert_array_equal(r1, r0)
        np.testing.assert_array_equal(s1, s0)
    assert results[True][5][2] == 1 and results[False][5][2] > 1, "the sixth identical-shape call must be one graph launch"


def test_tensor_path_slices_very_large_batches(monkeypatch):
    """AUTO / TENSOR mode must not reject a batch that exceeds what one tensor-path call takes: it goes through
    in slices (RR_TC_MAX_BATCH shrinks the slice so that a 600-query batch needs three)."""
    monkeypatch.setenv("RR_TC_MAX_BATCH", "256")
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import numpy as np, review_recommender_b200 as rr
        emb = rr.synth.embeddings(70_000, 128); q = rr.synth.queries(600, 128)
        ix = rr.engine.HybridIndex(emb)
        i2, s2, c2 = ix.dense_topk(q, 50, rr._lib.RR_DENSE_TENSOR)
        assert ix.dense_stats()["path"] == 2
        i1, s1, c1 = ix.dense_topk(q, 50, rr._lib.RR_DENSE_EXACT)
        assert (i1 == i2).all() and (s1 == s2).all() and (c1 == c2).all()
        print("sliced ok")
    """)
    import os
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ), capture_output=True, text=True, timeout=300,
                       cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
    assert r.returncode == 0 and "sliced ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
