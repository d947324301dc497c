"""Index preparation on the GPU (rr_normalize_rows): bit-identical to the reference's l2_normalize as captured in
tests/golden/primitives.npz, and to NumPy on larger random matrices; the fused bf16 copy is the round-to-nearest
copy of the normalised rows; HybridIndex(normalize=True) equals HybridIndex over host-normalised rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import primitives as P


def _normalize(x, want_bf16=False):
    import ctypes as C
    import torch
    import review_recommender_b200 as rr
    lib = rr._lib.load()
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    n, d = t.shape
    out = torch.empty_like(t)
    norms = torch.empty(n, dtype=torch.float32, device="cuda")
    dim_pad = (d + 63) // 64 * 64
    bf = torch.empty((n, dim_pad), dtype=torch.bfloat16, device="cuda") if want_bf16 else None
    rr._lib.check(lib.rr_normalize_rows(C.c_void_p(t.data_ptr()), n, d, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(bf.data_ptr() if bf is not None else 0), dim_pad,
                                        C.c_void_p(norms.data_ptr()), 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out.cpu().numpy(), norms.cpu().numpy(), bf


@pytest.mark.parametrize("D", [7, 100, 130, 384, 768])
def test_reference_golden_rows(golden_dir, D):
    prim = np.load(golden_dir / "primitives.npz")
    got, norms, _ = _normalize(prim[f"l2big_in_{D}"])
    np.testing.assert_array_equal(got, prim[f"l2big_out_{D}"])
    np.testing.assert_array_equal(norms, np.linalg.norm(prim[f"l2big_in_{D}"], axis=1))


def test_small_golden_and_random_matrices(golden_dir):
    import torch
    prim = np.load(golden_dir / "primitives.npz")
    got, _, _ = _normalize(prim["l2_in"])
    np.testing.assert_array_equal(got, prim["l2_out"])
    rng = np.random.default_rng(5)
    for n, d in ((1, 1), (33, 8), (1000, 129), (20_000, 384), (3000, 1000), (500, 2049)):
        x = (rng.standard_normal((n, d)) * rng.uniform(0.1, 10, size=(n, 1))).astype(np.float32)
        got, norms, bf = _normalize(x, want_bf16=True)
        want = P.l2_normalize(x, axis=1)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(norms, np.linalg.norm(x, axis=1))
        ref_bf = torch.from_numpy(want).to(torch.bfloat16)
        assert torch.equal(bf[:, :d].cpu(), ref_bf) and bool((bf[:, d:] == 0).all())


def test_index_normalize_flag_equals_host_normalisation():
    import review_recommender_b200 as rr
    rng = np.random.default_rng(6)
    raw = (rng.standard_normal((70_000, 384)) * 3.0).astype(np.float32)
    q = rr.synth.queries(40, 384)
    a = rr.engine.HybridIndex(raw, device="cuda:0", normalize=True)
    b = rr.engine.HybridIndex(P.l2_normalize(raw, axis=1), device="cuda:0")
    assert (a.emb == b.emb).all() and (a.emb_bf16 == b.emb_bf16).all()
    for mode in (rr._lib.RR_DENSE_TENSOR, rr._lib.RR_DENSE_EXACT):
        ia, sa, _ = a.dense_topk(q, 150, mode)
        ib, sb, _ = b.dense_topk(q, 150, mode)
        assert (ia == ib).all() and (sa == sb).all()
    a.close()
    b.close()
