"""Live check of the oracle restatements against the reference's own `utils.py`, on many random inputs.
Runs only where /root/reference exists (the build container); the GPU box relies on the committed goldens."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import gates as G
from oracle import primitives as P

REF = Path("/root/reference/utils.py")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="needs /root/reference (build container only)")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("ref_utils_live", str(REF))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_utils_live"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_numeric_primitives_on_random_inputs(ref):
    rng = np.random.default_rng(2024)
    for trial in range(60):
        n = int(rng.integers(1, 400))
        d = int(rng.choice([1, 3, 8, 17, 64, 130, 384, 768]))
        x = (rng.standard_normal((min(n, 40), d)) * rng.uniform(1e-3, 50)).astype(np.float32)
        if trial % 7 == 0:
            x[0] = 0
        np.testing.assert_array_equal(P.l2_normalize(x), ref.l2_normalize(x))
        for arr in (rng.standard_normal(n).astype(np.float32), rng.standard_normal(n),
                    np.abs(rng.standard_normal(n)).astype(np.float32) * (rng.random(n) < 0.3),
                    np.full(n, 2.5, dtype=np.float32)):
            if trial % 11 == 0 and n > 2:
                arr = arr.copy()
                arr[1] = np.nan
            a, b = P.minmax_normalize(arr), ref.minmax_normalize(arr)
            assert a.dtype == b.dtype
            np.testing.assert_array_equal(a, b)
        avg = np.round(np.clip(rng.normal(4.1, 0.6, n), 1, 5), 3)
        avg[rng.random(n) < 0.05] = np.nan
        cnt = np.clip(np.rint(rng.lognormal(np.log(12), 1.2, n)), 0, 5000).astype(np.int64)
        C = float(rng.choice([5.0, 20.0, 50.0]))
        np.testing.assert_array_equal(P.bayesian_prior(avg, cnt, C), ref.bayesian_prior(avg, cnt, C))
        mr = int(rng.choice([0, 1, 8, 20]))
        np.testing.assert_array_equal(P.trust_score_from_reviews(cnt, mr, 50), ref.trust_score_from_reviews(cnt, mr))
        np.testing.assert_array_equal(P.trust_score_from_reviews(cnt.astype(np.float64), mr, 80),
                                      ref.trust_score_from_reviews(cnt.astype(np.float64), mr, 80))


def test_cosine_search_on_random_inputs(ref):
    rng = np.random.default_rng(7)
    for n, d, k in ((50, 8, 10), (700, 48, 25), (3000, 384, 150), (20, 16, 100)):
        mat = ref.l2_normalize(rng.standard_normal((n, d)).astype(np.float32))
        q = ref.l2_normalize(rng.standard_normal((1, d)).astype(np.float32))[0]
        i0, s0 = ref.cosine_similarity_search(q, mat, k)
        i1, s1 = P.cosine_similarity_search(q, mat, k)
        np.testing.assert_array_equal(i0, i1)
        np.testing.assert_array_equal(s0, s1)


def test_tokenizer_and_gates_on_generated_queries(ref):
    rng = np.random.default_rng(5)
    words = ["yellow", "Golden", "cat's", "wireless", "headphones", "anc", "noise-canceling", "USB-C", "the", "of", "a",
             "keyboard", "Mechanical", "rose", "tan", "42", "x", "sock", "socks", "café", "design", "dog", "puppies"]
    texts = ["", "A yellow cat sock with noise cancelling", "GOLDEN RETRIEVER PUPPIES", "mechanical keyboards, wireless",
             "plain text without anything", "rose-tinted café design pattern 42"]
    for _ in range(200):
        q = " ".join(rng.choice(words, size=int(rng.integers(0, 9))))
        assert P.tokenize_query(q) == ref.tokenize_query(q)
        g0, g1 = ref.build_gate_groups(q), G.build_gate_groups(q)
        assert g0 == g1
        for t in texts:
            for pen in (0.5, 0.25, 1.0):
                assert G.calculate_gate_factor(t, g1, pen) == ref.calculate_gate_factor(t, g0, pen)
