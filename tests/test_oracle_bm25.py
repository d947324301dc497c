"""Oracle BM25 (restated rank_bm25.BM25Okapi) against the known answers of SURVEY.md 8c and
against its own vectorised twin.  CPU only."""
import json
import math

import numpy as np
import pytest

from oracle.bm25_okapi import BM25Okapi, BM25OkapiCSR, flatten_corpus


@pytest.fixture(scope="module")
def fx(golden_dir):
    return json.loads((golden_dir / "bm25_fixture.json").read_text())


def test_fixture_corpus_known_answers(fx):
    bm = BM25Okapi(fx["fixture_corpus"])          # corpus of the reference's tests/conftest.py:94-99
    assert bm.avgdl == fx["fixture_avgdl"]
    assert all(v == fx["fixture_idf"] for v in bm.idf.values())
    assert fx["fixture_idf"] == math.log(2.5) - math.log(1.5)
    for q, want in fx["fixture_scores"].items():
        got = bm.get_scores(q.split())
        assert got.dtype == np.float64
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-9)


def test_negative_idf_floor(fx):
    bm = BM25Okapi(fx["neg_corpus"])
    assert bm.average_idf == fx["neg_mean_idf"]
    assert bm.idf["a"] == fx["neg_idf_a"]
    np.testing.assert_allclose(bm.get_scores(["a", "e"]), fx["neg_scores"]["a e"], rtol=0, atol=5e-9)


def test_duplicate_and_unknown_tokens():
    corpus = [["x", "y", "x"], ["y", "z"], ["w"], ["x"]]
    bm = BM25Okapi(corpus)
    one = bm.get_scores(["x"])
    np.testing.assert_array_equal(bm.get_scores(["x", "x"]), one + one)
    np.testing.assert_array_equal(bm.get_scores(["nope"]), np.zeros(4))
    np.testing.assert_array_equal(bm.get_scores([]), np.zeros(4))


@pytest.mark.parametrize("seed,n,v", [(0, 50, 12), (1, 400, 60), (2, 1500, 300)])
def test_csr_twin_is_bit_identical(seed, n, v):
    rng = np.random.default_rng(seed)
    corpus = [[f"w{int(t)}" for t in rng.zipf(1.3, size=int(rng.integers(1, 30))) % v] for _ in range(n)]
    bm = BM25Okapi(corpus)
    offs, ids, vocab = flatten_corpus(corpus)
    csr = BM25OkapiCSR(offs, ids, len(vocab))
    assert csr.avgdl == bm.avgdl and csr.average_idf == bm.average_idf
    for t, w in enumerate(vocab):
        assert csr.idf[t] == bm.idf[w]
    for _ in range(10):
        q = [int(x) for x in rng.integers(0, len(vocab), size=int(rng.integers(1, 8)))]
        np.testing.assert_array_equal(csr.get_scores(q), bm.get_scores([vocab[i] for i in q]))


def test_pin_against_the_real_rank_bm25_when_installed(fx):
    """Flips oracle.bm25_okapi from "parity unpinned" to pinned wherever the real package exists (it is absent from
    the reference tree, its requirements.txt and this image): bit-for-bit on the reference's fixture corpus, the
    negative-idf corpus and random Zipf corpora -- construction statistics and get_scores."""
    rank_bm25 = pytest.importorskip("rank_bm25")
    from pathlib import Path
    if Path(rank_bm25.__file__).resolve().is_relative_to(Path(__file__).resolve().parent.parent):
        pytest.skip("`rank_bm25` resolves to this repo's drop-in shim, not the third-party package")
    rng = np.random.default_rng(123)
    corpora = [fx["fixture_corpus"], fx["neg_corpus"]]
    for n, v in ((60, 15), (800, 120)):
        corpora.append([[f"w{int(t)}" for t in rng.zipf(1.3, size=int(rng.integers(1, 40))) % v] for _ in range(n)])
    for corpus in corpora:
        real, mine = rank_bm25.BM25Okapi(corpus), BM25Okapi(corpus)
        assert real.avgdl == mine.avgdl and real.average_idf == mine.average_idf and real.corpus_size == mine.corpus_size
        assert list(real.idf.items()) == list(mine.idf.items())                  # values AND dict order
        vocab = list(mine.idf)
        for _ in range(20):
            q = [vocab[int(i)] for i in rng.integers(0, len(vocab), size=int(rng.integers(1, 9)))] + ["<unknown>"]
            np.testing.assert_array_equal(real.get_scores(q), mine.get_scores(q))
