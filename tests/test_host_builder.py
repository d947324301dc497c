"""Host-side code of the C-ABI library, checked on the CPU: every declared symbol is exported,
and the BM25 index builder (rr_bm25_local_stats / rr_bm25_idf / rr_bm25_build_postings) agrees
bit for bit with the oracle's float64 statistics and with fp32(oracle impact)."""
import re
from pathlib import Path

import numpy as np
import pytest

import review_recommender_b200 as rr
from oracle.bm25_okapi import BM25OkapiCSR

REPO = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from __graft_entry__ import build
    build()
    return rr._lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = (REPO / "include" / "rr_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # drop comments
    declared = set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(rr._lib.SIGNATURES), declared ^ set(rr._lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rr_abi_version() == 3


@pytest.mark.parametrize("n,v,tile", [(37, 50, 16), (3000, 400, 256), (5000, 2000, 16384), (9000, 5000, 64)])
def test_builder_matches_oracle(lib, n, v, tile):
    offs, toks = rr.synth.corpus_tokens(n, v)
    csr = BM25OkapiCSR(offs, toks, v)
    st = rr.engine.BM25Stats.local(offs, toks, v).finalize()
    np.testing.assert_array_equal(st.df, csr.df)
    assert st.avgdl == csr.avgdl and st.average_idf == csr.average_idf
    np.testing.assert_array_equal(st.idf, csr.idf)

    hp = rr.engine.build_postings(offs, toks, st, tile_docs=tile, n_threads=3)
    assert hp.n_tiles == (n + tile - 1) // tile
    docs = (hp.data & np.uint64(0xFFFFFFFF)).astype(np.int64)
    imp = (hp.data >> np.uint64(32)).astype(np.uint32).view(np.float32)
    # term classes: frequent <=> local df >= RR_DIR_MIN_PER_TILE * n_tiles; slots are term-ascending
    theta = lib.rr_bm25_dir_threshold(hp.n_tiles)
    assert theta == 8 * hp.n_tiles
    freq = np.nonzero(csr.df >= theta)[0]
    assert hp.n_freq == len(freq)
    np.testing.assert_array_equal(hp.term_slot[freq], np.arange(len(freq)))
    assert np.all(hp.term_slot[csr.df < theta] == -1)
    stride = hp.n_freq + 1
    seen = 0
    for t in range(hp.n_tiles):
        base = int(hp.tile_base[t])
        assert base % 2 == 0
        off = hp.dir[t * stride:(t + 1) * stride].astype(np.int64)
        assert off[0] == 0 and np.all(np.diff(off) >= 0)
        for slot in np.nonzero(np.diff(off))[0]:
            term = int(freq[slot])
            lo, hi = base + off[slot], base + off[slot + 1]
            d = docs[lo:hi]
            assert np.all(np.diff(d) > 0) and d[0] >= t * tile and d[-1] < min(n, (t + 1) * tile)
            odocs, ocontrib = csr.impacts(term)
            sel = (odocs >= t * tile) & (odocs < (t + 1) * tile)
            np.testing.assert_array_equal(d, odocs[sel])
            np.testing.assert_array_equal(imp[lo:hi], ocontrib[sel].astype(np.float32))
            seen += hi - lo
    # rare terms: one term-major, doc-ascending list each, behind the frequent region
    rare_base = int(hp.tile_base[hp.n_tiles])
    assert rare_base % 2 == 0 and hp.rare_off[0] == 0
    for term in np.nonzero(csr.df < theta)[0]:
        lo, hi = rare_base + int(hp.rare_off[term]), rare_base + int(hp.rare_off[term + 1])
        assert hi - lo == csr.df[term]
        if hi > lo:
            odocs, ocontrib = csr.impacts(int(term))
            np.testing.assert_array_equal(docs[lo:hi], odocs)
            np.testing.assert_array_equal(imp[lo:hi], ocontrib.astype(np.float32))
            seen += hi - lo
    for term in freq:
        assert hp.rare_off[term] == hp.rare_off[term + 1]
    assert seen == csr.post_doc.shape[0]
    assert hp.data.size == rare_base + (int(hp.rare_off[-1]) + 1) // 2 * 2


def test_sharded_stats_reduce_to_global(lib):
    n, v = 4000, 300
    offs, toks = rr.synth.corpus_tokens(n, v)
    whole = rr.engine.BM25Stats.local(offs, toks, v).finalize()
    cut = 1700
    a = rr.engine.BM25Stats.local(offs[:cut + 1], toks[:offs[cut]], v, token_pos0=0)
    b = rr.engine.BM25Stats.local(offs[cut:], toks[offs[cut]:], v, token_pos0=int(offs[cut]))
    merged = rr.engine.BM25Stats(v, a.df + b.df, np.minimum(a.first_pos, b.first_pos),
                                 a.total_tokens + b.total_tokens, a.n_docs + b.n_docs).finalize()
    np.testing.assert_array_equal(merged.idf, whole.idf)
    assert merged.avgdl == whole.avgdl and merged.average_idf == whole.average_idf


def test_flatten_corpus_matches_the_dict_walk():
    """drop_in.flatten_corpus (itertools.chain + pandas.factorize) assigns the ids a Python dict walk would
    (first-appearance order, which rank_bm25's idf mean depends on), including empty documents."""
    import review_recommender_b200 as rr
    from oracle.bm25_okapi import flatten_corpus as dict_walk
    offs, toks = rr.synth.corpus_tokens(2000, 300)
    corpus = rr.synth.corpus_as_lists(offs, toks)
    corpus[0] = []
    corpus[7] = []
    corpus.append(["new-token", "t1", "new-token"])
    o1, i1, v1 = rr.drop_in.flatten_corpus(corpus)
    o2, i2, v2 = dict_walk(corpus)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(i1, i2)
    assert list(v1.keys()) == v2 and list(v1.values()) == list(range(len(v2)))
    o3, i3, v3 = rr.drop_in.flatten_corpus([[], []])
    assert o3.tolist() == [0, 0, 0] and i3.size == 0 and v3 == {}
