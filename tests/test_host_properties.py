"""Property tests (hypothesis) of the host-side mirrors that run on the product path: the query tokenizer and the
gate-group builder of drop_in.py against the pinned oracle restatements, for arbitrary unicode query strings; the
round-1 shortlist rule of the sharded search; the corpus flattening."""
import numpy as np
from hypothesis import given, settings, strategies as st

import review_recommender_b200 as rr
from oracle import gates as G
from oracle import primitives as P
from oracle.bm25_okapi import flatten_corpus as dict_walk

WORDS = ["yellow", "Gold", "golden", "navy", "cat", "cats", "Cat's", "wireless", "bluetooth", "headphones", "noise",
         "anc", "design", "sock", "the", "of", "and", "x", "usb-c", "42", "straße", "İstanbul", "café", "keyboards",
         "tan", "rose", "supercalifragilistic"]
query = st.one_of(st.text(max_size=60),
                  st.lists(st.sampled_from(WORDS), max_size=10).map(" ".join),
                  st.lists(st.one_of(st.sampled_from(WORDS), st.text(max_size=8)), max_size=8).map(" ".join))


@settings(max_examples=400, deadline=None)
@given(query)
def test_tokenizer_and_gate_groups_mirror_the_oracle(q):
    assert rr.drop_in.tokenize_query(q) == P.tokenize_query(q)
    mine = rr.drop_in.build_gate_groups(q)
    want = G.build_gate_groups(q)
    assert [set(g) for g in mine] == want
    assert len(mine) <= rr.drop_in.GATE_MAX_GROUPS


def test_gate_tables_equal_the_oracle_tables():
    assert {k: set(v) for k, v in rr.drop_in.GATE_COLORS.items()} == G.COLORS
    assert {k: set(v) for k, v in rr.drop_in.GATE_SYNONYMS.items()} == G.SYNONYMS
    assert list(rr.drop_in.GATE_COLORS) == list(G.COLORS) and list(rr.drop_in.GATE_SYNONYMS) == list(G.SYNONYMS)
    assert len(rr.drop_in.GATE_FIXED_GROUPS) == 19


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 4096), st.integers(1, 64))
def test_round1_shortlist_rule(pool, world):
    m = rr.dist.local_pool(pool, world)
    assert 1 <= m <= pool
    if world == 1:
        assert m == pool
    else:
        assert m == pool or (m % 16 == 0 and m >= pool / world)      # never below the expected share of the pool
    assert rr.dist.local_pool(pool, world) >= rr.dist.local_pool(pool, world * 2) or m == pool


@settings(max_examples=100, deadline=None)
@given(st.lists(st.lists(st.sampled_from(WORDS + ["a", "b", "c"]), max_size=7), max_size=30))
def test_flatten_corpus_equals_the_dict_walk(corpus):
    o1, i1, v1 = rr.drop_in.flatten_corpus(corpus)
    o2, i2, v2 = dict_walk(corpus)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(i1, i2.astype(np.int32) if len(i2) else i1)
    assert list(v1.keys()) == v2


def test_synthetic_recipe_is_pinned():
    """The SURVEY 8d recipe (rr.synth) decides the benchmark corpus, the queries and hence bench.py's `result_digest`, which
    must be identical for every GPU count and every round: a digest of recipe samples guards against drift."""
    import hashlib
    import numpy as np
    import review_recommender_b200 as rr
    s = rr.synth
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s.embeddings(2000, 384)[::97]).tobytes())
    offs, toks = s.corpus_tokens(3000, 50000)
    h.update(offs.tobytes())
    h.update(toks.tobytes())
    nr, av = s.metadata(3000)
    h.update(nr.tobytes())
    h.update(av.tobytes())
    h.update(s.queries(64, 384).tobytes())
    h.update(s.query_terms(64, 4, offs, toks, 50000).tobytes())
    assert h.hexdigest() == "51d2ae0ec74a291c033ed77129101e59c208a6776ff59bafac551eb831039612"
    # rows of a shard are the same rows whichever rank generates them
    a = s.embeddings(700, 16, row0=1_999_800)
    b = s.embeddings(400, 16, row0=2_000_000)
    np.testing.assert_array_equal(a[200:600], b)


def test_ptr_takes_the_fast_route_and_falls_back():
    """engine._ptr (on the single-query latency path) returns the address ndarray.ctypes.data would, also for arrays
    the buffer-protocol route refuses (read-only, empty, non-contiguous), and None for None."""
    import numpy as np
    import torch
    from review_recommender_b200.engine import _ptr
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    ro = a.copy()
    ro.flags.writeable = False
    for arr in (a, a[1:], a[:, ::2], ro, np.empty((0, 6), np.float32), np.empty(0, np.int64)):
        assert _ptr(arr) == arr.ctypes.data
    t = torch.arange(5)
    assert _ptr(t) == t.data_ptr()
    assert _ptr(None) is None
    import pytest
    with pytest.raises(TypeError):
        _ptr([1, 2, 3])
