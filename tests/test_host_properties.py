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
