"""oracle.gates (and the host-side build_gate_groups mirror) against the reference's own gate functions
captured in tests/golden/gate_cases.json.  CPU only."""
import json

import numpy as np
import pytest

from oracle.gates import build_gate_groups, calculate_gate_factor, pool_gate_factors
from tests.golden_worlds import GATE_QUERIES, make_gate_texts
import review_recommender_b200 as rr


@pytest.fixture(scope="module")
def g(golden_dir):
    gc = json.loads((golden_dir / "gate_cases.json").read_text())
    z = np.load(golden_dir / "search_cases.npz")
    texts = make_gate_texts(z["emb"].shape[0], z["doc_offsets"], z["token_ids"])
    return gc, texts


def test_groups_match_reference(g):
    gc, _ = g
    assert gc["queries"] == GATE_QUERIES
    for q, want in zip(gc["queries"], gc["groups"]):
        assert [sorted(x) for x in build_gate_groups(q)] == want
        assert [sorted(x) for x in rr.drop_in.build_gate_groups(q)] == want       # host mirror used by the product
    assert any(len(w) == 6 for w in gc["groups"]) and any(len(w) == 0 for w in gc["groups"])


def test_factors_match_reference(g):
    gc, texts = g
    for c in gc["factor_cases"]:
        q = gc["queries"][c["query_index"]]
        groups = build_gate_groups(q)
        got = [calculate_gate_factor(str(texts[r])[:6000], groups, c["penalty"]) for r in c["rows"]]
        np.testing.assert_array_equal(np.array([f for f, _, _ in got], dtype=np.float32), np.float32(c["gate_f32"]))
        assert [h for _, h, _ in got] == c["hits"]
        np.testing.assert_array_equal(pool_gate_factors([texts[r] for r in c["rows"]], q, c["penalty"]),
                                      np.float32(c["gate_f32"]))
