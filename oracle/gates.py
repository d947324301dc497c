"""Attribute gates, restated.  TEST INFRASTRUCTURE ONLY.

`build_gate_groups`      utils.py:62-86   (= _build_gate_groups app/app_product_search.py:211-226, app/test.py:62-78)
`calculate_gate_factor`  utils.py:88-101  (= _gate_factor app/app_product_search.py:228-236)
`pool_gate_factors`      the per-candidate loop of app/app_product_search.py:297-302 / app/test.py:291-297

PINNED by tests/golden/make_golden.py::golden_gates (tests/golden/gate_cases.json), which runs all three
copies of the reference's functions and both drivers with gate_penalty = 0.5 on product text.
"""
from __future__ import annotations

from typing import List, Sequence, Set, Tuple

import numpy as np

from .primitives import tokenize_query

SYNONYMS = {                                                   # utils.py:15-24
    "sock": {"sock", "socks"},
    "headphone": {"headphone", "headphones", "earphone", "earphones", "earbud", "earbuds", "headset"},
    "keyboard": {"keyboard", "keyboards"},
    "wireless": {"wireless", "bluetooth"},
    "noise": {"noise cancelling", "noise-canceling", "noise canceling", "anc"},
    "cat": {"cat", "cats", "kitten", "kittens", "kitty"},
    "dog": {"dog", "dogs", "puppy", "puppies"},
    "design": {"design", "pattern", "print", "graphic", "artwork", "motif", "theme"},
}
COLORS = {                                                     # utils.py:26-38
    "yellow": {"yellow", "mustard", "lemon", "gold", "golden"},
    "red": {"red", "scarlet", "crimson", "maroon"},
    "blue": {"blue", "navy", "cobalt", "azure"},
    "green": {"green", "emerald", "olive"},
    "black": {"black"},
    "white": {"white", "ivory"},
    "pink": {"pink", "rose"},
    "purple": {"purple", "violet", "lavender"},
    "orange": {"orange", "amber"},
    "brown": {"brown", "tan", "beige", "khaki"},
    "gray": {"gray", "grey", "charcoal", "slate"},
}


def build_gate_groups(query: str) -> List[Set[str]]:
    ql = query.lower()
    groups: List[Set[str]] = []
    for _name, syns in COLORS.items():                         # :68-70
        if any(w in ql for w in syns):
            groups.append(syns)
    for t in tokenize_query(query):                            # :73-78
        if t in SYNONYMS:
            groups.append(SYNONYMS[t])
        elif len(t) >= 4:
            groups.append({t})
    uniq: List[Set[str]] = []                                  # :81-84
    for g in groups:
        if g not in uniq:
            uniq.append(g)
    return uniq[:6]                                            # :86


def calculate_gate_factor(text: str, groups: Sequence[Set[str]], penalty: float = 0.5) -> Tuple[float, int, int]:
    tl = text.lower()
    hits, factor = 0, 1.0
    for g in groups:
        if any(s in tl for s in g):
            hits += 1
        else:
            factor *= penalty
    return factor, hits, len(groups)


def pool_gate_factors(texts: Sequence, query: str, penalty: float) -> np.ndarray:
    """float32[len(texts)]: the `_gate` column (`agg_text.astype(str).str.slice(0, 6000)`, :299)."""
    groups = build_gate_groups(query)
    return np.array([calculate_gate_factor(str(t)[:6000], groups, penalty)[0] for t in texts], dtype=np.float32)
