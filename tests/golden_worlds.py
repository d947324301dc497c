"""Deterministic input generators shared by tests/golden/make_golden.py (which feeds them to the reference)
and the tests (which feed the same inputs to the oracle and the CUDA path)."""
from __future__ import annotations

import numpy as np

GATE_WORDS = ["yellow", "Mustard", "lemon", "GOLDEN", "red", "navy", "olive", "black", "ivory", "rose", "violet", "amber",
              "beige", "charcoal", "sock", "socks", "headphones", "earbuds", "keyboard", "wireless", "Bluetooth",
              "noise cancelling", "noise-canceling", "ANC", "cat", "kitten", "dog", "puppies", "design", "pattern",
              "caf\u00e9", "\u00dcber", "\u0130stanbul", "stra\u00dfe", "\u4e2d\u6587", "cotton", "soft", "comfortable",
              "great", "quality", "battery", "mechanical", "gaming", "retriever", "scatter", "broadband"]
GATE_QUERIES = ["yellow cat socks", "wireless headphones with noise cancelling", "Golden retriever puppy design",
                "BLACK mechanical keyboard", "t12 t345 cotton socks", "caf\u00e9 \u00fcber soft", "the of and", "anc dog rose gray navy red green",
                "scat band tan", "supercalifragilisticexpialidocious-extraordinarily-long-token-number-one-two-three headphones"]


def make_gate_texts(N: int, doc_offsets, token_ids, seed: int = 4242):
    """agg_text per product: its BM25 tokens ("t17 t4 ...") mixed with gate vocabulary; a few very long texts
    (match only beyond 6000 chars), upper case, non-ASCII and empty."""
    rng = np.random.default_rng(seed)
    texts = []
    for i in range(N):
        toks = [f"t{int(t) + 1}" for t in token_ids[doc_offsets[i]:doc_offsets[i + 1]]]
        words = [GATE_WORDS[j] for j in rng.integers(0, len(GATE_WORDS), size=int(rng.integers(0, 9)))]
        mix = toks + words
        rng.shuffle(mix)
        t = " ".join(mix)
        if i % 53 == 0:
            t = ("x" * 41 + " ") * 150 + "yellow kitten wireless " + t      # gate words only after char 6000
        if i % 59 == 0:
            t = "\u0130" * 3000 + " socks " + "\u4e2d" * 2990 + " headphones cat" + t   # multi-byte text around the cut
        if i % 61 == 0:
            t = t.upper()
        texts.append(t)
    texts[3] = ""
    texts[8] = "NaN"          # (a real NaN makes the reference raise under pandas 3: .astype(str) keeps it missing)
    return texts
