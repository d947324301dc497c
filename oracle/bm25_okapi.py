"""Restatement of `rank_bm25.BM25Okapi` (third-party, rank_bm25 0.2.x).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the package is not vendored in /root/reference, not pinned in its
requirements.txt, not installed and not fetchable here.  The algorithm below is the
published one, restated; it is anchored on

* the reference's construction call sites  `BM25Okapi(blob["corpus"])`
  (app/test.py:156, app/app_product_search.py:142): library defaults k1=1.5, b=0.75,
  epsilon=0.25, corpus = list of token lists;
* the reference's scoring call sites `bm25.get_scores(tokens)` -> float64[N]
  (app/test.py:170, app/app_product_search.py:206);
* the hand-checked golden vectors of SURVEY.md section 8c on the reference's own fixture
  corpus (tests/conftest.py:94-99).

Two implementations of the same arithmetic:

`BM25Okapi`     dict-per-document, pure Python, exactly the data structures and loop
                order of the library (this is what the reference executes; it is also
                the CPU baseline that bench.py times).
`BM25OkapiCSR`  float64 NumPy over integer token ids, for corpora too large for N dicts.
                Bit-identical to `BM25Okapi` (tests/test_oracle_bm25.py checks this).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import numpy as np

K1_DEFAULT = 1.5
B_DEFAULT = 0.75
EPSILON_DEFAULT = 0.25


class BM25Okapi:
    """rank_bm25.BM25Okapi, restated (see module docstring for anchors)."""

    def __init__(self, corpus: Sequence[Sequence[str]], tokenizer=None,
                 k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        self.tokenizer = tokenizer
        if tokenizer:
            corpus = [tokenizer(doc) for doc in corpus]
        nd = self._initialize(corpus)
        self._calc_idf(nd)

    def _initialize(self, corpus) -> Dict[str, int]:
        nd: Dict[str, int] = {}  # word -> number of documents containing it
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies: Dict[str, int] = {}
            for word in document:
                if word not in frequencies:
                    frequencies[word] = 0
                frequencies[word] += 1
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                if word in nd:
                    nd[word] += 1
                else:
                    nd[word] = 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        return nd

    def _calc_idf(self, nd: Dict[str, int]) -> None:
        # idf = ln(N - df + 0.5) - ln(df + 0.5); negatives are floored to
        # epsilon * mean(idf) where the mean is taken over ALL terms, negatives included,
        # summed sequentially in dict (first-appearance) order.
        idf_sum = 0
        negative_idfs = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative_idfs.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative_idfs:
            self.idf[word] = eps

    def get_scores(self, query: Iterable[str]) -> np.ndarray:
        """float64[N].  Duplicate query tokens are summed once per occurrence; unknown
        tokens (and tokens whose idf is exactly 0) contribute 0."""
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def idf_from_stats(df: np.ndarray, first_pos: np.ndarray, corpus_size: int, epsilon: float = EPSILON_DEFAULT):
    """`BM25Okapi._calc_idf` from corpus statistics: df int64[V] (documents containing the term) and first_pos int64[V]
    (flat position of the term's first occurrence -- the insertion order of the library's dict, in which the idf
    mean is accumulated).  -> (idf float64[V] with the epsilon floor applied, average_idf, number of present terms)."""
    present = np.nonzero(np.asarray(df) > 0)[0]
    appear = present[np.argsort(np.asarray(first_pos)[present], kind="stable")]
    idf = np.zeros(len(df), dtype=np.float64)
    idf_sum = 0
    negative = []
    for t in appear.tolist():
        freq = int(df[t])
        x = math.log(corpus_size - freq + 0.5) - math.log(freq + 0.5)
        idf[t] = x
        idf_sum += x
        if x < 0:
            negative.append(t)
    n_terms = int(appear.shape[0])
    average_idf = idf_sum / n_terms
    eps = epsilon * average_idf
    for t in negative:
        idf[t] = eps
    return idf, average_idf, n_terms


class BM25OkapiCSR:
    """Same arithmetic as `BM25Okapi`, over integer token ids (term id in [0, V)).

    corpus is given flat: `doc_offsets` int64[N+1], `token_ids` int[total].
    """

    def __init__(self, doc_offsets: np.ndarray, token_ids: np.ndarray, vocab_size: int,
                 k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPSILON_DEFAULT):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        doc_offsets = np.asarray(doc_offsets, dtype=np.int64)
        token_ids = np.asarray(token_ids, dtype=np.int64)
        n = int(doc_offsets.shape[0] - 1)
        v = int(vocab_size)
        self.corpus_size = n
        self.vocab_size = v
        self.doc_len = np.diff(doc_offsets)                       # int64[N]
        self.avgdl = int(self.doc_len.sum()) / n                  # python int / int, as the library
        doc_of_tok = np.repeat(np.arange(n, dtype=np.int64), self.doc_len)
        keys, tf = np.unique(doc_of_tok * v + token_ids, return_counts=True)
        p_doc = keys // v
        p_term = keys % v
        self.df = np.bincount(p_term, minlength=v).astype(np.int64)
        # term-major postings (doc ascending inside a term)
        order = np.lexsort((p_doc, p_term))
        self.post_doc = p_doc[order]
        self.post_tf = tf[order].astype(np.int64)
        self.term_off = np.zeros(v + 1, dtype=np.int64)
        np.cumsum(self.df, out=self.term_off[1:])
        # idf, with the mean accumulated in first-appearance order like the dict walk
        present, first_idx = np.unique(token_ids, return_index=True)
        first_pos = np.full(v, np.iinfo(np.int64).max, dtype=np.int64)
        first_pos[present] = first_idx
        self.idf, self.average_idf, self.n_terms = idf_from_stats(self.df, first_pos, n, self.epsilon)

    def impacts(self, term: int) -> tuple[np.ndarray, np.ndarray]:
        """(doc ids, float64 per-posting contribution) of one term."""
        lo, hi = int(self.term_off[term]), int(self.term_off[term + 1])
        docs = self.post_doc[lo:hi]
        q_freq = self.post_tf[lo:hi]
        doc_len = self.doc_len[docs]
        contrib = (self.idf[term] or 0) * (q_freq * (self.k1 + 1) /
                                           (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return docs, contrib

    def get_scores(self, query_ids: Iterable[int]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        for q in query_ids:
            q = int(q)
            if q < 0 or q >= self.vocab_size or self.df[q] == 0:
                continue  # unknown term: idf.get(q) is None -> contributes 0
            docs, contrib = self.impacts(q)
            score[docs] += contrib      # docs are unique inside one term
        return score


def flatten_corpus(corpus: Sequence[Sequence[str]]):
    """list[list[str]] -> (doc_offsets int64[N+1], token_ids int32[total], vocab list[str]).
    Term ids are assigned in first-appearance order (the dict order of the library)."""
    vocab: Dict[str, int] = {}
    ids: List[int] = []
    offs = [0]
    for doc in corpus:
        for w in doc:
            i = vocab.get(w)
            if i is None:
                i = len(vocab)
                vocab[w] = i
            ids.append(i)
        offs.append(len(ids))
    return (np.asarray(offs, dtype=np.int64), np.asarray(ids, dtype=np.int32), list(vocab.keys()))
