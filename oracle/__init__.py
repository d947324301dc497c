"""CPU oracle for the hybrid-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy / pure Python, the arithmetic of the reference's
first-stage hybrid retrieval path (`utils.py`, `app/test.py`,
`app/app_product_search.py:179-370` and the third-party `rank_bm25.BM25Okapi`).
It exists to *check* the CUDA path; it is never the product:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
  `--impl reference` legs may import it;
* nothing under `review-recommender_b200/` imports it, and the product path raises if
  the CUDA extension is missing rather than falling back to this code.

Pinning status
--------------
* `oracle.primitives`, `oracle.hybrid`, `oracle.snippets`, `oracle.gates`: PINNED.  Checked against the reference's own
  functions run in the build container (`/root/reference/utils.py`,
  `/root/reference/app/test.py`, and `app/app_product_search.py` imported under a
  stub `streamlit`), via `tests/golden/make_golden.py` -> `tests/golden/*.json|npz`
  and live in `tests/test_oracle_vs_reference.py` when `/root/reference` exists.
* `oracle.bm25_okapi`: **PARITY UNPINNED**.  `rank_bm25` is a third-party dependency
  that is absent from `/root/reference`, absent from its `requirements.txt`, not
  installed and not fetchable (no network).  The class is restated from the
  published `rank_bm25` 0.2.x `BM25`/`BM25Okapi` algorithm and anchored on the
  reference's call sites (`app/test.py:156,170`, `app/app_product_search.py:142,206`),
  on the reference's fixture corpus (`tests/conftest.py:94-99`) and on the hand-checked
  golden vectors in SURVEY.md section 8c.
"""
