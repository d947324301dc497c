"""Zero-edit drop-in for the third-party `rank_bm25` module.

Put this directory on sys.path ahead of site-packages and the reference's own imports
(`from rank_bm25 import BM25Okapi`, app/app_product_search.py:122, app/test.py:101) pick up the
GPU-backed class.  See INTEGRATION.md."""
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from review_recommender_b200.drop_in import BM25Okapi  # noqa: E402,F401

__all__ = ["BM25Okapi"]
