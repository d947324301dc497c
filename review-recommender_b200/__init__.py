"""B200-native first-stage hybrid retrieval (BM25 + dense cosine + fusion + top-k) behind the
search-function signatures of Ntropy86/review-recommender.

Submodules are imported lazily so that `synth` (pure NumPy) can be used without the CUDA
library; everything that computes goes through `librr_b200.so` (see `_lib`) and raises if it
is missing -- there is no CPU fallback.
"""
import importlib as _importlib

__all__ = ["synth", "_lib", "engine", "drop_in", "dist"]


def __getattr__(name):
    if name in __all__:
        return _importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
