"""semantic-search-kd_b200 -- B200-native exact inner-product top-k for the semantic-search-kd
retrieval path (import name: ``semantic_search_kd_b200``).

Only the hot path lives here: the C-ABI CUDA library (``csrc/``, ``include/b200search.h``) and the
host-side mirror of the reference's index surface (``FAISSIndexBuilder`` -> ``FlatIPIndex``).
"""
from . import _lib  # noqa: F401
from .errors import (DeviceError, IndexBuildError, IndexNotBuiltError, IndexNotFoundError,  # noqa: F401
                     SearchIndexError)
from .index import FAISSIndexBuilder, FlatIPIndex  # noqa: F401
from .mining import ANCEMiner, maxsim_topk, retrieve_topk, similarity  # noqa: F401

__all__ = ["FlatIPIndex", "FAISSIndexBuilder", "ANCEMiner", "similarity", "retrieve_topk", "maxsim_topk", "IndexNotFoundError", "IndexNotBuiltError",
           "IndexBuildError", "DeviceError", "SearchIndexError"]
