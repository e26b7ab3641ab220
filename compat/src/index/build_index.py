"""Shim that makes the reference's import resolve to the B200 index:

    from src.index.build_index import FAISSIndexBuilder      # /root/reference/src/serve/app.py:21,
                                                              # scripts/build_faiss_index.py:9

The reference's own src/index/build_index.py is absent from its tree.  Put this
repo's ``compat/`` directory on PYTHONPATH ahead of (or copy this file into) the reference checkout.
"""
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[3]
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))

from semantic_search_kd_b200 import FAISSIndexBuilder, FlatIPIndex  # noqa: E402,F401
from semantic_search_kd_b200.errors import (IndexBuildError, IndexNotBuiltError,  # noqa: E402,F401
                                            IndexNotFoundError)
