"""Error types of the index surface.

Mirrors the reference's index error convention (``/root/reference/src/exceptions.py:104-139``:
``IndexNotFoundError`` / ``IndexNotBuiltError`` / ``IndexBuildError`` with ``error_code`` and
``details``) so that callers written against the reference keep working.  When the reference's
own ``src.exceptions`` module is importable (i.e. this package runs inside the reference tree)
the classes below derive from the reference's, so ``except SemanticKDError`` still catches them.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

try:  # inside the reference tree: share its hierarchy
    from src.exceptions import SemanticKDError as _Base  # type: ignore
except Exception:  # standalone
    class _Base(Exception):  # type: ignore[no-redef]
        def __init__(self, message: str, error_code: Optional[str] = None,
                     details: Optional[Dict[str, Any]] = None) -> None:
            super().__init__(message)
            self.message = message
            self.error_code = error_code or "SEMANTIC_KD_ERROR"
            self.details = details or {}

        def to_dict(self) -> Dict[str, Any]:
            return {"error": self.error_code, "message": self.message, "details": self.details}


class SearchIndexError(_Base):
    """Base of the index errors (the reference names it ``IndexError``, shadowing the builtin)."""


class IndexNotFoundError(SearchIndexError):
    def __init__(self, index_path: str) -> None:
        super().__init__(f"Index not found at: {index_path}", error_code="INDEX_NOT_FOUND",
                         details={"index_path": str(index_path)})


class IndexNotBuiltError(SearchIndexError):
    def __init__(self) -> None:
        super().__init__("Index has not been built. Call build() first.", error_code="INDEX_NOT_BUILT")


class IndexBuildError(SearchIndexError):
    def __init__(self, message: str, documents_processed: int = 0) -> None:
        super().__init__(message, error_code="INDEX_BUILD_ERROR",
                         details={"documents_processed": documents_processed})


class DeviceError(SearchIndexError):
    """The CUDA extension is missing or no B200 is visible.  There is no CPU fallback."""

    def __init__(self, message: str, code: int = 0) -> None:
        super().__init__(message, error_code="B200_DEVICE_ERROR", details={"code": code})
