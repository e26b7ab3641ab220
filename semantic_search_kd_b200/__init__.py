"""Importable alias of the ``semantic-search-kd_b200/`` package directory.

The directory name is fixed by the repo contract and is not a valid Python identifier, so this
shim points ``semantic_search_kd_b200``'s package path at it and runs its ``__init__``.
"""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "semantic-search-kd_b200"
__path__ = [str(_real)]
__file__ = str(_real / "__init__.py")
exec(compile((_real / "__init__.py").read_text(), __file__, "exec"))
