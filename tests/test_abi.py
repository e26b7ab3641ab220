"""CPU: the C-ABI library builds, loads and exports every symbol include/b200search.h declares;
without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def L():
    import semantic_search_kd_b200 as pkg
    pkg._lib.build()
    return pkg._lib.lib()


def declared_symbols():
    text = (ROOT / "include" / "b200search.h").read_text()
    return sorted(set(re.findall(r"B2S_API\s+[\w\s\*]+?\b(b2s_\w+)\s*\(", text)))


def test_header_symbols_are_exported(L):
    import semantic_search_kd_b200 as pkg
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(pkg._lib.EXPORTS) == syms
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b200search.h but not exported"


def test_no_torch_types_in_signatures():
    text = (ROOT / "include" / "b200search.h").read_text()
    assert "torch" not in text.lower() and "at::" not in text and "std::" not in text


def test_version_and_error_string(L):
    assert L.b2s_version() >= 100
    assert isinstance(L.b2s_last_error(), bytes)


def test_argument_validation_without_device(L):
    import semantic_search_kd_b200 as pkg
    h = ctypes.c_void_p()
    assert L.b2s_create(383, 0, 0, ctypes.byref(h)) == pkg._lib.B2S_ERR_UNSUPPORTED
    assert b"dim" in L.b2s_last_error()
    assert L.b2s_create(384, 7, 0, ctypes.byref(h)) == pkg._lib.B2S_ERR_INVALID
    assert L.b2s_create(384, 0, 0, None) == pkg._lib.B2S_ERR_INVALID
    assert L.b2s_search(None, None, 1, 1, None, None) == pkg._lib.B2S_ERR_INVALID
    assert L.b2s_ntotal(None) == 0
    assert L.b2s_destroy(None) == 0


def test_fails_loudly_without_gpu(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import semantic_search_kd_b200 as pkg
    h = ctypes.c_void_p()
    assert L.b2s_create(384, 0, 0, ctypes.byref(h)) == pkg._lib.B2S_ERR_NO_DEVICE
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    with pytest.raises(pkg.DeviceError):
        idx.add(np.zeros((4, 384), np.float32))
    with pytest.raises(pkg.IndexNotBuiltError):
        idx.search(np.zeros((1, 384), np.float32), 5)
    q = np.zeros((1, 384), np.float32)
    out = np.zeros((1, 1), np.float32)
    rc = L.b2s_similarity(0, q.ctypes.data_as(ctypes.c_void_p), 1, q.ctypes.data_as(ctypes.c_void_p), 1, 384,
                          out.ctypes.data_as(ctypes.c_void_p))
    assert rc == pkg._lib.B2S_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (tier rule 3)."""
    pkg_dir = ROOT / "semantic-search-kd_b200"
    for p in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cuh")) + \
            list(pkg_dir.rglob("*.inl")):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p
        assert "_oracle.so" not in text, p
