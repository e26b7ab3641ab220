"""CPU: host-side logic of the index wrapper that needs no GPU (file formats, errors, shapes)."""
import json
import struct

import numpy as np
import pytest

import semantic_search_kd_b200 as pkg
from semantic_search_kd_b200 import faiss_io
from conftest import unit_rows


def test_faiss_flat_file_layout(tmp_path):
    """The conftest fixture of the reference (10 x 384 IndexFlatIP) is 15 405 bytes."""
    X = unit_rows(10, 384, 42)
    n = faiss_io.write_flat_ip(tmp_path / "index.faiss", [X[:4], X[4:]], 10, 384)
    assert n == 15405 == (tmp_path / "index.faiss").stat().st_size
    raw = (tmp_path / "index.faiss").read_bytes()
    assert raw[:4] == b"IxFI"
    d, ntotal = struct.unpack("<iq", raw[4:16])
    assert (d, ntotal) == (384, 10)
    assert struct.unpack("<Q", raw[37:45])[0] == 3840
    back, metric = faiss_io.read_flat(tmp_path / "index.faiss")
    assert metric == 0 and np.array_equal(np.asarray(back), X)


def test_faiss_reader_rejects_other_index_types(tmp_path):
    p = tmp_path / "index.faiss"
    p.write_bytes(b"IHNf" + b"\0" * 100)
    with pytest.raises(ValueError, match="unsupported FAISS index type"):
        faiss_io.read_flat(p)
    p.write_bytes(b"IxFI" + b"\0" * 10)
    with pytest.raises(ValueError, match="truncated"):
        faiss_io.read_flat(p)


def test_faiss_writer_checks_row_count(tmp_path):
    with pytest.raises(ValueError):
        faiss_io.write_flat_ip(tmp_path / "x.faiss", [unit_rows(3, 384, 0)], 5, 384)


def test_constructor_mirrors_reference_signature():
    # scripts/build_faiss_index.py:49-53 and src/serve/app.py:427-429
    a = pkg.FAISSIndexBuilder(embedding_dim=384, index_type="HNSW", metric="cosine")
    b = pkg.FAISSIndexBuilder(embedding_dim=384)
    assert a.ntotal == 0 and b.ntotal == 0 and a.doc_ids == []
    assert a.index is a
    with pytest.raises(pkg.IndexBuildError):
        pkg.FAISSIndexBuilder(embedding_dim=384, metric="l2")


def test_errors_follow_reference_convention(tmp_path):
    idx = pkg.FAISSIndexBuilder(embedding_dim=384)
    with pytest.raises(pkg.IndexNotBuiltError) as e:
        idx.search(np.zeros((1, 384), np.float32), 3)
    assert e.value.error_code == "INDEX_NOT_BUILT"
    with pytest.raises(pkg.IndexNotFoundError) as e:
        idx.load(tmp_path / "missing")
    assert e.value.error_code == "INDEX_NOT_FOUND" and "missing" in e.value.details["index_path"]
    (tmp_path / "empty").mkdir()
    with pytest.raises(pkg.IndexNotFoundError):
        idx.load(tmp_path / "empty")
    with pytest.raises(pkg.IndexNotBuiltError):
        idx.save(tmp_path / "out")
    err = pkg.IndexBuildError("boom", documents_processed=7)
    assert err.to_dict() == {"error": "INDEX_BUILD_ERROR", "message": "boom",
                             "details": {"documents_processed": 7}}


def test_bad_shapes_rejected_before_touching_the_device():
    idx = pkg.FAISSIndexBuilder(embedding_dim=384)
    with pytest.raises(pkg.IndexBuildError):
        idx.add(np.zeros((3, 100), np.float32))
    with pytest.raises(pkg.IndexBuildError):
        idx.add(np.zeros((384,), np.float32))
