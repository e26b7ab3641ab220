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
    p.write_bytes(b"IwPQ" + b"\0" * 100)
    with pytest.raises(ValueError, match="unsupported FAISS index type"):
        faiss_io.read_flat(p)
    p.write_bytes(b"IHNf" + b"\0" * 100)             # an HNSW file without a flat storage tail
    with pytest.raises(ValueError):
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


def test_small_call_staging_path():
    """index._small_call (the serving-sized host path: cached staging arrays instead of three ctypes pointers per
    call) with a fake C entry point: what it passes, what it returns, when it declines."""
    import ctypes
    import threading
    from semantic_search_kd_b200 import index as ix

    class Owner:
        _h = ctypes.c_void_p(1234)

    calls = []

    def fake(handle, aq, nq, k, as_, ai):
        q = np.ctypeslib.as_array((ctypes.c_float * (nq * 4)).from_address(aq)).copy()
        calls.append((handle, q, nq, k))
        s = np.ctypeslib.as_array((ctypes.c_float * (nq * k)).from_address(as_))
        i = np.ctypeslib.as_array((ctypes.c_int64 * (nq * k)).from_address(ai))
        s[:] = np.arange(nq * k, dtype=np.float32)[::-1]
        i[:] = np.arange(nq * k, dtype=np.int64)
        return 0

    o = Owner()
    q = np.arange(8, dtype=np.float32).reshape(2, 4)
    D, I = ix._small_call(o, fake, "fake", q, 2, 3)
    assert D.shape == (2, 3) and I.shape == (2, 3) and D.dtype == np.float32 and I.dtype == np.int64
    assert I.tolist() == [[0, 1, 2], [3, 4, 5]] and D[0, 0] == 5.0
    assert calls[0][0] is o._h and np.array_equal(calls[0][1], q.reshape(-1)) and calls[0][2:] == (2, 3)
    D2, I2 = ix._small_call(o, fake, "fake", q + 1, 2, 3)        # results are copies, not views of the staging arrays
    assert D is not D2 and I.tolist() == [[0, 1, 2], [3, 4, 5]]
    # too large for the staging arrays -> the caller takes the general path
    assert ix._small_call(o, fake, "fake", np.zeros((1, ix._SMALL_Q_FLOATS + 1), np.float32), 1, 3) is None
    assert ix._small_call(o, fake, "fake", q, 2, ix._SMALL_OUT) is None
    # another thread is inside -> decline instead of sharing the staging arrays
    lock = o.__dict__["_stage"][0]
    assert lock.acquire(False)
    try:
        out = []
        t = threading.Thread(target=lambda: out.append(ix._small_call(o, fake, "fake", q, 2, 3)))
        t.start()
        t.join()
        assert out == [None]
    finally:
        lock.release()
    # an error code from the library surfaces as the package's exception and releases the lock
    with pytest.raises(pkg.DeviceError):
        ix._small_call(o, lambda *a: pkg._lib.B2S_ERR_CUDA, "fake", q, 2, 3)
    assert ix._small_call(o, fake, "fake", q, 2, 3) is not None


def _write_hnsw_flat(path, X, M=32, extra_tail_fields=0):
    """An IndexHNSWFlat file as faiss 1.7.x lays it out (restated from its index_write.cpp; faiss is absent here):
    "IHNf" | index header | HNSW {assign_probas f64[], cum_nneighbor_per_level i32[], levels i32[], offsets u64[],
    neighbors i32[], entry_point, max_level, efConstruction, efSearch, upper_beam (i32 each)} | storage = flat index."""
    n, d = X.shape

    def vec(fmt, vals):
        a = np.asarray(vals, dtype=fmt)
        return struct.pack("<Q", a.size) + a.tobytes()

    levels = np.ones(n, dtype="<i4")
    offsets = np.arange(n + 1, dtype="<u8") * (2 * M)
    neighbors = np.full(n * 2 * M, -1, dtype="<i4")
    body = b"IHNf" + struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 0)
    body += vec("<f8", [1.0 - 1.0 / M, 0.0]) + vec("<i4", [0, 2 * M, 3 * M]) + vec("<i4", levels) + vec("<u8", offsets)
    body += vec("<i4", neighbors) + struct.pack("<5i", 0, 0, 200, 64, 1) + b"\0" * (4 * extra_tail_fields)
    path.write_bytes(body)
    with open(path, "ab") as f:
        f.write(b"IxFI" + struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 0) + struct.pack("<Q", n * d))
        f.write(np.ascontiguousarray(X, dtype="<f4").tobytes())


def test_hnsw_flat_file_gives_up_its_vectors(tmp_path):
    """The reference's own build script writes IndexHNSWFlat (index_type="HNSW"): its directory must load.  The reader
    takes the flat storage index from the END of the file, so it does not depend on the graph's exact layout
    (a version with extra HNSW fields reads the same)."""
    X = unit_rows(123, 384, 7)
    for extra in (0, 3):
        p = tmp_path / f"hnsw{extra}.faiss"
        _write_hnsw_flat(p, X, extra_tail_fields=extra)
        assert faiss_io.read_header(p) == (384, 123, 0)
        rows, metric = faiss_io.read_flat(p)
        assert metric == 0 and rows.shape == (123, 384) and np.array_equal(np.asarray(rows), X)
    # header says more vectors than the tail holds -> refused, not mis-read
    raw = bytearray((tmp_path / "hnsw0.faiss").read_bytes())
    raw[8:16] = struct.pack("<q", 124)
    (tmp_path / "bad.faiss").write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        faiss_io.read_flat(tmp_path / "bad.faiss")
