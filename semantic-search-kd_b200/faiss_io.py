"""Reader / writer of the FAISS flat index file the reference's index directory holds.

Directory layout of the reference (``/root/reference/tests/conftest.py:187-198``,
``src/serve/app.py:430-441``): ``index.faiss`` + ``doc_ids.json`` (+ ``texts.json``).  The fixture
writes ``faiss.write_index(faiss.IndexFlatIP(384), ...)``; that binary layout, restated from
faiss' published ``index_write.cpp`` (faiss is not installable here, so this is unverified
against a real file -- SURVEY.md App. C):

    "IxFI" | int32 d | int64 ntotal | int64 1<<20 | int64 1<<20 | uint8 is_trained
           | int32 metric_type (0 = inner product, 1 = L2) | uint64 count (= ntotal*d) | float32[count]

("IxF2" is the L2 flavour, same body).  Little-endian throughout.

An ``IndexHNSWFlat`` file -- what the reference's ``scripts/build_faiss_index.py`` leaves behind with its default
``index_type="HNSW"`` -- is ``"IHNf" | the same header | the HNSW graph (five length-prefixed vectors and five
ints) | the storage index``, and the storage index is a complete flat file ("IxFI" ...) holding every vector.  It is
the LAST thing in the file, so its position follows from the file size and the outer header alone: the graph
(whose exact layout differs between faiss versions) is skipped without being parsed, the inner header is
validated.  An exact index has no use for the graph; the vectors are all it needs.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Iterable, Tuple

import numpy as np

FOURCC_IP = b"IxFI"
FOURCC_L2 = b"IxF2"
FOURCC_HNSW_FLAT = b"IHNf"   # IndexHNSWFlat: what the reference's own build script writes (index_type="HNSW")
_HEADER = struct.Struct("<4siqqqBi")  # fourcc, d, ntotal, dummy, dummy, is_trained, metric
HEADER_BYTES = _HEADER.size + 8       # + uint64 count = 45


def write_flat_ip(path: Path, blocks: Iterable[np.ndarray], ntotal: int, dim: int) -> int:
    """Stream fp32 row blocks into an IndexFlatIP file.  Returns bytes written."""
    path = Path(path)
    written = 0
    with open(path, "wb") as f:
        f.write(_HEADER.pack(FOURCC_IP, dim, ntotal, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", ntotal * dim))
        written += HEADER_BYTES
        rows = 0
        for b in blocks:
            b = np.ascontiguousarray(b, dtype="<f4")
            if b.ndim != 2 or b.shape[1] != dim:
                raise ValueError(f"block of shape {b.shape} does not match dim {dim}")
            f.write(memoryview(b).cast("B"))
            rows += b.shape[0]
            written += b.nbytes
    if rows != ntotal:
        raise ValueError(f"wrote {rows} rows, header says {ntotal}")
    return written


def _locate_flat(path: Path) -> Tuple[int, int, int, int]:
    """(offset of the flat index inside the file, d, ntotal, metric_type) for a flat file or an IndexHNSWFlat file."""
    path = Path(path)
    size = path.stat().st_size
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
        if len(head) < _HEADER.size:
            raise ValueError(f"{path}: truncated FAISS header")
        fourcc, d, ntotal, _d1, _d2, _trained, metric = _HEADER.unpack(head[:_HEADER.size])
        if fourcc in (FOURCC_IP, FOURCC_L2):
            off = 0
        elif fourcc == FOURCC_HNSW_FLAT:
            if d <= 0 or ntotal < 0:
                raise ValueError(f"{path}: implausible header (d={d}, ntotal={ntotal})")
            off = size - (HEADER_BYTES + ntotal * d * 4)     # the storage index is the tail of the file
            if off < _HEADER.size:
                raise ValueError(f"{path}: IndexHNSWFlat file too short for {ntotal} x {d} stored vectors")
            f.seek(off)
            inner = f.read(HEADER_BYTES)
            ifourcc, idim, intotal, _a, _b, _t, imetric = _HEADER.unpack(inner[:_HEADER.size])
            if ifourcc not in (FOURCC_IP, FOURCC_L2) or idim != d or intotal != ntotal:
                raise ValueError(f"{path}: no flat storage index ({ntotal} x {d}) at the end of the IndexHNSWFlat file "
                                 f"(found {ifourcc!r}, {intotal} x {idim}): unsupported faiss version or storage type")
            metric = imetric
            head = inner
        else:
            raise ValueError(f"{path}: unsupported FAISS index type {fourcc!r}; only flat indexes (IxFI / IxF2) and "
                             "IndexHNSWFlat (IHNf, whose flat storage is read) can be loaded into the exact-search index")
    if len(head) < HEADER_BYTES:
        raise ValueError(f"{path}: truncated FAISS header")
    if metric > 1:
        raise ValueError(f"{path}: unsupported metric_type {metric}")
    (count,) = struct.unpack("<Q", head[_HEADER.size:HEADER_BYTES])
    if count != ntotal * d:
        raise ValueError(f"{path}: vector count {count} != ntotal*d {ntotal * d}")
    if size < off + HEADER_BYTES + count * 4:
        raise ValueError(f"{path}: file shorter than header claims ({size} < {off + HEADER_BYTES + count * 4})")
    return off, int(d), int(ntotal), int(metric)


def read_header(path: Path) -> Tuple[int, int, int]:
    """(d, ntotal, metric_type) of an IndexFlat / IndexHNSWFlat file; ValueError for anything else."""
    _off, d, ntotal, metric = _locate_flat(path)
    return d, ntotal, metric


def read_flat(path: Path) -> Tuple[np.ndarray, int]:
    """Memory-map the stored vectors of an IndexFlat or IndexHNSWFlat file.  Returns (fp32 [ntotal, d] memmap, metric)."""
    path = Path(path)
    off, d, ntotal, metric = _locate_flat(path)
    if ntotal == 0:
        return np.zeros((0, d), dtype=np.float32), metric
    data = np.memmap(path, dtype="<f4", mode="r", offset=off + HEADER_BYTES, shape=(ntotal, d))
    return data, metric
