"""Reader / writer of the FAISS flat index file the reference's index directory holds.

Directory layout of the reference (``/root/reference/tests/conftest.py:187-198``,
``src/serve/app.py:430-441``): ``index.faiss`` + ``doc_ids.json`` (+ ``texts.json``).  The fixture
writes ``faiss.write_index(faiss.IndexFlatIP(384), ...)``; that binary layout, restated from
faiss' published ``index_write.cpp`` (faiss is not installable here, so this is unverified
against a real file -- SURVEY.md App. C):

    "IxFI" | int32 d | int64 ntotal | int64 1<<20 | int64 1<<20 | uint8 is_trained
           | int32 metric_type (0 = inner product, 1 = L2) | uint64 count (= ntotal*d) | float32[count]

("IxF2" is the L2 flavour, same body).  Little-endian throughout.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Iterable, Tuple

import numpy as np

FOURCC_IP = b"IxFI"
FOURCC_L2 = b"IxF2"
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


def read_header(path: Path) -> Tuple[int, int, int]:
    """(d, ntotal, metric_type) of an IndexFlat file; ValueError for anything else."""
    with open(Path(path), "rb") as f:
        head = f.read(_HEADER.size)
    if len(head) < _HEADER.size:
        raise ValueError(f"{path}: truncated FAISS header")
    fourcc, d, ntotal, _d1, _d2, _trained, metric = _HEADER.unpack(head)
    if fourcc not in (FOURCC_IP, FOURCC_L2):
        raise ValueError(f"{path}: unsupported FAISS index type {fourcc!r}")
    return int(d), int(ntotal), int(metric)


def read_flat(path: Path) -> Tuple[np.ndarray, int]:
    """Memory-map the vectors of an IndexFlat file.  Returns (fp32 [ntotal, d] memmap, metric)."""
    path = Path(path)
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
    if len(head) < HEADER_BYTES:
        raise ValueError(f"{path}: truncated FAISS header")
    fourcc, d, ntotal, _d1, _d2, _trained, metric = _HEADER.unpack(head[:_HEADER.size])
    if fourcc not in (FOURCC_IP, FOURCC_L2):
        raise ValueError(f"{path}: unsupported FAISS index type {fourcc!r}; only flat indexes "
                         "(IxFI / IxF2) can be loaded into the exact-search index")
    if metric > 1:
        raise ValueError(f"{path}: unsupported metric_type {metric}")
    (count,) = struct.unpack("<Q", head[_HEADER.size:])
    if count != ntotal * d:
        raise ValueError(f"{path}: vector count {count} != ntotal*d {ntotal * d}")
    expect = HEADER_BYTES + count * 4
    if path.stat().st_size < expect:
        raise ValueError(f"{path}: file shorter than header claims ({path.stat().st_size} < {expect})")
    if ntotal == 0:
        return np.zeros((0, d), dtype=np.float32), metric
    data = np.memmap(path, dtype="<f4", mode="r", offset=HEADER_BYTES, shape=(ntotal, d))
    return data, metric
