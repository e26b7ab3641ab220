"""Host-side mirrors of the reference's callers of the dense path: ``compute_similarity``,
``ANCEMiner`` and the exact-retrieval evaluation loop -- same names, argument meaning and results,
with the arithmetic on the B200 (C ABI ``b2s_similarity`` / ``b2s_search*`` /
``b2s_score_rows_device`` / ``b2s_ance_filter_device``).  No CPU fallback.

* ``similarity(q, d)``            <- ``StudentModel.compute_similarity``
                                     (``/root/reference/tests/test_student_model.py:104-124``)
* ``ANCEMiner.mine(...)``         <- ``/root/reference/src/mining/miners.py:184-253`` (same signature:
                                     per-query candidate lists, margin filter, top-k)
* ``ANCEMiner.mine_corpus(...)``  <- the corpus-wide ANCE the reference's design calls for
                                     (``docs/decisions/adr-003``; ``configs/kd.yaml:93-100``): candidates
                                     are the query's exact top-(top_k + positives) of the WHOLE index
* ``retrieve_topk(...)``          <- the scan + ``argsort[::-1][:k]`` of
                                     ``/root/reference/src/kd/eval.py:65-86`` and
                                     ``scripts/simple_eval.py:25,35`` without the ``[Q, N]`` matrix
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .errors import IndexBuildError
from .index import FlatIPIndex, _check

try:
    import torch
except Exception:  # pragma: no cover
    torch = None  # type: ignore


def similarity(q: np.ndarray, d: np.ndarray, device: int = 0) -> np.ndarray:
    """``compute_similarity(q, d) -> float32 [nq, nd]`` (dense ``q @ d.T`` in fp32 on the GPU)."""
    q = np.ascontiguousarray(np.atleast_2d(np.asarray(q, dtype=np.float32)))
    d = np.ascontiguousarray(np.atleast_2d(np.asarray(d, dtype=np.float32)))
    if d.shape[0] and q.shape[1] != d.shape[1]:
        raise IndexBuildError(f"dimension mismatch: {q.shape} vs {d.shape}")
    out = np.empty((q.shape[0], d.shape[0]), dtype=np.float32)
    if out.size:
        _check(_lib.lib().b2s_similarity(int(device), q.ctypes.data_as(ctypes.c_void_p), q.shape[0],
                                         d.ctypes.data_as(ctypes.c_void_p), d.shape[0], q.shape[1],
                                         out.ctypes.data_as(ctypes.c_void_p)), "b2s_similarity")
    return out


def retrieve_topk(index: FlatIPIndex, query_embs, k: int):
    """Exact top-k ids per query (what ``np.argsort(scores)[::-1][:k]`` selects, ties by ascending id)."""
    return index.search(query_embs, k)[1]


def maxsim_topk(index: FlatIPIndex, query_embs, k: int, chunk_to_doc, k_chunks: Optional[int] = None):
    """Document-level top-k by MaxSim: search ``k_chunks`` (default 4k) chunk rows per query, keep each
    document's best chunk (``b2s_maxsim_device``).  ``chunk_to_doc``: int64 ``[ntotal]``, chunk row ->
    document id.  Returns numpy ``(doc_scores float32 [nq,k], doc_ids int64 [nq,k] (-1 padded))``."""
    if torch is None or not torch.cuda.is_available():
        raise IndexBuildError("maxsim_topk needs torch with CUDA (device tensors are the plumbing)")
    dev = torch.device("cuda", index.device if index.device is not None else 0)
    q = query_embs if isinstance(query_embs, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(np.atleast_2d(np.asarray(query_embs, dtype=np.float32))))
    q = q.to(dev).contiguous()
    c2d = chunk_to_doc if isinstance(chunk_to_doc, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(np.asarray(chunk_to_doc, dtype=np.int64)))
    c2d = c2d.to(dev).contiguous()
    k_in = min(2048, int(k_chunks) if k_chunks else 4 * int(k))
    scores, ids = index.search_device(q, k_in)
    nq = q.shape[0]
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.int64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _check(_lib.lib().b2s_maxsim_device(dev.index, ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ids.data_ptr()), nq,
                                        k_in, ctypes.c_void_p(c2d.data_ptr()), c2d.shape[0], int(k),
                                        ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_d.data_ptr()), None, stream),
           "b2s_maxsim_device")
    return out_s.cpu().numpy(), out_d.cpu().numpy()


class ANCEMiner:
    """Stage-3 adversarial negative mining with the student's own scores."""

    def __init__(self, student_model: Any, margin: float = 0.1, device: int = 0) -> None:
        self.student = student_model
        self.margin = float(margin)
        self.device = int(device)

    # ------------------------------------------------------------- reference signature
    def mine(self, queries: List[str], positives: List[List[str]], candidates: List[List[str]],
             candidate_texts: Dict[str, str], positive_texts: Dict[str, str], top_k: int = 5) -> List[List[str]]:
        """Same contract as the reference: per query, candidates whose student score is within
        ``margin`` of the best positive, by descending score, first ``top_k`` doc ids."""
        out: List[List[str]] = []
        for query, pos_ids, cand_ids in zip(queries, positives, candidates):
            q = np.asarray(self.student.encode_queries([query]), dtype=np.float32)[0]
            pos_embs = np.asarray(self.student.encode_documents([positive_texts.get(i, "") for i in pos_ids]),
                                  dtype=np.float32).reshape(len(pos_ids), -1)
            cand_embs = np.asarray(self.student.encode_documents([candidate_texts.get(i, "") for i in cand_ids]),
                                   dtype=np.float32).reshape(len(cand_ids), -1)
            pos_scores = similarity(q.reshape(1, -1), pos_embs, self.device)[0] if len(pos_ids) else np.zeros(0)
            cand_scores = similarity(q.reshape(1, -1), cand_embs, self.device)[0] if len(cand_ids) else np.zeros(0)
            max_pos = float(pos_scores.max()) if len(pos_scores) > 0 else 0.0
            adv = [(doc, s) for doc, s in zip(cand_ids, cand_scores) if s >= max_pos - self.margin]
            adv.sort(key=lambda x: x[1], reverse=True)      # stable, like the reference
            out.append([doc for doc, _ in adv[:top_k]])
        return out

    # ------------------------------------------------------------- corpus-wide
    def mine_corpus(self, index: FlatIPIndex, query_embs, positive_ids: Sequence[Sequence[int]],
                    top_k: int = 200) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Candidates = the whole index.  ``positive_ids[i]`` are row ids of query i's positives.

        Returns ``(neg_ids int64 [nq, top_k] (-1 padded), neg_scores float32, counts int32 [nq])``.
        One batched exact search for ``top_k + max positives`` neighbours (tensor path), a gather-dot
        for the positives' scores and one filter kernel; nothing is materialised per candidate on
        the host."""
        if torch is None or not torch.cuda.is_available():
            raise IndexBuildError("mine_corpus needs torch with CUDA (device tensors are the plumbing)")
        if index.metric == "cosine":
            raise IndexBuildError("mine_corpus expects unit-norm embeddings in an inner_product index")
        dev = torch.device("cuda", index.device if index.device is not None else self.device)
        q = query_embs if isinstance(query_embs, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(query_embs, dtype=np.float32)))
        q = q.to(dev).contiguous()
        nq = q.shape[0]
        n_pos = max(1, max((len(p) for p in positive_ids), default=0))
        pos = np.full((nq, n_pos), -1, dtype=np.int64)
        for i, p in enumerate(positive_ids):
            pos[i, :len(p)] = np.asarray(p, dtype=np.int64)
        pos_d = torch.from_numpy(pos).to(dev)
        k_in = min(2048, top_k + n_pos)
        scores, ids = index.search_device(q, k_in)
        L = _lib.lib()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        pos_scores = torch.empty((nq, n_pos), dtype=torch.float32, device=dev)
        dt = _lib.DTYPE_BF16 if q.dtype == torch.bfloat16 else _lib.DTYPE_F32
        round_q = 1 if index.stats()["path"] == _lib.PATH_TENSOR else 0
        _check(L.b2s_score_rows_device(index._h, ctypes.c_void_p(q.data_ptr()), dt, round_q, nq,
                                       ctypes.c_void_p(pos_d.data_ptr()), n_pos,
                                       ctypes.c_void_p(pos_scores.data_ptr()), stream), "b2s_score_rows_device")
        out_ids = torch.empty((nq, top_k), dtype=torch.int64, device=dev)
        out_scores = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        _check(L.b2s_ance_filter_device(dev.index, ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                                        nq, k_in, ctypes.c_void_p(pos_d.data_ptr()),
                                        ctypes.c_void_p(pos_scores.data_ptr()), n_pos, ctypes.c_float(self.margin),
                                        int(top_k), ctypes.c_void_p(out_ids.data_ptr()),
                                        ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(counts.data_ptr()),
                                        stream), "b2s_ance_filter_device")
        return out_ids.cpu().numpy(), out_scores.cpu().numpy(), counts.cpu().numpy()
