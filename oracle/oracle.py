"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the exact inner-product top-k path.  Two restatements:

* ``flat_ip_topk_np``  -- numpy, follows the reference's own in-repo exact
  search (``/root/reference/src/kd/eval.py:75,86``,
  ``scripts/simple_eval.py:25,35``: ``q @ c.T`` then ``argsort[::-1][:k]``) with
  faiss.IndexFlatIP's output convention (``tests/conftest.py:184-185``;
  ``src/serve/app.py:299-301``: shape ``[nq,k]``, id ``-1`` = no result).
* ``flat_ip_topk``     -- ctypes binding of ``oracle/flat_ip.c`` (same
  semantics, OpenMP, streams over corpus blocks) for sizes numpy cannot hold.

PINNING: pinned against outputs of the reference's own exact-search code run in
the build container (``tests/golden/make_ref_golden.py`` imports
``src/kd/eval.py``, ``scripts/simple_eval.py`` and ``src/mining/miners.py``
unmodified from ``/root/reference`` with the absent model stubbed; fixtures
``tests/golden/ref_eval.npz``, ``ref_ance.json``; test
``tests/test_reference_golden.py``).  faiss-cpu itself (``pyproject.toml:15``,
``^1.7.4``) is neither in ``/root/reference`` nor installable, and no reference
test pins a retrieved id or score, so the faiss half (tie order, -1 padding)
follows faiss' published semantics, see ``flat_ip.c``.  The older vectors under
``tests/golden/`` (``make_golden.py``) come from the numpy restatement on the
reference's fixture recipe (``tests/conftest.py:65-73``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_oracle.so"
FLT_LOWEST = np.float32(-3.4028234663852886e38)  # faiss' "no result" score

_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle/_oracle.so with gcc (see oracle/Makefile)."""
    srcs = [_HERE / "flat_ip.c", _HERE / "hnsw.cpp", _HERE / "Makefile"]
    stale = (not _SO.exists()) or any(s.stat().st_mtime > _SO.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-s", "-C", str(_HERE), "-B", "_oracle.so"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        L = ctypes.CDLL(str(_SO))
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.orc_flat_ip_topk.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                       f32p, i64p, ctypes.c_int]
        L.orc_flat_ip_topk.restype = ctypes.c_int
        L.orc_similarity.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int64,
                                     ctypes.c_int, f32p]
        L.orc_similarity.restype = ctypes.c_int
        L.orc_round_bf16.argtypes = [f32p, f32p, ctypes.c_int64]
        L.orc_round_bf16.restype = None
        L.orc_to_bf16_bits.argtypes = [f32p, ctypes.POINTER(ctypes.c_uint16), ctypes.c_int64]
        L.orc_to_bf16_bits.restype = None
        L.orc_gen_unit_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64,
                                        ctypes.c_int64]
        L.orc_gen_unit_rows.restype = None
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_hnsw_build.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        L.orc_hnsw_build.restype = ctypes.c_void_p
        L.orc_hnsw_search.argtypes = [ctypes.c_void_p, f32p, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_int, f32p, i64p, ctypes.c_int]
        L.orc_hnsw_search.restype = ctypes.c_int
        L.orc_hnsw_free.argtypes = [ctypes.c_void_p]
        L.orc_hnsw_free.restype = None
        L.orc_hnsw_max_level.argtypes = [ctypes.c_void_p]
        L.orc_hnsw_max_level.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


# ---------------------------------------------------------------------------
# numpy restatement (small cases)
# ---------------------------------------------------------------------------

def flat_ip_topk_np(X: np.ndarray, Q: np.ndarray, k: int,
                    acc: str = "f64") -> Tuple[np.ndarray, np.ndarray]:
    """IndexFlatIP.search restated with numpy.

    scores = Q @ X.T (``src/kd/eval.py:75``), k best per row in descending order
    (``eval.py:86``), ties by ascending id, padding ``(-FLT_MAX, -1)`` when
    ``k > n`` (guarded by ``src/serve/app.py:300``).
    """
    X = _f32(X).reshape(-1, X.shape[-1]) if X.size else _f32(X).reshape(0, Q.shape[-1])
    Q = _f32(Q)
    nq, n = Q.shape[0], X.shape[0]
    D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if nq == 0 or k == 0 or n == 0:
        return D, I
    if acc == "f64":
        S = (Q.astype(np.float64) @ X.astype(np.float64).T).astype(np.float32)
    else:
        S = Q @ X.T
    ids = np.arange(n, dtype=np.int64)
    kk = min(k, n)
    for i in range(nq):
        order = np.lexsort((ids, -S[i].astype(np.float64)))[:kk]  # score desc, id asc
        D[i, :kk] = S[i, order]
        I[i, :kk] = order
    return D, I


def similarity_np(Q: np.ndarray, X: np.ndarray) -> np.ndarray:
    """StudentModel.compute_similarity restated (``tests/test_student_model.py:104-124``)."""
    return (_f32(Q).astype(np.float64) @ _f32(X).astype(np.float64).T).astype(np.float32)


def round_bf16_np(a: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round-to-nearest-even) -> fp32, pure numpy."""
    u = _f32(a).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    u = (u + 0x7FFF + lsb) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def ance_filter_ref(cand_ids, cand_scores, pos_scores, margin: float, top_k: int):
    """ANCEMiner.mine's selection, restated (``src/mining/miners.py:237-247``).

    keep candidates with ``score >= max(pos_scores) - margin`` (0.0 when there
    are no positives), stable sort by descending score, first ``top_k`` ids.
    """
    max_pos = max(pos_scores) if len(pos_scores) > 0 else 0.0
    adv = [(d, s) for d, s in zip(cand_ids, cand_scores) if s >= max_pos - margin]
    adv.sort(key=lambda x: x[1], reverse=True)
    return [d for d, _ in adv[:top_k]]


def maxsim_ref(chunk_scores) -> dict:
    """MaxSim, restated (``src/utils/chunk.py:123-148``): ``[("{doc}_{chunk_idx}", score), ...]`` ->
    ``{doc: best chunk score}``; an id without ``_`` is its own document."""
    best: dict = {}
    for cid, sc in chunk_scores:
        doc = cid.rsplit("_", 1)[0] if "_" in cid else cid
        best[doc] = max(sc, best[doc]) if doc in best else sc
    return best


# ---------------------------------------------------------------------------
# C oracle (large cases, streaming)
# ---------------------------------------------------------------------------

def flat_ip_topk(X: np.ndarray, Q: np.ndarray, k: int, acc: str = "f64", id_offset: int = 0,
                 merge_into: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                 nthreads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    X = _f32(X)
    Q = _f32(Q)
    nq = Q.shape[0]
    d = Q.shape[1]
    n = X.shape[0] if X.ndim == 2 else 0
    if merge_into is None:
        D = np.full((nq, k), FLT_LOWEST, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        merge = 0
    else:
        D, I = merge_into
        assert D.dtype == np.float32 and I.dtype == np.int64 and D.flags.c_contiguous
        merge = 1
    if nq == 0 or k == 0 or n == 0:
        return D, I
    rc = lib().orc_flat_ip_topk(_p(X, ctypes.c_float), n, d, _p(Q, ctypes.c_float), nq, k,
                                1 if acc == "f64" else 0, id_offset, merge,
                                _p(D, ctypes.c_float), _p(I, ctypes.c_int64), nthreads)
    if rc != 0:
        raise RuntimeError("orc_flat_ip_topk failed")
    return D, I


def round_bf16(a: np.ndarray) -> np.ndarray:
    a = _f32(a)
    out = np.empty_like(a)
    lib().orc_round_bf16(_p(a, ctypes.c_float), _p(out, ctypes.c_float), a.size)
    return out


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    a = _f32(a)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().orc_to_bf16_bits(_p(a, ctypes.c_float), _p(out, ctypes.c_uint16), a.size)
    return out


def gen_unit_rows(n: int, d: int, seed: int, row_offset: int = 0) -> np.ndarray:
    out = np.empty((n, d), dtype=np.float32)
    lib().orc_gen_unit_rows(_p(out, ctypes.c_float), n, d, seed, row_offset)
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


class HnswRef:
    """HNSW restatement (M, efConstruction, efSearch as ``src/config.py:133-135``)."""

    def __init__(self, X: np.ndarray, M: int = 32, ef_construction: int = 200,
                 nthreads: int = 0, seed: int = 12345):
        self.X = _f32(X)  # kept alive: the C side borrows it
        n, d = self.X.shape
        self._h = lib().orc_hnsw_build(_p(self.X, ctypes.c_float), n, d, M, ef_construction,
                                       nthreads, seed)
        if not self._h:
            raise RuntimeError("orc_hnsw_build failed")

    def search(self, Q: np.ndarray, k: int, ef_search: int = 64, nthreads: int = 0):
        Q = _f32(Q)
        nq = Q.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        rc = lib().orc_hnsw_search(self._h, _p(Q, ctypes.c_float), nq, k, ef_search,
                                   _p(D, ctypes.c_float), _p(I, ctypes.c_int64), nthreads)
        if rc != 0:
            raise RuntimeError("orc_hnsw_search failed")
        return D, I

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.orc_hnsw_free(h)
            self._h = None


# ---------------------------------------------------------------------------
# parity rule (BASELINE.json north_star)
# ---------------------------------------------------------------------------

def recall_at_k(I_test: np.ndarray, I_ref: np.ndarray) -> float:
    """|test ∩ ref| / |ref| averaged over queries (ids < 0 ignored)."""
    tot = 0.0
    for a, b in zip(I_test, I_ref):
        ref = set(int(x) for x in b if x >= 0)
        if not ref:
            tot += 1.0
            continue
        tot += len(ref & set(int(x) for x in a if x >= 0)) / len(ref)
    return tot / max(1, len(I_ref))


def compare_topk(D_test: np.ndarray, I_test: np.ndarray, D_ref: np.ndarray, I_ref: np.ndarray,
                 X: np.ndarray, Q: np.ndarray, tie_tol: float = 1e-3) -> dict:
    """Apply the parity rule.

    Ids must equal the oracle's.  After removing common ids, an id only we
    returned (and symmetrically an id only the oracle returned) is accepted iff
    its fp32 score (recomputed here in fp64 from X, Q) is within ``tie_tol`` of
    the oracle's k-th score.  ORDER is checked as well, position by position: where our id at
    rank j differs from the oracle's, its true score must be within ``2 * tie_tol`` of the oracle's
    j-th score (a storage error of eps per row moves the j-th order statistic by at most eps and
    the row's own score by another eps), and the true scores of our list may never increase by
    more than ``2 * tie_tol`` from one rank to the next.  Returns counters; ``ok`` is the verdict.
    """
    X = _f32(X)
    Q = _f32(Q)
    nq, k = I_ref.shape
    exact_order = set_match = swaps = bad = order_bad = 0
    max_err = 0.0
    bad_examples = []
    for i in range(nq):
        a, b = I_test[i], I_ref[i]
        if np.array_equal(a, b):
            exact_order += 1
        sa = set(int(x) for x in a if x >= 0)
        sb = set(int(x) for x in b if x >= 0)
        if sa == sb:
            set_match += 1
        valid = b >= 0
        if len(sa) != len(sb) or (a >= 0).sum() != valid.sum():
            bad += 1
            bad_examples.append((i, "count"))
            continue
        if valid.any():
            kth = float(D_ref[i][valid][-1])
            prev = None
            for j in np.nonzero(a >= 0)[0]:
                s = float(np.dot(Q[i].astype(np.float64), X[int(a[j])].astype(np.float64)))
                max_err = max(max_err, abs(s - float(D_test[i, j])))
                if (a[j] != b[j] and abs(s - float(D_ref[i, j])) > 2 * tie_tol) or \
                        (prev is not None and s > prev + 2 * tie_tol):
                    order_bad += 1
                    bad_examples.append((i, "order", int(j), int(a[j]), s, float(D_ref[i, j])))
                prev = s
            for x in (sa - sb) | (sb - sa):
                s = float(np.dot(Q[i].astype(np.float64), X[x].astype(np.float64)))
                if abs(s - kth) <= tie_tol:
                    swaps += 1
                else:
                    bad += 1
                    bad_examples.append((i, x, s, kth))
    return {"nq": nq, "k": k, "exact_order": exact_order, "set_match": set_match,
            "tie_swaps": swaps, "violations": bad, "order_violations": order_bad, "max_abs_score_err": max_err,
            "ok": bad == 0 and order_bad == 0, "examples": bad_examples[:5]}
