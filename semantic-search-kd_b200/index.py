"""Dense index wrapper: the reference's ``FAISSIndexBuilder`` surface over libb200search.

The reference's wrapper (``src/index/build_index.py``) is absent from its tree; the surface is
reconstructed from its call sites:

* ctor           ``/root/reference/scripts/build_faiss_index.py:49-53``, ``src/serve/app.py:427-429``
* build          ``scripts/build_faiss_index.py:55-62`` (``build_from_parquet`` -> obj with ``.ntotal``)
* save / load    ``scripts/build_faiss_index.py:66``; ``src/serve/app.py:430-433``; layout
                 ``tests/conftest.py:187-198`` (``index.faiss``, ``doc_ids.json``, ``texts.json``)
* search         ``src/serve/app.py:293-295`` -> ``(distances float32 [nq,k], indices int64 [nq,k])``,
                 ``-1`` ids for missing results (guard at ``app.py:300``)

Behind it there is no FAISS and no HNSW graph: ``search`` is an EXACT scan of the bf16 corpus on
one B200 (``include/b200search.h``), so every ``index_type`` the reference accepts
(``src/config.py:129``: Flat|IVF|HNSW|PQ) is served with recall 1.0.  No CPU fallback exists: if
the CUDA library or a B200 is missing, calls raise ``DeviceError``.
"""
from __future__ import annotations

import ctypes
import json
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from . import faiss_io
from .errors import DeviceError, IndexBuildError, IndexNotBuiltError, IndexNotFoundError

try:  # torch is plumbing only (device tensors, streams)
    import torch
except Exception:  # pragma: no cover
    torch = None  # type: ignore

ArrayLike = Union[np.ndarray, "torch.Tensor"]
_SIDE_CAR = "rows.bf16"
_SIDE_META = "b200_meta.json"
_METRIC_ALIASES = {"ip": "inner_product", "dot": "inner_product"}


_SMALL_Q_FLOATS = 4096      # serving-sized calls go through per-object staging arrays whose addresses are known:
_SMALL_OUT = 2048           # building three ctypes pointers per call costs ~12 us of Python, a tenth of a sharded step


def _small_call(owner: Any, fn: Any, what: str, q: np.ndarray, nq: int, k: int):
    """Host-buffer search of a serving-sized call (``fn`` = b2s_search / b2s_search_sharded) through staging arrays
    owned by ``owner``: the query is copied in, the answers are copied out, no ctypes object is created.  Returns
    None when the call is too large or another thread is inside (the caller then takes the general path)."""
    if q.size > _SMALL_Q_FLOATS or nq * k > _SMALL_OUT:
        return None
    st = owner.__dict__.get("_stage")
    if st is None:
        import threading
        hq = np.empty(_SMALL_Q_FLOATS, dtype=np.float32)
        hs = np.empty(_SMALL_OUT, dtype=np.float32)
        hi = np.empty(_SMALL_OUT, dtype=np.int64)
        st = owner.__dict__["_stage"] = (threading.Lock(), hq, hs, hi, hq.ctypes.data, hs.ctypes.data, hi.ctypes.data)
    lock, hq, hs, hi, aq, as_, ai = st
    if not lock.acquire(False):
        return None
    try:
        hq[:q.size] = q.reshape(-1)
        handle = owner._h if hasattr(owner, "_h") else owner.local._h
        _check(fn(handle, aq, nq, k, as_, ai), what)
        n = nq * k
        return hs[:n].reshape(nq, k).copy(), hi[:n].reshape(nq, k).copy()
    finally:
        lock.release()


def _check(rc: int, what: str) -> None:
    if rc == _lib.B2S_OK:
        return
    msg = f"{what}: {_lib.last_error()} (code {rc})"
    if rc == _lib.B2S_ERR_NO_DEVICE:
        raise DeviceError(msg, rc)
    if rc in (_lib.B2S_ERR_INVALID, _lib.B2S_ERR_UNSUPPORTED, _lib.B2S_ERR_NOMEM):
        raise IndexBuildError(msg)
    raise DeviceError(msg, rc)


class FlatIPIndex:
    """Exact inner-product / cosine index on one GPU (one corpus shard)."""

    def __init__(self, embedding_dim: int = 384, index_type: str = "HNSW", metric: str = "cosine",
                 device: Optional[int] = None, keep_fp32: bool = False) -> None:
        if metric in ("cosine",):
            self._metric = _lib.METRIC_COSINE
        elif metric in ("inner_product", "ip", "dot"):
            self._metric = _lib.METRIC_INNER_PRODUCT
        else:
            raise IndexBuildError(f"metric {metric!r} is not supported (cosine | inner_product)")
        self.embedding_dim = int(embedding_dim)
        self.index_type = index_type  # accepted for compatibility; the search is always exact
        self.metric = metric
        self.doc_ids: List[str] = []
        self.doc_texts: Dict[str, str] = {}
        self._device = device
        self._keep_fp32 = bool(keep_fp32)
        self._h: Optional[ctypes.c_void_p] = None
        self._id_offset = 0
        self._pending_opts: Dict[str, int] = {}

    # ------------------------------------------------------------------ handle
    def _ensure(self) -> ctypes.c_void_p:
        if self._h is None:
            L = _lib.lib()
            dev = self._device
            if dev is None:
                dev = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
            h = ctypes.c_void_p()
            _check(L.b2s_create(self.embedding_dim, self._metric, int(dev), ctypes.byref(h)), "b2s_create")
            self._h = h
            self._device = int(dev)
            if self._keep_fp32:
                _check(L.b2s_set_option(h, b"keep_f32", 1), "b2s_set_option")
            if self._id_offset:
                _check(L.b2s_set_id_offset(h, self._id_offset), "b2s_set_id_offset")
            for name, val in self._pending_opts.items():
                _check(L.b2s_set_option(h, name.encode(), int(val)), "b2s_set_option")
        return self._h

    def close(self) -> None:
        if self._h is not None:
            _lib.lib().b2s_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # best effort
        try:
            self.close()
        except Exception:
            pass

    # --------------------------------------------------------------- properties
    @property
    def ntotal(self) -> int:
        return 0 if self._h is None else int(_lib.lib().b2s_ntotal(self._h))

    @property
    def index(self) -> "FlatIPIndex":
        """The object ``build_from_parquet`` returns; exposes ``.ntotal`` like a faiss index."""
        return self

    @property
    def device(self) -> Optional[int]:
        return self._device

    def set_option(self, name: str, value: int) -> None:
        self._pending_opts[name] = int(value)
        if self._h is not None:
            _check(_lib.lib().b2s_set_option(self._h, name.encode(), int(value)), "b2s_set_option")

    def set_id_offset(self, offset: int) -> None:
        self._id_offset = int(offset)
        if self._h is not None:
            _check(_lib.lib().b2s_set_id_offset(self._h, self._id_offset), "b2s_set_id_offset")

    def reserve(self, n_rows: int) -> None:
        _check(_lib.lib().b2s_reserve(self._ensure(), int(n_rows)), "b2s_reserve")

    # --------------------------------------------------------------------- build
    def add(self, embeddings: ArrayLike, doc_ids: Optional[Sequence[str]] = None) -> None:
        """``index.add(x)`` (``tests/conftest.py:185``): append rows ``[n, dim]`` (fp32 / bf16)."""
        L = _lib.lib()
        n_before = self.ntotal
        if torch is not None and isinstance(embeddings, torch.Tensor):
            t = embeddings
            if t.dim() != 2 or t.shape[1] != self.embedding_dim:
                raise IndexBuildError(f"expected [n, {self.embedding_dim}] embeddings, got {tuple(t.shape)}",
                                      n_before)
            if t.is_cuda:
                h = self._ensure()
                if t.device.index != self._device:
                    raise IndexBuildError(f"embeddings live on cuda:{t.device.index}, index on cuda:{self._device}",
                                          n_before)
                if t.dtype not in (torch.float32, torch.bfloat16):
                    t = t.float()
                t = t.contiguous()
                torch.cuda.current_stream(t.device).synchronize()
                fn = L.b2s_add_bf16 if t.dtype == torch.bfloat16 else L.b2s_add_f32
                _check(fn(h, ctypes.c_void_p(t.data_ptr()), t.shape[0], 1), "b2s_add")
                n_added = t.shape[0]
            else:
                if t.dtype == torch.bfloat16:
                    t = t.contiguous()
                    _check(L.b2s_add_bf16(self._ensure(), ctypes.c_void_p(t.data_ptr()), t.shape[0], 0),
                           "b2s_add_bf16")
                    n_added = t.shape[0]
                else:
                    return self.add(t.float().numpy(), doc_ids)
        else:
            a = np.asarray(embeddings)
            if a.ndim != 2 or a.shape[1] != self.embedding_dim:
                raise IndexBuildError(f"expected [n, {self.embedding_dim}] embeddings, got {a.shape}", n_before)
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.shape[0]:
                _check(L.b2s_add_f32(self._ensure(), a.ctypes.data_as(ctypes.c_void_p), a.shape[0], 0),
                       "b2s_add_f32")
            else:
                self._ensure()
            n_added = a.shape[0]
        if doc_ids is not None:
            if len(doc_ids) != n_added:
                raise IndexBuildError(f"{len(doc_ids)} doc_ids for {n_added} rows", n_before)
            if len(self.doc_ids) != n_before:
                raise IndexBuildError("doc_ids were not supplied for earlier rows", n_before)
            self.doc_ids.extend(str(d) for d in doc_ids)

    def build_from_embeddings(self, embeddings: ArrayLike,
                              doc_ids: Optional[Sequence[str]] = None) -> "FlatIPIndex":
        """north_star: "build from an embedding array"."""
        if self._h is not None:
            _check(_lib.lib().b2s_reset(self._h), "b2s_reset")
        self.doc_ids = []
        self.add(embeddings, doc_ids)
        return self

    def build_from_parquet(self, model: Any, parquet_path: Path, batch_size: int = 32,
                           max_docs: Optional[int] = None, hnsw_m: int = 32,
                           hnsw_ef_construction: int = 200, text_column: str = "text",
                           id_column: Optional[str] = None) -> "FlatIPIndex":
        """``scripts/build_faiss_index.py:55-62``: parquet -> ``model.encode_documents`` -> index.

        ``hnsw_m`` / ``hnsw_ef_construction`` are accepted and ignored: there is no graph to build,
        the "build" of an exact index is a bf16 copy of the rows into HBM.
        """
        import pandas as pd  # local: only the build path needs it

        parquet_path = Path(parquet_path)
        if not parquet_path.exists():
            raise IndexNotFoundError(str(parquet_path))
        df = pd.read_parquet(parquet_path)
        if text_column not in df.columns:
            raise IndexBuildError(f"parquet has no {text_column!r} column (columns: {list(df.columns)})")
        if id_column is None:
            id_column = "doc_id" if "doc_id" in df.columns else ("chunk_id" if "chunk_id" in df.columns else None)
        if max_docs is not None:
            df = df.iloc[: int(max_docs)]
        texts = [str(t) for t in df[text_column].tolist()]
        ids = [str(x) for x in df[id_column].tolist()] if id_column else [f"doc_{i}" for i in range(len(texts))]
        if self._h is not None:
            _check(_lib.lib().b2s_reset(self._h), "b2s_reset")
        self.doc_ids = []
        self.doc_texts = {}
        done = 0
        try:
            for s in range(0, len(texts), max(1, int(batch_size))):
                chunk = texts[s:s + batch_size]
                emb = model.encode_documents(chunk)
                self.add(np.asarray(emb, dtype=np.float32), ids[s:s + batch_size])
                done += len(chunk)
        except IndexBuildError:
            raise
        except Exception as e:  # encoder failure etc.
            raise IndexBuildError(f"index build failed after {done} documents: {e}", done) from e
        for i, t in zip(ids, texts):
            self.doc_texts.setdefault(i, t)
        return self

    # ---------------------------------------------------------------- persistence
    def save(self, output_dir: Path, write_faiss: bool = True, block_rows: int = 1 << 18) -> None:
        """Write ``index.faiss`` (IndexFlatIP layout), ``doc_ids.json``, optional ``texts.json`` and a
        bf16 side-car (``rows.bf16`` + ``b200_meta.json``) that ``load`` prefers."""
        if self._h is None:
            raise IndexNotBuiltError()
        L = _lib.lib()
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        n, d = self.ntotal, self.embedding_dim

        def blocks() -> Iterable[np.ndarray]:
            for s in range(0, n, block_rows):
                m = min(block_rows, n - s)
                buf = np.empty((m, d), dtype=np.float32)
                _check(L.b2s_read_rows_f32(self._h, s, m, buf.ctypes.data_as(ctypes.c_void_p)), "b2s_read_rows_f32")
                yield buf

        with open(out / _SIDE_CAR, "wb") as side:
            def tee() -> Iterable[np.ndarray]:
                for b in blocks():
                    # exact: the rows are bf16 values held in fp32, the high 16 bits are the bf16 pattern
                    side.write((b.view(np.uint32) >> 16).astype("<u2").tobytes())
                    yield b
            if write_faiss:
                faiss_io.write_flat_ip(out / "index.faiss", tee(), n, d)
            else:
                for _ in tee():
                    pass
        (out / _SIDE_META).write_text(json.dumps({"ntotal": n, "dim": d, "metric": self.metric,
                                                  "dtype": "bf16", "format": 1}) + "\n")
        doc_ids = self.doc_ids if self.doc_ids else [f"doc_{i}" for i in range(n)]
        with open(out / "doc_ids.json", "w") as f:
            json.dump(doc_ids, f)
        if self.doc_texts:
            with open(out / "texts.json", "w") as f:
                json.dump(self.doc_texts, f)

    def load(self, index_dir: Path, block_rows: int = 1 << 18) -> "FlatIPIndex":
        """``src/serve/app.py:430-433``: restore rows and ``doc_ids`` from an index directory.

        Rows come back VERBATIM (``b2s_add_prepared``): a directory written by ``save`` holds the rows as
        the index held them (already unit-normalised under ``metric="cosine"``), so any number of
        save / load cycles is bit-stable and an ``inner_product`` index is never silently normalised.
        The bf16 side-car is used only if it belongs to this ``index.faiss`` (same row count and
        dimension, not older than it) and was written under the metric this object was constructed
        with; a mismatching metric raises ``IndexBuildError``.
        """
        d = Path(index_dir)
        if not d.exists():
            raise IndexNotFoundError(str(d))
        meta_p, side_p, faiss_p = d / _SIDE_META, d / _SIDE_CAR, d / "index.faiss"
        L = _lib.lib()
        rows_src: Optional[np.ndarray] = None
        is_bf16 = False
        use_side = meta_p.exists() and side_p.exists()
        meta: Dict[str, Any] = {}
        if use_side:
            meta = json.loads(meta_p.read_text())
            if faiss_p.exists():
                # a side-car left behind by an earlier save must not shadow a newer / different index.faiss
                try:
                    f_dim, f_n, _m = faiss_io.read_header(faiss_p)
                except ValueError:
                    f_dim, f_n = meta.get("dim"), meta.get("ntotal")   # not a flat file: the side-car decides
                stale = faiss_p.stat().st_mtime > max(side_p.stat().st_mtime, meta_p.stat().st_mtime) + 2.0
                if f_dim != meta.get("dim") or f_n != meta.get("ntotal") or stale:
                    use_side = False
        if use_side:
            if meta.get("dim") != self.embedding_dim:
                raise IndexBuildError(f"index dim {meta.get('dim')} != embedding_dim {self.embedding_dim}")
            saved_metric = meta.get("metric", self.metric)
            if _METRIC_ALIASES.get(saved_metric, saved_metric) != _METRIC_ALIASES.get(self.metric, self.metric):
                raise IndexBuildError(f"index was saved with metric {saved_metric!r}, this index uses {self.metric!r}")
            n = int(meta["ntotal"])
            if side_p.stat().st_size != n * self.embedding_dim * 2:
                raise IndexBuildError(f"{side_p} has the wrong size for {n} rows")
            rows_src = (np.memmap(side_p, dtype="<u2", mode="r", shape=(n, self.embedding_dim))
                        if n else np.zeros((0, self.embedding_dim), dtype=np.uint16))
            is_bf16 = True
        elif faiss_p.exists():
            try:
                rows_src, f_metric = faiss_io.read_flat(faiss_p)
            except ValueError as e:
                raise IndexBuildError(str(e)) from e
            if rows_src.shape[1] != self.embedding_dim:
                raise IndexBuildError(f"index dim {rows_src.shape[1]} != embedding_dim {self.embedding_dim}")
            if f_metric != 0:   # faiss METRIC_INNER_PRODUCT = 0: cosine and inner_product are both stored that way
                raise IndexBuildError(f"index.faiss has metric_type {f_metric}; only inner-product (0) files are served")
        else:
            raise IndexNotFoundError(str(faiss_p))
        h = self._ensure()
        _check(L.b2s_reset(h), "b2s_reset")
        n = rows_src.shape[0]
        if n:
            _check(L.b2s_reserve(h, n), "b2s_reserve")
        dt = _lib.DTYPE_BF16 if is_bf16 else _lib.DTYPE_F32
        for s in range(0, n, block_rows):
            blk = np.ascontiguousarray(rows_src[s:s + block_rows])
            _check(L.b2s_add_prepared(h, blk.ctypes.data_as(ctypes.c_void_p), dt, blk.shape[0], 0), "b2s_add_prepared")
        ids_p = d / "doc_ids.json"
        if ids_p.exists():
            with open(ids_p) as f:
                self.doc_ids = [str(x) for x in json.load(f)]
            if len(self.doc_ids) != n:
                raise IndexBuildError(f"doc_ids.json has {len(self.doc_ids)} ids for {n} vectors")
        else:
            self.doc_ids = [f"doc_{i}" for i in range(n)]
        return self

    # --------------------------------------------------------------------- search
    def search(self, query_emb: ArrayLike, k: int = 10) -> Tuple[ArrayLike, ArrayLike]:
        """Exact top-k.  numpy in -> numpy out (host path, copies inside); CUDA tensor in -> CUDA
        tensors out (device path on the current stream).  ``scores`` float32 ``[nq,k]`` descending,
        ``ids`` int64 ``[nq,k]``, unfilled slots ``(-FLT_MAX, -1)``."""
        if self._h is None:
            raise IndexNotBuiltError()
        k = int(k)
        if k < 0:
            raise IndexBuildError("k must be >= 0")
        L = _lib.lib()
        if torch is not None and isinstance(query_emb, torch.Tensor):
            q = query_emb
            if q.dim() == 1:
                q = q.unsqueeze(0)
            if q.dim() != 2 or q.shape[1] != self.embedding_dim:
                raise IndexBuildError(f"expected [nq, {self.embedding_dim}] queries, got {tuple(q.shape)}")
            if q.is_cuda:
                return self.search_device(q, k)
            s, i = self.search(q.float().numpy(), k)
            return torch.from_numpy(s), torch.from_numpy(i)
        q = np.asarray(query_emb)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.ndim != 2 or q.shape[1] != self.embedding_dim:
            raise IndexBuildError(f"expected [nq, {self.embedding_dim}] queries, got {q.shape}")
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        if nq and k:
            fast = _small_call(self, L.b2s_search, "b2s_search", q, nq, k)
            if fast is not None:
                return fast
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        if nq and k:
            _check(L.b2s_search(self._h, q.ctypes.data_as(ctypes.c_void_p), nq, k,
                                scores.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p)),
                   "b2s_search")
        return scores, ids

    def search_device(self, q: "torch.Tensor", k: int,
                      out: Optional[Tuple["torch.Tensor", "torch.Tensor"]] = None, stable_queries: bool = False):
        """Device-resident search on torch's current stream (no host sync, no copies).

        ``stable_queries=True`` is the caller's promise that ``q`` was completely written before the
        previous operation on this stream was enqueued (``B2S_SEARCH_STABLE_QUERIES``): a batch-1/2
        scan may then overlap the tail of the previous search."""
        if self._h is None:
            raise IndexNotBuiltError()
        if q.device.index != self._device:
            raise IndexBuildError(f"queries live on cuda:{q.device.index}, index on cuda:{self._device}")
        if q.dtype not in (torch.float32, torch.bfloat16):
            q = q.float()
        q = q.contiguous()
        nq = q.shape[0]
        if out is None:
            scores = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        else:
            scores, ids = out
        if nq and k:
            stream = torch.cuda.current_stream(q.device).cuda_stream
            dt = _lib.DTYPE_BF16 if q.dtype == torch.bfloat16 else _lib.DTYPE_F32
            _check(_lib.lib().b2s_search_device(self._h, ctypes.c_void_p(q.data_ptr()), dt, nq, int(k),
                                                ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                                                ctypes.c_void_p(stream),
                                                _lib.SEARCH_STABLE_QUERIES if stable_queries else 0), "b2s_search_device")
        return scores, ids

    def read_trace(self, raw: bool = False, previous: bool = False) -> Optional[Dict[str, Any]]:
        """Phase stamps of the last batch-1/2 scan launch (option ``trace`` = 1), in microseconds relative to
        the earliest CTA start: see ``b2s_read_trace``.  ``previous=True`` adds the launch before it under
        ``"previous"`` and ``"period_us"`` = distance between the two launches' earliest CTA starts (they may
        overlap under programmatic dependent launch).  Synchronises the device."""
        if self._h is None:
            raise IndexNotBuiltError()
        stride, half = 512, 16 + 6 * 512
        buf = np.zeros(2 * half, dtype=np.uint64)
        n = _lib.lib().b2s_read_trace(self._h, buf.ctypes.data_as(ctypes.c_void_p), len(buf))
        if n <= 0:
            return None

        def parse(b):
            if b[1] == 0 or not (0 < int(b[0]) <= stride):
                return None, 0
            g = int(b[0])
            arr = lambda i: b[16 + i * stride:16 + i * stride + g].astype(np.int64)   # noqa: E731
            starts, ends = arr(0), arr(1)
            t0 = int(starts.min())
            us = lambda v: (int(v) - t0) / 1e3   # noqa: E731
            out = {"grid": g, "cta_start_spread_us": us(starts.max()), "cta_start_p90_us": us(np.percentile(starts, 90)),
                   "scan_end_first_us": us(ends.min()),
                   "scan_end_median_us": us(np.median(ends)), "scan_end_last_us": us(ends.max()), "ticket_us": us(b[1]),
                   "local_topk_us": us(b[2]), "done_us": us(b[5])}
            tb, te = arr(2), arr(3)
            if tb.min() > 0 and te.min() > 0:
                out["transition_wait_median_us"] = float(np.median(te - tb)) / 1e3
                out["transition_wait_max_us"] = float((te - tb).max()) / 1e3
            if b[13]:
                out["phase_b_offers"], out["phase_b_inserts"], out["transition_keys"] = int(b[11]), int(b[12]), int(b[13])
            if b[3] and b[4]:
                out["pushed_us"], out["peers_seen_us"] = us(b[3]), us(b[4])
            if raw:
                out["raw"] = {"start": (starts - t0) / 1e3, "end": (ends - t0) / 1e3, "trans_begin": (tb - t0) / 1e3,
                              "trans_end": (te - t0) / 1e3, "static_end": (arr(4) - t0) / 1e3, "smid": arr(5)}
            return out, t0

        out, t0 = parse(buf[:half])
        if out is not None and previous:
            prev, p0 = parse(buf[half:])
            if prev is not None:
                out["previous"] = prev
                out["period_us"] = (t0 - p0) / 1e3
        return out

    def stats(self) -> Dict[str, Any]:
        if self._h is None:
            raise IndexNotBuiltError()
        st = _lib.Stats()
        _check(_lib.lib().b2s_last_stats(self._h, ctypes.byref(st)), "b2s_last_stats")
        return {f: getattr(st, f) for f, _ in st._fields_ if f != "reserved"}


# The name the reference imports: ``from src.index.build_index import FAISSIndexBuilder``
# (``src/serve/app.py:21``, ``scripts/build_faiss_index.py:9``).
FAISSIndexBuilder = FlatIPIndex
