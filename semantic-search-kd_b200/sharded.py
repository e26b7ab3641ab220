"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, backend ``nccl``).  Rank r holds the contiguous row
range ``[r*ceil(N/G), min(N, (r+1)*ceil(N/G)))``; global id = range start + local row, so the
``doc_ids[idx]`` mapping of ``/root/reference/src/serve/app.py:303`` is unchanged.  Queries are
replicated (``dim*4`` bytes each).  A search is

    exchange="peer" (default with NCCL/CUDA): K1/K2, then ONE fused kernel per query batch that merges
      the local lists, stores the local top-k (k*12 bytes per query) straight into every peer's
      exchange buffer over NVLink (CUDA-IPC mapped memory + release flags), waits for the peers'
      flags and merges G*k candidates (csrc/exchange.cuh; ``b2s_search_sharded_device``).  No
      collective call on the data path; torch.distributed only carries the 64-byte IPC handles once.
    exchange="nccl": local top-k on every rank (K1/K2 + K3, ids already global)
      -> ONE all-gather of the packed candidate block (k*12 bytes per query per rank) over NVLink
      -> K4 merge of G*k candidates per query on every rank (ties -> lower id).

The reference has no multi-device path; this realises the "shard across
instances, fan out, merge" prose of ``docs/operations/scaling-and-performance.md:154-172``.

``local_index`` / ``merge_fn`` are injectable so that the plumbing (ranges, gather order, packing)
is testable with the ``gloo`` backend on CPU; the defaults are the CUDA implementations and there
is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib
from .errors import IndexBuildError, IndexNotBuiltError
from .index import FlatIPIndex, _check, _small_call

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None  # type: ignore
    dist = None  # type: ignore


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank`` (empty ranges allowed when ``n_total < world``)."""
    per = -(-n_total // world) if n_total > 0 else 0
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


def packed_bytes(nq: int, k: int) -> int:
    """Bytes of one rank's packed candidate block: [ids int64 nq*k][scores fp32 nq*k], 16-aligned."""
    return (nq * k * 12 + 15) // 16 * 16


class ShardedFlatIPIndex:
    """Exact top-k over a corpus row-sharded across the ranks of a process group."""

    def __init__(self, embedding_dim: int = 384, metric: str = "cosine", group=None,
                 local_index: Optional[FlatIPIndex] = None,
                 merge_fn: Optional[Callable] = None, device: Optional[int] = None,
                 exchange: str = "auto", exchange_slot_bytes: int = 4 << 20, exchange_max_nq: int = 16384,
                 shard: str = "corpus") -> None:
        if dist is None or not dist.is_initialized():
            raise IndexBuildError("torch.distributed must be initialised before ShardedFlatIPIndex")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.embedding_dim = int(embedding_dim)
        self.local = local_index if local_index is not None else FlatIPIndex(embedding_dim, metric=metric,
                                                                              device=device)
        self._merge_fn = merge_fn
        if shard not in ("corpus", "queries"):
            raise IndexBuildError("shard must be 'corpus' or 'queries'")
        # shard="queries": every rank holds the WHOLE corpus and searches its slice of the query batch; the
        # only traffic is the all-gather of the answers.  For sweeps whose corpus fits one GPU (ANCE mining
        # over 8.8M rows = 6.8 GB): no candidate exchange, no merge, same FLOPs.
        self.shard = shard
        if exchange not in ("auto", "peer", "nccl"):
            raise IndexBuildError("exchange must be 'auto', 'peer' or 'nccl'")
        # the fused peer-memory exchange needs real GPUs and the CUDA merge (not the injected test doubles)
        if exchange == "auto":
            real_gpu_path = (merge_fn is None and (local_index is None or isinstance(local_index, FlatIPIndex)) and
                             torch is not None and torch.cuda.is_available() and dist.get_backend(group) == "nccl")
            exchange = "peer" if real_gpu_path else "nccl"
        self.exchange = exchange
        self._ex_slot_bytes = int(exchange_slot_bytes)
        self._ex_max_nq = int(exchange_max_nq)
        self._ex_ready = False
        self.exchange_fallback_reason = None
        self.n_total = 0
        self.range = (0, 0)
        self._bufs = {}

    # ------------------------------------------------------------------ build
    def build_from_embeddings(self, embeddings, n_total: Optional[int] = None) -> "ShardedFlatIPIndex":
        """Every rank passes the FULL array (or only needs its slice to be valid); each keeps its range."""
        n = int(n_total if n_total is not None else embeddings.shape[0])
        if self.shard == "queries":
            self.local.build_from_embeddings(embeddings)
            self.local.set_id_offset(0)
            self.n_total, self.range = n, (0, n)
            return self
        lo, hi = shard_range(n, self.world, self.rank)
        self.local.build_from_embeddings(embeddings[lo:hi])
        self.local.set_id_offset(lo)
        self.n_total, self.range = n, (lo, hi)
        return self

    def add_local(self, local_rows, n_total: int) -> "ShardedFlatIPIndex":
        """Append this rank's own rows (generated or loaded shard-locally); ``n_total`` = global rows."""
        lo, hi = shard_range(int(n_total), self.world, self.rank)
        self.local.add(local_rows)
        if self.local.ntotal > hi - lo:
            raise IndexBuildError(f"rank {self.rank} holds {self.local.ntotal} rows, its range has {hi - lo}")
        self.local.set_id_offset(lo)
        self.n_total, self.range = int(n_total), (lo, hi)
        return self

    @property
    def ntotal(self) -> int:
        return self.n_total

    # ------------------------------------------------------------------ search
    def _buffers(self, nq: int, k: int, device):
        key = (nq, k, str(device))
        b = self._bufs.get(key)
        if b is None:
            per = packed_bytes(nq, k)
            gathered = torch.empty((self.world, per), dtype=torch.uint8, device=device)
            mine = gathered[self.rank]
            ids = mine[: nq * k * 8].view(torch.int64).view(nq, k)
            scores = mine[nq * k * 8: nq * k * 12].view(torch.float32).view(nq, k)
            b = (gathered, mine, scores, ids, None, None)
            self._bufs = {key: b}  # keep only the latest shape
        return b

    def _connect_exchange(self, device) -> None:
        """One-time set-up of the peer-memory exchange: allocate this rank's buffer, all-gather the
        64-byte CUDA IPC handles, map every peer's buffer."""
        L = _lib.lib()
        h = self.local._ensure()
        handle = (ctypes.c_ubyte * 64)()
        rc = L.b2s_exchange_create(h, self.world, self.rank, self._ex_slot_bytes, self._ex_max_nq, handle)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        allh = torch.empty((self.world * 64,), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        if rc == _lib.B2S_OK:
            buf = np.ascontiguousarray(allh.cpu().numpy())
            rc = L.b2s_exchange_connect(h, buf.ctypes.data_as(ctypes.c_void_p), 0)
        # every rank must end up on the same protocol: if any rank could not map its peers (no P2P / IPC on
        # this box) all of them use the NCCL all-gather instead
        ok = torch.tensor([1 if rc == _lib.B2S_OK else 0], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # also: every buffer is mapped before a push
        if int(ok.item()) == 1:
            self._ex_ready = True
        else:
            self.exchange = "nccl"
            self.exchange_fallback_reason = _lib.last_error() if rc != _lib.B2S_OK else "a peer rank could not map the buffers"

    def exchange_status(self) -> int:
        """0, or the sequence number of a sharded call whose wait for a peer timed out (synchronises)."""
        if not self._ex_ready:
            return 0
        return int(_lib.lib().b2s_exchange_status(self.local._h))

    def _peer_ok(self, nq: int, k: int) -> bool:
        # (fp32 re-ranking happens after a LOCAL search only: with keep_fp32 the all-gather path is used)
        return (self.exchange == "peer" and self.world > 1 and nq <= self._ex_max_nq and
                packed_bytes(nq, k) <= self._ex_slot_bytes and not getattr(self.local, "_keep_fp32", False))

    def search_device(self, q: "torch.Tensor", k: int, stable_queries: bool = False):
        """Device-resident sharded search on the current stream; returns freshly allocated CUDA tensors
        (all ranks hold the global answer).  ``stable_queries``: see ``FlatIPIndex.search_device``.
        A peer-exchange timeout of an earlier call surfaces here as ``DeviceError``."""
        if self.n_total == 0 and self.local.ntotal == 0 and self.local._h is None:
            raise IndexNotBuiltError()
        nq = q.shape[0]
        if self.shard == "queries":
            return self._search_query_sharded(q, k)
        if nq and k and self._peer_ok(nq, k) and not self._ex_ready:
            self._connect_exchange(q.device)
        if nq and k and self._peer_ok(nq, k):
            if q.dtype not in (torch.float32, torch.bfloat16):
                q = q.float()
            q = q.contiguous()
            out_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            stream = torch.cuda.current_stream(q.device).cuda_stream
            dt = _lib.DTYPE_BF16 if q.dtype == torch.bfloat16 else _lib.DTYPE_F32
            _check(_lib.lib().b2s_search_sharded_device(self.local._h, ctypes.c_void_p(q.data_ptr()), dt, nq, int(k),
                                                        ctypes.c_void_p(out_s.data_ptr()),
                                                        ctypes.c_void_p(out_i.data_ptr()), ctypes.c_void_p(stream), 0,
                                                        _lib.SEARCH_STABLE_QUERIES if stable_queries else 0),
                   "b2s_search_sharded_device")
            return out_s, out_i
        gathered, mine, scores, ids, _os, _oi = self._buffers(nq, k, q.device)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        self.local.search_device(q, k, out=(scores, ids), **({"stable_queries": True} if stable_queries else {}))
        if self.world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), mine, group=self.group)
        if self._merge_fn is not None:
            self._merge_fn(gathered, self.world, nq, k, out_s, out_i)
        else:
            stream = torch.cuda.current_stream(q.device).cuda_stream
            _check(_lib.lib().b2s_merge_packed_device(q.device.index, ctypes.c_void_p(gathered.data_ptr()),
                                                      self.world, nq, int(k), ctypes.c_void_p(out_s.data_ptr()),
                                                      ctypes.c_void_p(out_i.data_ptr()), ctypes.c_void_p(stream)),
                   "b2s_merge_packed_device")
        return out_s, out_i

    def _search_query_sharded(self, q: "torch.Tensor", k: int):
        """Rank r searches queries [r*ceil(nq/G), ...) against its full copy of the corpus; one all-gather
        of the (padded) answers gives every rank all of them."""
        nq = q.shape[0]
        per = -(-nq // self.world) if nq else 0
        lo, hi = min(nq, self.rank * per), min(nq, self.rank * per + per)
        s_all = torch.empty((self.world, per, k), dtype=torch.float32, device=q.device)
        i_all = torch.empty((self.world, per, k), dtype=torch.int64, device=q.device)
        mine_s, mine_i = s_all[self.rank], i_all[self.rank]
        if hi > lo:
            self.local.search_device(q[lo:hi].contiguous(), k, out=(mine_s[: hi - lo], mine_i[: hi - lo]))
        if self.world > 1 and per:
            dist.all_gather_into_tensor(s_all.view(-1), mine_s.reshape(-1), group=self.group)
            dist.all_gather_into_tensor(i_all.view(-1), mine_i.reshape(-1), group=self.group)
        return s_all.view(-1, k)[:nq], i_all.view(-1, k)[:nq]

    def search(self, query_emb, k: int = 10):
        """``search(query_emb, k) -> (scores, ids)``: numpy in -> numpy out (H2D / D2H inside),
        CUDA tensor in -> CUDA tensors out.  Every rank must call it with the same queries."""
        if torch is not None and isinstance(query_emb, torch.Tensor) and query_emb.is_cuda:
            return self.search_device(query_emb, k)
        q = np.ascontiguousarray(query_emb, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        k = int(k)
        peer = bool(nq and k and self.shard == "corpus" and self._peer_ok(nq, k))
        if peer and self._ex_ready:     # the serving path: nothing below is needed
            fast = _small_call(self, _lib.lib().b2s_search_sharded, "b2s_search_sharded", q, nq, k)
            if fast is not None:
                return fast
        dev = torch.device("cuda", self.local.device if self.local.device is not None else torch.cuda.current_device())
        if self.shard == "queries":
            qd = torch.from_numpy(q).to(dev) if dev.type == "cuda" else torch.from_numpy(q)
            s, i = self._search_query_sharded(qd, k)
            return s.cpu().numpy(), i.cpu().numpy()
        if peer and not self._ex_ready:
            self._connect_exchange(dev)
            peer = self._peer_ok(nq, k)      # the connect may have fallen back to the all-gather path
        if peer:
            scores = np.empty((nq, k), dtype=np.float32)
            ids = np.empty((nq, k), dtype=np.int64)
            _check(_lib.lib().b2s_search_sharded(self.local._h, q.ctypes.data_as(ctypes.c_void_p), nq, k,
                                                 scores.ctypes.data_as(ctypes.c_void_p),
                                                 ids.ctypes.data_as(ctypes.c_void_p)), "b2s_search_sharded")
            return scores, ids
        qd = torch.from_numpy(q).to(dev, non_blocking=False)
        s, i = self.search_device(qd, k)
        return s.cpu().numpy(), i.cpu().numpy()
