"""CPU, world_size 2, gloo: the host-side plumbing of the row-sharded search (shard ranges, global
id offsets, packed candidate block, all-gather order, merge input layout).  The local search and
the merge are injected test doubles backed by the ORACLE -- the product defaults are CUDA-only."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleLocalIndex:
    """Test double with FlatIPIndex's surface; search = CPU oracle on bf16-rounded rows."""

    def __init__(self, dim):
        from oracle import oracle as orc
        self.orc = orc
        self.dim = dim
        self.X = np.zeros((0, dim), np.float32)
        self.offset = 0
        self.device = None
        self._h = object()

    @property
    def ntotal(self):
        return self.X.shape[0]

    def build_from_embeddings(self, emb, doc_ids=None):
        self.X = self.orc.round_bf16(np.asarray(emb, np.float32).reshape(-1, self.dim))
        return self

    def add(self, emb, doc_ids=None):
        self.X = np.concatenate([self.X, self.orc.round_bf16(np.asarray(emb, np.float32))])

    def set_id_offset(self, off):
        self.offset = int(off)

    def search_device(self, q, k, out=None):
        D, I = self.orc.flat_ip_topk(self.X, q.numpy(), k, id_offset=self.offset)
        scores, ids = out
        scores.copy_(torch.from_numpy(D))
        ids.copy_(torch.from_numpy(I))
        return scores, ids


def oracle_merge(gathered, world, nq, k, out_s, out_i):
    """Reference for b2s_merge_packed_device: per rank [ids int64 nq*k][scores f32 nq*k]."""
    S, I = [], []
    for r in range(world):
        blk = gathered[r]
        I.append(blk[: nq * k * 8].view(torch.int64).view(nq, k).numpy())
        S.append(blk[nq * k * 8: nq * k * 12].view(torch.float32).view(nq, k).numpy())
    S, I = np.stack(S), np.stack(I)                      # [G, nq, k]
    for q in range(nq):
        s = S[:, q, :].reshape(-1)
        i = I[:, q, :].reshape(-1)
        pos = np.arange(s.size)
        valid = i >= 0
        order = np.lexsort((pos[valid], -s[valid].astype(np.float64)))[:k]   # score desc, rank-major pos asc
        ss = np.full(k, np.float32(-3.4028234663852886e38), np.float32)
        ii = np.full(k, -1, np.int64)
        ss[: order.size] = s[valid][order]
        ii[: order.size] = i[valid][order]
        out_s[q] = torch.from_numpy(ss)
        out_i[q] = torch.from_numpy(ii)


def _worker(rank, world, port, n, nq, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import unit_rows
        from oracle import oracle as orc
        from semantic_search_kd_b200.sharded import ShardedFlatIPIndex, shard_range, packed_bytes
        X = unit_rows(n, 384, 5)
        if n > 40:
            X[n - 3:] = X[2]                 # exact ties that straddle the shards
        Q = unit_rows(nq, 384, 6)
        Q[0] = X[min(2, n - 1)]
        idx = ShardedFlatIPIndex(384, metric="inner_product", local_index=OracleLocalIndex(384),
                                 merge_fn=oracle_merge)
        idx.build_from_embeddings(X)
        lo, hi = shard_range(n, world, rank)
        assert idx.range == (lo, hi) and idx.local.ntotal == hi - lo and idx.ntotal == n
        assert packed_bytes(nq, k) % 16 == 0 and packed_bytes(nq, k) >= nq * k * 12
        s, i = idx.search_device(torch.from_numpy(Q), k)
        Dr, Ir = orc.flat_ip_topk(orc.round_bf16(X), Q, k)
        ok = np.array_equal(i.numpy(), Ir) and np.allclose(s.numpy(), Dr, atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,nq,k", [(1000, 5, 10), (7, 3, 10), (1, 2, 4), (513, 1, 100)])
def test_sharded_search_world2_gloo(n, nq, k):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, nq, k, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_ranges_cover_exactly():
    from semantic_search_kd_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 9, 8841823, 100_000_000):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(hi >= lo for lo, hi in rs)


def _worker_queries(rank, world, port, n, nq, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import unit_rows
        from oracle import oracle as orc
        from semantic_search_kd_b200.sharded import ShardedFlatIPIndex
        X, Q = unit_rows(n, 384, 15), unit_rows(nq, 384, 16)
        idx = ShardedFlatIPIndex(384, metric="inner_product", local_index=OracleLocalIndex(384),
                                 merge_fn=oracle_merge, shard="queries")
        idx.build_from_embeddings(X)
        assert idx.local.ntotal == n and idx.range == (0, n)          # the whole corpus on every rank
        s, i = idx.search_device(torch.from_numpy(Q), k)
        Dr, Ir = orc.flat_ip_topk(orc.round_bf16(X), Q, k)
        ret[rank] = bool(np.array_equal(i.numpy(), Ir) and np.allclose(s.numpy(), Dr, atol=1e-6) and s.shape == (nq, k))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,nq,k", [(800, 7, 10), (800, 1, 5), (50, 4, 10)])
def test_query_sharded_search_world2_gloo(n, nq, k):
    """shard="queries": replicated corpus, each rank searches its slice of the batch, answers all-gathered."""
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_queries, args=(world, port, n, nq, k, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
