"""Run under torchrun on >= 2 GPUs: row-sharded search, peer-memory exchange vs NCCL all-gather vs oracle."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def overlapped_stable_calls(pkg, ShardedFlatIPIndex, rank, world, local, dev):
    """B2S_SEARCH_STABLE_QUERIES on a sharded index: the scan of call i+1 overlaps the NVLink exchange of call i.
    300 back-to-back calls must equal the same calls issued one at a time, on every rank, and a torch fp32
    reference (shard-local top-k, gathered, merged) on the rows the index holds."""
    from semantic_search_kd_b200.sharded import shard_range
    n_per = 700_000                                   # large enough for the cascade select
    total = n_per * world
    lo, hi = shard_range(total, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    Xl = torch.randn((hi - lo, 384), generator=g, device=dev)
    Xl = Xl / Xl.norm(dim=1, keepdim=True)
    loc = pkg.FlatIPIndex(384, metric="inner_product", device=local)
    loc.add(Xl)
    idx = ShardedFlatIPIndex(384, metric="inner_product", local_index=loc, exchange="peer")
    idx.local.set_id_offset(lo)
    idx.n_total, idx.range = total, (lo, hi)
    g.manual_seed(55)                                 # same queries on every rank
    Q = torch.randn((300, 384), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
    ks = [(10, 10, 10, 100, 1, 16)[i % 6] for i in range(300)]
    serial = []
    for i, k in enumerate(ks):
        s, ids = idx.search_device(Q[i:i + 1], k)
        torch.cuda.synchronize()
        serial.append((s.clone(), ids.clone()))
    dist.barrier()
    outs = [idx.search_device(Q[i:i + 1], k, stable_queries=True) for i, k in enumerate(ks)]
    torch.cuda.synchronize()
    assert idx._ex_ready and idx.exchange_status() == 0
    for i, ((s, ids), (rs, ri)) in enumerate(zip(outs, serial)):
        assert torch.equal(ids, ri) and torch.equal(s, rs), ("stable != serial", i, ks[i], rank)
    # torch fp32 reference on the bf16 rows, 24 of the k = 10 calls
    Xb = Xl.to(torch.bfloat16).float()
    sel = [i for i, k in enumerate(ks) if k == 10][:24]
    ls, li = torch.topk(Q[sel] @ Xb.T, 10, dim=1)
    gs = [torch.empty_like(ls) for _ in range(world)]
    gi = [torch.empty_like(li) for _ in range(world)]
    dist.all_gather(gs, ls.contiguous())
    dist.all_gather(gi, (li + lo).contiguous())
    cs, ci = torch.cat(gs, dim=1), torch.cat(gi, dim=1)
    ts, pick = torch.topk(cs, 10, dim=1)
    ref_i = torch.gather(ci, 1, pick)
    for j, i in enumerate(sel):
        assert torch.equal(outs[i][1][0], ref_i[j]), ("stable != torch reference", i, rank)
        assert torch.allclose(outs[i][0][0], ts[j], atol=2e-5)
    dist.barrier()
    loc.close()


def main():
    import semantic_search_kd_b200 as pkg
    from semantic_search_kd_b200.sharded import ShardedFlatIPIndex
    from conftest import unit_rows
    from oracle import oracle as orc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 50001
    X = unit_rows(n, 384, 11)
    X[n // 2 + 3] = X[17]
    for nq, k in ((1, 10), (7, 10), (200, 10), (300, 100), (3, 1000)):
        Q = unit_rows(nq, 384, 300 + nq)
        Dr, Ir = orc.flat_ip_topk(X, Q, k)
        res = {}
        for mode in ("peer", "nccl"):
            idx = ShardedFlatIPIndex(384, metric="inner_product", device=local, exchange=mode)
            idx.build_from_embeddings(X)
            for rep in range(3):
                D, I = idx.search(Q, k)                                   # host buffers
                s, i = idx.search_device(torch.from_numpy(Q).to(dev), k)  # device buffers
                torch.cuda.synchronize()
                assert np.array_equal(I, i.cpu().numpy()), (mode, nq, k, rep)
                rep_ = orc.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=1e-3)
                assert rep_["ok"], (mode, nq, k, rep_)
            if mode == "peer":
                assert idx._ex_ready
                assert pkg._lib.lib().b2s_exchange_status(idx.local._h) == 0
            res[mode] = I
            dist.barrier()
            idx.local.close()
        assert np.array_equal(res["peer"], res["nccl"]), (nq, k)
        # every rank holds the same answer
        t = torch.from_numpy(res["peer"]).to(dev)
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0)
    overlapped_stable_calls(pkg, ShardedFlatIPIndex, rank, world, local, dev)
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
