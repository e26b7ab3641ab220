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
    for nq, k in ((1, 10), (7, 10), (200, 10), (300, 100)):
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
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
