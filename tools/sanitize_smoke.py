#!/usr/bin/env python3
"""Small end-to-end pass over every kernel family, meant to run under `compute-sanitizer --tool memcheck`
(sizes are tiny: the sanitizer slows kernels down by orders of magnitude)."""
import ctypes
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import semantic_search_kd_b200 as pkg  # noqa: E402


def unit(n, seed):
    a = np.random.default_rng(seed).standard_normal((n, 384)).astype(np.float32)
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def main():
    X, Q = unit(3001, 0), unit(300, 1)
    ref = (Q @ X.T)
    for opts in ({"path": 1}, {"path": 2}, {"path": 2, "tc_single_cta": 0}, {"path": 2, "seed": 0}, {"path": 2, "tc_shared_thr": 0}):
        idx = pkg.FlatIPIndex(384, metric="inner_product")
        for k_, v in opts.items():
            idx.set_option(k_, v)
        idx.add(X)
        for nq, k in ((1, 10), (2, 40), (5, 10), (130, 10), (300, 100), (3, 600)):
            D, I = idx.search(Q[:nq], k)
            top1 = ref[:nq].argmax(axis=1)
            assert (I[:, 0] == top1).mean() > 0.9, (opts, nq, k)
        idx.close()
    # fp32 re-ranking, cosine, MaxSim, ANCE filter, exchange with itself
    idx = pkg.FlatIPIndex(384, metric="inner_product", keep_fp32=True)
    idx.add(X)
    D, I = idx.search(Q[:7], 10)
    assert (I[:, 0] == ref[:7].argmax(axis=1)).all()
    S, Dd = pkg.maxsim_topk(idx, Q[:9], 5, np.arange(len(X)) // 3)
    ids, sc, cnt = pkg.ANCEMiner(None, margin=0.05).mine_corpus(idx, Q[:9], [[int(i)] for i in ref[:9].argmax(axis=1)], top_k=20)
    idx.close()
    L = pkg._lib.lib()
    idx = pkg.FlatIPIndex(384, metric="inner_product")     # (the peer exchange refuses fp32 re-ranking)
    idx.add(X)
    assert L.b2s_exchange_create(idx._h, 1, 0, 1 << 20, 4096, None) == 0
    ptrs = (ctypes.c_void_p * 1)(L.b2s_exchange_local(idx._h))
    assert L.b2s_exchange_connect(idx._h, ptrs, 1) == 0
    D2 = np.empty((4, 10), np.float32)
    I2 = np.empty((4, 10), np.int64)
    for nq in (1, 4):
        assert L.b2s_search_sharded(idx._h, Q.ctypes.data_as(ctypes.c_void_p), nq, 10, D2.ctypes.data_as(ctypes.c_void_p),
                                    I2.ctypes.data_as(ctypes.c_void_p)) == 0
        assert (I2[:nq, 0] == ref[:nq].argmax(axis=1)).all()
    idx.close()
    # cascade select + dynamic tail + overlapped (stable-query) launches + trace, on the smallest shard that takes them
    Xc = unit(140_000, 3)
    refc = Q[:12] @ Xc.T
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    idx.set_option("cascade_min_units", 4)
    idx.add(Xc)
    dev = torch.device("cuda", 0)
    Qd = torch.from_numpy(Q[:12]).to(dev)
    for tm in (0, 1):
        idx.set_option("transition_mode", tm)
        idx.set_option("trace", tm)
        outs = [idx.search_device(Qd[i:i + 1], (10, 16, 3)[i % 3], stable_queries=True) for i in range(12)]
        torch.cuda.synchronize()
        for i, (s_, i_) in enumerate(outs):
            assert int(i_[0, 0]) == int(refc[i].argmax()), (tm, i)
    idx.set_option("trace", 0)
    torch.cuda.synchronize()
    print("SANITIZE_SMOKE_OK")


if __name__ == "__main__":
    main()
