"""GPU tests of the fused candidate exchange (csrc/exchange.cuh) behind b2s_search_sharded_device.

On ONE GPU the ranks of a row-sharded corpus are emulated as several index handles of one process
whose exchange buffers are connected by raw pointers; kernels that wait on other ranks must not be
co-scheduled on one device, so every rank first runs phase 1 (local search + push) and then phase 2
(wait + merge).  With >= 2 GPUs the real thing (one process per GPU, CUDA IPC, phase 0 = the fused
kernel) runs under torchrun: tests/mp_sharded_gpu.py."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import unit_rows

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _emulated_ranks(pkg, X, world, path):
    import torch
    from semantic_search_kd_b200.sharded import shard_range
    L = pkg._lib.lib()
    ranks = []
    for r in range(world):
        lo, hi = shard_range(len(X), world, r)
        idx = pkg.FlatIPIndex(X.shape[1], metric="inner_product", device=0)
        idx.set_option("path", path)
        idx.add(X[lo:hi]) if hi > lo else idx._ensure()
        idx.set_id_offset(lo)
        rc = L.b2s_exchange_create(idx._h, world, r, 1 << 20, 4096, None)
        idx.set_option("exchange_timeout_ms", 2000)
        assert rc == 0, pkg._lib.last_error()
        ranks.append(idx)
    ptrs = (ctypes.c_void_p * world)(*[L.b2s_exchange_local(i._h) for i in ranks])
    for idx in ranks:
        assert L.b2s_exchange_connect(idx._h, ptrs, 1) == 0, pkg._lib.last_error()
    return ranks


@pytest.mark.parametrize("world,n,nq,k,path", [(2, 20000, 1, 10, 1), (3, 20001, 5, 10, 1), (4, 30000, 70, 100, 2),
                                               (8, 5000, 200, 10, 2), (3, 2, 4, 10, 1), (2, 9000, 300, 7, 0),
                                               (8, 40000, 3, 1000, 2), (5, 30000, 2, 1000, 0)])   # world * k > 4096: run merge by binary search
def test_emulated_ranks_two_phase(oracle, world, n, nq, k, path):
    import torch
    import semantic_search_kd_b200 as pkg
    L = pkg._lib.lib()
    X, Q = unit_rows(n, 384, 100 + n % 97), unit_rows(nq, 384, 200 + nq)
    if n > 100:
        X[n // 2 + 1] = X[3]            # an exact tie that straddles shards
    ranks = _emulated_ranks(pkg, X, world, path)
    dev = torch.device("cuda", 0)
    q = torch.from_numpy(Q).to(dev)
    outs = [(torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
            for _ in ranks]
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for rep in range(3):                 # several calls: sequence numbers and both slot parities
        for phase in (1, 2):
            for idx, (s, i) in zip(ranks, outs):
                rc = L.b2s_search_sharded_device(idx._h, ctypes.c_void_p(q.data_ptr()), 0, nq, k,
                                                 ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()), stream, phase, 0)
                assert rc == 0, pkg._lib.last_error()
        torch.cuda.synchronize()
        for idx in ranks:
            assert L.b2s_exchange_status(idx._h) == 0
        Dr, Ir = oracle.flat_ip_topk(X, Q, k)
        for s, i in outs:
            D, I = s.cpu().numpy(), i.cpu().numpy()
            rep_ = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=1e-3)
            assert rep_["ok"], rep_
            assert np.array_equal(I, outs[0][1].cpu().numpy())      # every rank holds the same answer
    if n > 100:
        D, I = outs[0][0].cpu().numpy(), outs[0][1].cpu().numpy()
    for idx in ranks:
        idx.close()


def test_world_one_fused(oracle):
    """world = 1: the fused kernel (push to itself, wait on itself, merge) end to end, host API included."""
    import torch
    import semantic_search_kd_b200 as pkg
    L = pkg._lib.lib()
    X, Q = unit_rows(15000, 384, 5), unit_rows(9, 384, 6)
    (idx,) = _emulated_ranks(pkg, X, 1, 0)
    Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
    D = np.empty((9, 10), np.float32)
    I = np.empty((9, 10), np.int64)
    for _ in range(3):
        rc = L.b2s_search_sharded(idx._h, Q.ctypes.data_as(ctypes.c_void_p), 9, 10, D.ctypes.data_as(ctypes.c_void_p),
                                  I.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, pkg._lib.last_error()
        assert oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=1e-3)["ok"]
    assert L.b2s_exchange_status(idx._h) == 0
    idx.close()


def test_exchange_rejects_oversized_calls():
    import semantic_search_kd_b200 as pkg
    L = pkg._lib.lib()
    X = unit_rows(1000, 384, 1)
    (idx,) = _emulated_ranks(pkg, X, 1, 0)
    rc = L.b2s_search_sharded_device(idx._h, ctypes.c_void_p(1), 0, 5000, 10, ctypes.c_void_p(1), ctypes.c_void_p(1), None, 0, 0)
    assert rc == pkg._lib.B2S_ERR_UNSUPPORTED
    idx.close()


def test_multi_gpu_processes():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "mp_sharded_gpu.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_OK" in r.stdout


def test_missing_peer_fails_the_call_instead_of_hanging():
    """A rank whose peer never issues the call must get an ERROR after `exchange_timeout_ms`, not a hang and not
    silent garbage: the kernel stores the call's sequence number into a mapped host word, the
    host-buffer call that hit it returns B2S_ERR_CUDA, and so does every later sharded call on the handle."""
    import time
    import semantic_search_kd_b200 as pkg
    L = pkg._lib.lib()
    X, Q = unit_rows(5000, 384, 3), unit_rows(2, 384, 4)
    ranks = _emulated_ranks(pkg, X, 2, 1)
    for idx in ranks:
        idx.set_option("exchange_timeout_ms", 50)
    D = np.empty((1, 10), np.float32)
    I = np.empty((1, 10), np.int64)
    args = (Q.ctypes.data_as(ctypes.c_void_p), 1, 10, D.ctypes.data_as(ctypes.c_void_p), I.ctypes.data_as(ctypes.c_void_p))
    t0 = time.perf_counter()
    rc = L.b2s_search_sharded(ranks[0]._h, *args)          # rank 1 never calls
    dt = time.perf_counter() - t0
    assert rc == pkg._lib.B2S_ERR_CUDA, (rc, pkg._lib.last_error())
    assert "timed out" in pkg._lib.last_error()
    assert 0.02 < dt < 5.0, dt                             # one deadline for the whole call
    assert L.b2s_exchange_status(ranks[0]._h) != 0
    assert L.b2s_search_sharded(ranks[0]._h, *args) == pkg._lib.B2S_ERR_CUDA   # the handle stays failed: no silent reuse
    assert L.b2s_search(ranks[0]._h, *args) == 0            # the plain (unsharded) search of the handle still works
    for idx in ranks:
        idx.close()
