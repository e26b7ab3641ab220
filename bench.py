#!/usr/bin/env python3
"""bench.py -- headline benchmark of the exact top-k hot path (BASELINE.json).

Metric: queries/sec, exact top-10 over an 8.8M x 384 bf16 corpus (BASELINE.json configs[1],
batch-1 serving search).  One "step" = one query = one pass of the hot path over the corpus.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

The line printed by rank 0 follows the driver's contract: `value` = whole-job queries/s with
queries resident in HBM (device API, CUDA events, max over ranks), `e2e` = the same through the
host-buffer API (`FAISSIndexBuilder.search(np.ndarray, k)` -> b2s_search; H2D + D2H inside),
`roofline` for the dominant kernel (K1 scan) timed live with CUDA events inside the library,
`cpu_baseline` = the CPU oracle port on a bounded sample, `clocks` from nvidia-smi during the
timed region.  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ROWS = 8_841_823          # MS MARCO passage count (docs/reference/api-reference.md:338 of the reference)
DIM = 384
K = 10
METRIC = "queries/sec exact top-10 over 8.8Mx384"
BLOCK = 1 << 20             # synthetic corpus is generated in 1 Mi-row blocks, seed = (seed, block)
CPU_SAMPLE_DIV = 8          # the CPU legs scan 1/8 of the rows and scale the time by 8


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU legs (oracle port) -- the only place bench.py touches oracle/
# --------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to
    its workers, which would silently turn the reference arm into a single-core run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_flat_qps(n_queries: int, warmup: int = 1):
    """Batch-1 exact flat search on the host cores over a 1/8 row sample; returns (qps_full, info)."""
    from oracle import oracle as orc
    orc.build()
    rows = N_ROWS // CPU_SAMPLE_DIV
    X = orc.gen_unit_rows(rows, DIM, 0)
    Q = orc.gen_unit_rows(max(n_queries, 1) + warmup, DIM, 1)
    nt = host_threads()
    for i in range(warmup):
        orc.flat_ip_topk(X, Q[i:i + 1], K, acc="f32", nthreads=nt)
    t0 = time.perf_counter()
    for i in range(n_queries):
        orc.flat_ip_topk(X, Q[warmup + i:warmup + i + 1], K, acc="f32", nthreads=nt)
    dt = time.perf_counter() - t0
    per_query_full = dt / max(1, n_queries) * CPU_SAMPLE_DIV
    info = {"cores": nt, "kind": "port",
            "sample": f"{n_queries} batch-1 queries, fp32 flat scan (oracle/flat_ip.c, OpenMP) over "
                      f"{rows} of {N_ROWS} rows (1/{CPU_SAMPLE_DIV} sample); per-query time x{CPU_SAMPLE_DIV}",
            "ms_per_query_full_corpus": per_query_full * 1e3}
    return 1.0 / per_query_full, info


def cfg0_data(kind: str, n=100_000, nq=1000):
    """BASELINE configs[0] inputs.  'isotropic' = i.i.d. Gaussian directions (the literal synthetic recipe,
    tests/conftest.py:66-73 of the reference -- the worst case for any graph index: every point is almost
    equidistant from every other); 'clustered' = a 48-dimensional latent mixture embedded in 384-d plus small
    noise, queries = perturbed corpus rows (closer to what a text encoder produces)."""
    from oracle import oracle as orc
    if kind == "isotropic":
        return orc.gen_unit_rows(n, DIM, 0), orc.gen_unit_rows(nq, DIM, 1)
    rng = np.random.default_rng(1234)
    proj = rng.standard_normal((48, DIM)).astype(np.float32) / np.sqrt(48)
    centers = rng.standard_normal((256, 48)).astype(np.float32)
    z = centers[rng.integers(0, 256, n)] + 0.6 * rng.standard_normal((n, 48)).astype(np.float32)
    X = z @ proj + 0.05 * rng.standard_normal((n, DIM)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Q = X[rng.integers(0, n, nq)] + 0.05 * rng.standard_normal((nq, DIM)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    return np.ascontiguousarray(X, dtype=np.float32), np.ascontiguousarray(Q, dtype=np.float32)


def cpu_hnsw_report(kind="isotropic", n=100_000, nq=1000):
    """BASELINE configs[0]: HNSW(32/200/64) restatement build + search + recall@10 vs flat, CPU."""
    from oracle import oracle as orc
    X, Q = cfg0_data(kind, n, nq)
    t0 = time.perf_counter()
    nt = host_threads()
    h = orc.HnswRef(X, 32, 200, nthreads=nt)
    tb = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, Ih = h.search(Q, K, 64, nthreads=nt)
    ts = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, If = orc.flat_ip_topk(X, Q, K, acc="f32", nthreads=nt)
    tf = time.perf_counter() - t0
    return {"data": kind, "n": n, "nq": nq, "M": 32, "efConstruction": 200, "efSearch": 64, "build_s": round(tb, 2),
            "hnsw_qps": round(nq / ts, 1), "flat_qps_batched": round(nq / tf, 1),
            "recall_at_10_vs_flat": round(orc.recall_at_k(Ih, If), 4), "cores": nt,
            "note": "HNSW restatement (oracle/hnsw.cpp) with the reference's parameters (src/config.py:126-139), not faiss"}


def gpu_cfg0_report(torch, pkg, dev, kinds=("clustered", "isotropic")):
    """Our path on BASELINE configs[0] (100k x 384 fp32 rows, 1k queries, k=10): queries/s through the host
    API and recall@10 against the CPU flat oracle (the checker) -- exact search, so 1.0 up to bf16 near-ties."""
    from oracle import oracle as orc
    out = []
    for kind in kinds:
        X, Q = cfg0_data(kind)
        idx = pkg.FlatIPIndex(DIM, metric="inner_product", device=dev.index)
        t0 = time.perf_counter()
        idx.add(X)
        torch.cuda.synchronize()
        tb = time.perf_counter() - t0
        idx.search(Q[:8], K)
        idx.search(Q, K)                              # warm: workspaces, tensor maps
        t0 = time.perf_counter()
        D, I = idx.search(Q, K)                       # one batched call (tensor path)
        t_batch = time.perf_counter() - t0
        t0 = time.perf_counter()
        for i in range(200):
            idx.search(Q[i:i + 1], K)                 # serving style, one query per call (scan path)
        t_one = (time.perf_counter() - t0) / 200
        Df, If = orc.flat_ip_topk(X, Q, K, acc="f32")
        rep = orc.compare_topk(D, I, Df, If, X, Q, tie_tol=1e-3)
        exact = pkg.FlatIPIndex(DIM, metric="inner_product", device=dev.index, keep_fp32=True)   # + fp32 re-ranking
        exact.add(X)
        _, Ie = exact.search(Q, K)
        exact.close()
        out.append({"data": kind, "n": len(X), "nq": len(Q), "build_s": round(tb, 3),
                    "qps_one_batch_of_1000": round(len(Q) / t_batch, 1), "qps_batch1": round(1.0 / t_one, 1),
                    "recall_at_10_vs_flat": round(orc.recall_at_k(I, If), 4),
                    "recall_at_10_vs_flat_keep_fp32": round(orc.recall_at_k(Ie, If), 4),
                    "ids_ok_under_parity_rule": rep["ok"],
                    "tie_swaps": rep["tie_swaps"]})
        idx.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())   # before the OpenMP runtime of the oracle starts
    steps = max(1, args.steps)
    qps, info = cpu_flat_qps(steps, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 / qps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": dict(info, value=qps, unit="queries/s"),
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_hnsw and args.gpus == 1:
        # BASELINE configs[0], once per round (the N=1 run): the reference's HNSW configuration on the host cores
        line["hnsw_cfg0"] = []
        for kind in ("clustered", "isotropic"):   # the clustered set first: it is the one that says something about a graph index
            try:
                line["hnsw_cfg0"].append(cpu_hnsw_report(kind))
            except Exception as e:  # never lose the main number
                line["hnsw_cfg0"].append({"data": kind, "error": str(e)})
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: 8,841,823 x 384 bf16 corpus, batch-1 serving search, exact top-10",
            "rows": N_ROWS, "dim": DIM, "k": K, "batch": 1,
            "parallelism": f"corpus row-sharded x{n_gpus}" + (", k candidates per query exchanged over NVLink peer memory "
                                                              "inside the merge kernel" if n_gpus > 1 else ""),
            "l2": "inputs larger than L2: every step streams the whole bf16 shard "
                  f"({N_ROWS * DIM * 2 / n_gpus / 1e9:.2f} GB per GPU vs 126 MB L2); 1024 distinct queries cycled"}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def make_rows(torch, lo, hi, device, seed=0):
    """Yield fp32 unit-norm row blocks covering global rows [lo, hi); block b is seeded by (seed, b)."""
    b0, b1 = lo // BLOCK, (hi - 1) // BLOCK if hi > lo else -1
    for b in range(b0, b1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + b)
        x = torch.randn((BLOCK, DIM), generator=g, device=device, dtype=torch.float32)
        x = x / x.norm(dim=1, keepdim=True)
        s, e = max(lo, b * BLOCK) - b * BLOCK, min(hi, (b + 1) * BLOCK) - b * BLOCK
        yield x[s:e].contiguous()


def batched_report(torch, index, dev, nq=1024, reps=5):
    """The tensor path (K2: TMA + tcgen05 pair MMA + fused select) on the same corpus: nq queries per
    call, exact top-10 -- reported against the measured dense bf16 peak (not part of `value`)."""
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    Q = torch.randn((nq, DIM), generator=g, device=dev, dtype=torch.float32)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
    for _ in range(2):
        index.search_device(Q, K)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        index.search_device(Q, K)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = index.stats()
    flops = 2.0 * nq * index.ntotal * DIM
    peak = None
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("bf16_tflops", 0)) or None
    peak = peak or 1590.0
    tf = flops / (best * 1e-3) / 1e12
    return {"workload": f"{nq} queries per call, exact top-{K}, same corpus", "kernel": "gemm_topk_kernel (K2)",
            "path": st["path"], "ms_per_call": best, "queries_per_s": nq / best * 1e3, "achieved_tflops": tf,
            "peak_tflops": peak, "frac": tf / peak, "bound": "tensor",
            "note": "whole call (query prep + threshold pre-pass + main pass + merge), best of %d" % reps}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import semantic_search_kd_b200 as pkg
    from semantic_search_kd_b200.sharded import ShardedFlatIPIndex, shard_range
    sys.path.insert(0, str(ROOT / "tools"))
    import bench_extras as bx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    # ---- build the (sharded) index --------------------------------------------------------
    lo, hi = shard_range(args.rows, world, rank)
    local = pkg.FlatIPIndex(DIM, metric="inner_product", device=local_rank)
    if args.path:
        local.set_option("path", args.path)
    for o in args.opt:
        name, val = o.split("=")
        local.set_option(name, int(val))
    local.reserve(hi - lo)
    t_build = time.perf_counter()
    for blk in make_rows(torch, lo, hi, dev):
        local.add(blk)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    if world > 1:
        idx = ShardedFlatIPIndex(DIM, metric="inner_product", local_index=local, exchange=args.exchange,
                                 exchange_slot_bytes=16 << 20)
        idx.local.set_id_offset(lo)
        idx.n_total, idx.range = args.rows, (lo, hi)
    else:
        idx = local

    # ---- queries: 1024 distinct, written once: resident in HBM (value) and in host memory (e2e) ----
    nqd = 1024
    g = torch.Generator(device=dev)
    g.manual_seed(1_000_003)
    Qd = torch.randn((nqd, DIM), generator=g, device=dev, dtype=torch.float32)
    Qd = (Qd / Qd.norm(dim=1, keepdim=True)).contiguous()
    Qh = Qd.cpu().numpy()
    torch.cuda.synchronize()

    # The query set is complete before the first search is enqueued and is never written again: exactly the
    # promise of B2S_SEARCH_STABLE_QUERIES (include/b200search.h), so the scan of query i+1 may overlap the
    # candidate merge / NVLink exchange of query i.  `--no-stable` times the same loop without the promise.
    stable = not args.no_stable

    def step_dev(i, st=stable):
        return idx.search_device(Qd[i % nqd:i % nqd + 1], K, stable_queries=st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_loop(steps, st):
        for i in range(max(3, min(args.warmup, 50))):
            step_dev(i, st)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step_dev(args.warmup + i, st)
        b.record()
        barrier()
        return a.elapsed_time(b)

    local.set_option("timing", 0)
    # set-up, not warm-up: the first searches allocate the handle's workspaces, connect the peer exchange (CUDA IPC)
    # and bring the GPU out of its idle clocks; then the W warm-up steps the caller asked for
    SETUP_SEARCHES = 32
    for i in range(SETUP_SEARCHES):
        step_dev(i)
    barrier()
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    barrier()
    # ---- the timed region: exactly K steps, one kernel launch each, CUDA events on the launching stream ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # The ranks leave the host barrier tens of microseconds apart -- visible when K = 20 steps are 2.5 ms.  One
        # untimed search is enqueued in front of the start event: its candidate exchange makes every GPU reach the
        # event within the same few microseconds.  The event sits between that kernel and step 1, so nothing of
        # it overlaps the timed region (an event between two kernels switches their overlap off).
        step_dev(args.warmup - 1)
    e0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    st = local.stats()
    exchange = getattr(idx, "exchange", None) if world > 1 else None
    launches_per_step = st["kernel_launches"] + (1 if exchange == "nccl" else 0)
    clk = clocks.stop() if rank == 0 else None

    # ---- the same K steps without the stable-queries promise (informational) ----
    other_ms = timed_loop(args.steps, not stable)

    # ---- the dominant kernel alone: CUDA events recorded by the library around every launch of a separate loop
    # (an event between two kernels switches their overlap off, so this is NOT done inside the timed region) ----
    iso_n = max(8, min(64, args.steps))
    local.set_option("timing", 1)   # resets the ring
    for i in range(iso_n):
        step_dev(i, False)
    barrier()
    dom = np.zeros(iso_n, np.float32)
    tot = np.zeros(iso_n, np.float32)
    got = pkg._lib.lib().b2s_read_timings(local._h, dom.ctypes.data_as(ctypes.c_void_p),
                                          tot.ctypes.data_as(ctypes.c_void_p), iso_n)
    dom = dom[:got][dom[:got] > 0]
    tot = tot[:got][tot[:got] > 0]
    k1_iso_ms = float(np.median(dom)) if len(dom) else float("nan")
    local.set_option("timing", 0)

    # ---- where a step's time goes: the kernel's own %globaltimer stamps ----
    try:
        breakdown = bx.tail_breakdown(torch, dist, idx, local, Qd, K, world, rank, n=16)
    except Exception as e:  # never lose the main number
        breakdown = {"error": str(e)}

    # ---- e2e: host buffers through the public search(), copies inside the timed region -------
    e2e_steps = max(10, min(args.steps, args.e2e_steps))
    for i in range(3):
        idx.search(Qh[i:i + 1], K)
    barrier()
    lat = np.zeros(e2e_steps)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        t1 = time.perf_counter()
        s_h, i_h = idx.search(Qh[(7 + i) % nqd:(7 + i) % nqd + 1], K)
        lat[i] = time.perf_counter() - t1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_s, k1_iso_ms, other_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, k1_iso_ms, other_ms = [float(x) for x in t.tolist()]

    # ---- parity (untimed): device path vs a torch fp32 reference over the same rows; host path == device path ----
    try:
        parity = bx.verify_parity(torch, dist, idx, make_rows, args.rows, DIM, world, rank, dev, lo, hi)
    except Exception as e:
        parity = {"ok": False, "error": repr(e)}
    sd, idd = idx.search_device(Qd[5:6], K)
    sh, ih = idx.search(Qh[5:6], K)
    agree = bool(np.array_equal(idd.cpu().numpy(), ih))
    parity["host_api_equals_device_api"] = agree
    parity["ok"] = bool(parity.get("ok")) and agree

    ex_status = idx.exchange_status() if world > 1 and hasattr(idx, "exchange_status") else 0
    exchange = getattr(idx, "exchange", None) if world > 1 else None   # may have fallen back to nccl
    extras = {}
    if (n_gpus == 8 or args.extras) and not args.no_extras:
        # BASELINE configs[2] and configs[3], bounded (the driver's 8-GPU run makes them visible)
        try:
            extras["cfg_ance"] = bx.cfg_ance_report(torch, dist, pkg, idx, local, DIM, args.rows, world, rank, local_rank, dev,
                                                    nq_total=args.ance_queries)
        except Exception as e:
            extras["cfg_ance"] = {"error": repr(e)}
        try:
            # the same sweep with the corpus replicated (6.8 GB per GPU) and the QUERIES sharded: no exchange
            full = pkg.FlatIPIndex(DIM, metric="inner_product", device=local_rank)
            full.reserve(args.rows)
            for blk in make_rows(torch, 0, args.rows, dev):
                full.add(blk)
            torch.cuda.synchronize()
            extras["cfg_ance_query_sharded"] = bx.cfg_ance_report(torch, dist, pkg, full, full, DIM, args.rows, world, rank,
                                                                  local_rank, dev, nq_total=args.ance_queries * max(1, world // 2),
                                                                  mode="query_sharded")
            full.close()
            del full
            torch.cuda.empty_cache()
        except Exception as e:
            extras["cfg_ance_query_sharded"] = {"error": repr(e)}
        try:
            extras["cfg_100m"] = bx.cfg_100m_report(torch, dist, pkg, make_rows, DIM, world, rank, local_rank, dev,
                                                    rows=args.rows_100m)
        except Exception as e:
            extras["cfg_100m"] = {"error": repr(e)}
    if rank == 0:
        peak, peak_src = peaks()
        local_bytes = (hi - lo) * DIM * 2   # rank 0's shard (ranges differ by at most one row)
        step_ms = ms / args.steps
        k_launch_ms = step_ms / max(1, st["kernel_launches"])
        achieved = local_bytes / (k_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                shard = tj.get("shard_%d_rows" % (hi - lo))   # the capture of this very shard size, if there is one
                if n_gpus == 1 and args.rows == N_ROWS:
                    traffic = tj.get("scan_topk_kernel_dram_bytes_per_launch")
                elif shard:
                    traffic = shard["dram_bytes_read"] + shard["dram_bytes_write"]
                if traffic:
                    traffic_src = ("static file profiles/traffic.json (one `ncu --set full` capture of this kernel on ONE GPU holding "
                                   "a shard of this size: %s), not measured in this run" % tj.get("source"))
            except Exception:
                traffic = None
        value = args.steps / (ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(n_gpus), stable_queries=stable,
                               queries="device-resident set written before the first search and never rewritten: the loop "
                                       "passes B2S_SEARCH_STABLE_QUERIES" if stable else "no promise about the query buffer"),
                "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel (K1; its last CTA also finishes the top-k"
                                                       + (" and exchanges it with the peers over NVLink)" if n_gpus > 1 else ")"),
                             "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": local_bytes,
                             "kernel_ms_avg": k_launch_ms, "kernel_samples": args.steps,
                             "timing": "steady state: the timed region's CUDA events (launching stream) bracket exactly "
                                       f"{args.steps} back-to-back launches of this kernel, one per step; average = region / launches",
                             "isolated_kernel_ms": k1_iso_ms, "isolated_samples": int(len(dom)),
                             "isolated_frac": (local_bytes / (k1_iso_ms * 1e-3) / 1e9 / peak) if k1_iso_ms == k1_iso_ms else None,
                             "isolated_timing": "median of CUDA events the library records around each launch of a separate "
                                                "loop after the timed region (no overlap between consecutive launches)"},
                "e2e": {"value": e2e_steps / e2e_s, "unit": "queries/s", "steps": e2e_steps,
                        "h2d_bytes_per_step": DIM * 4, "d2h_bytes_per_step": K * 12,
                        "api": "FlatIPIndex.search(np.ndarray, k) -> b2s_search (host buffers)" if world == 1
                        else f"ShardedFlatIPIndex.search(np.ndarray, k) [exchange={exchange}]"},
                "latency_ms": {"e2e_p50": float(np.percentile(lat, 50) * 1e3), "e2e_p99": float(np.percentile(lat, 99) * 1e3),
                               "device_p50": float(np.percentile(tot, 50)) if len(tot) else None,
                               "device_p99": float(np.percentile(tot, 99)) if len(tot) else None,
                               "note": "e2e = wall time of one search(np.ndarray, k) call on rank 0; device = CUDA events "
                                       "around one whole search call (rank 0, separate loop)"},
                ("without_stable_queries" if stable else "with_stable_queries"):
                    {"value": args.steps / (other_ms * 1e-3), "unit": "queries/s", "ms_per_step": other_ms / args.steps,
                     "note": "NOT the headline: the same K steps " + ("without" if stable else "with") +
                             " the B2S_SEARCH_STABLE_QUERIES promise"},
                "tail_breakdown_us": breakdown,
                "parity": parity,
                "gpu_launches": launches_per_step * args.steps,
                "setup_searches": 32, "aligned_start": world > 1, "clocks": clk, "build_s": round(t_build, 2), "paths_agree": agree, "exchange_timeouts": ex_status}
        line.update(extras)
        if n_gpus == 1 and not args.no_batched:
            try:
                line["batched"] = batched_report(torch, local, dev)
            except Exception as e:
                line["batched"] = {"error": str(e)}
            try:
                line["sweep"] = bx.sweep_report(torch, local, dev, DIM)
            except Exception as e:
                line["sweep"] = {"error": repr(e)}
        if n_gpus == 1 and not args.no_cpu:
            try:
                qps, info = cpu_flat_qps(args.cpu_queries)
                line["cpu_baseline"] = dict(info, value=qps, unit="queries/s")
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)}
            try:
                line["cfg0_recall"] = gpu_cfg0_report(torch, pkg, dev)
            except Exception as e:
                line["cfg0_recall"] = {"error": str(e)}
            try:
                # the reference's HNSW configuration on the same host cores, same run (clustered cfg0 data: the
                # isotropic set takes minutes to build and is timed by `--impl reference` at N = 1)
                line["cpu_hnsw"] = cpu_hnsw_report("clustered")
            except Exception as e:
                line["cpu_hnsw"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity.get("ok"):
        raise SystemExit("bench.py: parity violated: " + json.dumps(parity)[:2000])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 scan, 2 tensor")
    ap.add_argument("--e2e-steps", type=int, default=1000)
    ap.add_argument("--cpu-queries", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hnsw", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-stable", action="store_true", help="time the loop without B2S_SEARCH_STABLE_QUERIES")
    ap.add_argument("--extras", action="store_true", help="run the bounded configs[2]/[3] legs at this N (default: N = 8 only)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--rows-100m", type=int, default=100_000_000)
    ap.add_argument("--ance-queries", type=int, default=32768)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU candidate exchange: fused peer-memory kernel (default) or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under torch.distributed.run
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
