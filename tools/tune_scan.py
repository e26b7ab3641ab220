#!/usr/bin/env python3
"""tools/tune_scan.py -- A/B the batch-1 scan kernel's options on one GPU (K1, scan_topk.cuh).

Builds a shard of --rows rows (default: the 8-GPU shard of BASELINE configs[1], 1,105,228 rows) and times
back-to-back batch-1 searches (device API, CUDA events around the whole loop) for a list of option sets,
with and without B2S_SEARCH_STABLE_QUERIES, and prints the kernel's own phase stamps (option "trace").
One JSON line per configuration; nothing here is a bench number (it is how the defaults were chosen)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import semantic_search_kd_b200 as pkg  # noqa: E402
from bench import make_rows, DIM  # noqa: E402

BASE = dict(cascade=1, dynamic_tail=3, prefetch_iters=6, peek_every=0, phase_a=0, phase_a_stagger=64, pdl_early=1, grid_spare=1, transition_mode=0)
CONFIGS = {
    "r01_like": dict(BASE, cascade=0, dynamic_tail=0, prefetch_iters=0, pdl_early=0, grid_spare=0),
    "lists": dict(BASE, cascade=0),
    "lists_late": dict(BASE, cascade=0, pdl_early=0),
    "casc": dict(BASE),
    "casc_spare0": dict(BASE, grid_spare=0),
    "casc_tm1": dict(BASE, transition_mode=1),
    "casc_tm1_spare0": dict(BASE, transition_mode=1, grid_spare=0),
    "casc_spare2": dict(BASE, grid_spare=2),
    "casc_late": dict(BASE, pdl_early=0),
    "casc_late_spare0": dict(BASE, pdl_early=0, grid_spare=0),
    "casc_dyn2": dict(BASE, dynamic_tail=2),
    "casc_dyn6": dict(BASE, dynamic_tail=6),
    "casc_dyn12": dict(BASE, dynamic_tail=12),
    "casc_a4": dict(BASE, phase_a=4),
    "casc_a12": dict(BASE, phase_a=12),
    "casc_stag16": dict(BASE, phase_a_stagger=16),
    "casc_stag32": dict(BASE, phase_a_stagger=32),
    "casc_pf0": dict(BASE, prefetch_iters=0),
    "casc_pf12": dict(BASE, prefetch_iters=12),
}


def trace_medians(idx, local, Q, k, n=24, stable=False, burst=4):
    local.set_option("trace", 1)
    rows = []
    for i in range(n):
        # a few back-to-back calls so that the traced (last) one runs behind predecessors, as in the loop
        for j in range(burst):
            idx.search_device(Q[(i + j) % 1024:(i + j) % 1024 + 1], k, stable_queries=stable)
        t = local.read_trace(previous=True)
        if t:
            t.pop("previous", None)
            rows.append(t)
    local.set_option("trace", 0)
    if not rows:
        return None
    keys = [key for key in rows[0] if all(key in r for r in rows)]
    return {key: round(float(np.median([r[key] for r in rows])), 2) for key in keys}


def timeline(idx, Q, k):
    """Per-CTA timeline of ONE traced launch: where the slow CTAs lose their time."""
    idx.set_option("trace", 1)
    idx.search_device(Q[:1], k)
    idx.search_device(Q[1:2], k)
    t = idx.read_trace(raw=True)
    idx.set_option("trace", 0)
    if not t:
        return None
    r = t["raw"]
    order = np.argsort(r["end"])
    pick = [order[0], order[len(order) // 4], order[len(order) // 2], order[3 * len(order) // 4], order[-1]]
    rows = []
    for b in pick:
        rows.append({"cta": int(b), "sm": int(r["smid"][b]), "start": round(float(r["start"][b]), 1),
                     "trans_begin": round(float(r["trans_begin"][b]), 1), "trans_end": round(float(r["trans_end"][b]), 1),
                     "static_end": round(float(r["static_end"][b]), 1), "end": round(float(r["end"][b]), 1)})
    sm_of = r["smid"]
    slow = r["end"] > np.median(r["end"]) * 1.2
    return {"quantile_ctas": rows, "n_slow": int(slow.sum()),
            "slow_sms_shared_by_two_slow": int(sum(np.sum(sm_of[slow] == s) == 2 for s in set(sm_of[slow].tolist()))),
            "slow_cta_mod32_hist": np.bincount((np.nonzero(slow)[0] % 32), minlength=32).tolist()}


def main():
    import os
    import torch.distributed as dist
    from semantic_search_kd_b200.sharded import ShardedFlatIPIndex, shard_range
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_105_228, help="rows PER GPU")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--configs", type=str, default=",".join(CONFIGS))
    ap.add_argument("--out", type=str, default="")
    ap.add_argument("--burst", type=int, default=200)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = args.rows * world
    lo, hi = shard_range(total, world, rank)
    local = pkg.FlatIPIndex(DIM, metric="inner_product", device=lr)
    local.reserve(hi - lo)
    for blk in make_rows(torch, lo, hi, dev):
        local.add(blk)
    if world > 1:
        idx = ShardedFlatIPIndex(DIM, metric="inner_product", local_index=local)
        idx.local.set_id_offset(lo)
        idx.n_total, idx.range = total, (lo, hi)
    else:
        idx = local
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    Q = torch.randn((1024, DIM), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ref = None
    lines = []
    for name in args.configs.split(","):
        for o, v in CONFIGS[name].items():
            local.set_option(o, v)
        res = {"config": name, "rows_per_gpu": args.rows, "world": world, "k": args.k, **CONFIGS[name]}
        s, i = idx.search_device(Q[:1], args.k)
        sync()
        if ref is None:
            ref = i.clone()
        res["ids_equal_first_config"] = bool(torch.equal(ref, i))
        for stable in (False, True):
            for w in range(50):
                idx.search_device(Q[w:w + 1], args.k, stable_queries=stable)
            sync()
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for j in range(args.steps):
                    idx.search_device(Q[j % 1024:j % 1024 + 1], args.k, stable_queries=stable)
                e1.record()
                sync()
                ms = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                best = min(best, ms / args.steps * 1e3)
            res["us_per_step_stable" if stable else "us_per_step"] = round(best, 2)
        res["ideal_us_at_7.2TBs"] = round(args.rows * DIM * 2 / 7.2e12 * 1e6, 2)
        res["trace_us"] = trace_medians(idx, local, Q, args.k)
        sync()
        res["trace_us_stable"] = trace_medians(idx, local, Q, args.k, stable=True)
        sync()
        res["trace_us_stable_burst"] = trace_medians(idx, local, Q, args.k, n=8, stable=True, burst=args.burst)
        sync()
        # per-CTA picture of the last launch of one long stable burst
        local.set_option("trace", 1)
        for j in range(args.burst):
            idx.search_device(Q[j % 1024:j % 1024 + 1], args.k, stable_queries=True)
        t = local.read_trace(raw=True, previous=True)
        local.set_option("trace", 0)
        sync()
        if t and rank == 0:
            r = t["raw"]
            order = np.argsort(r["start"])
            print("# late starters (cta, sm, start, trans_begin, trans_end, static_end, end):",
                  [(int(b), int(r["smid"][b]), round(float(r["start"][b]), 1), round(float(r["trans_begin"][b]), 1),
                    round(float(r["trans_end"][b]), 1), round(float(r["static_end"][b]), 1), round(float(r["end"][b]), 1))
                   for b in order[-6:]], "period", t.get("period_us"), flush=True)
            w = r["trans_end"] - r["trans_begin"]
            print("# transition wait percentiles (p10,p50,p90,max):", [round(float(np.percentile(w, q)), 1) for q in (10, 50, 90, 100)],
                  "start percentiles:", [round(float(np.percentile(r["start"], q)), 1) for q in (50, 90, 99, 100)],
                  "end percentiles:", [round(float(np.percentile(r["end"], q)), 1) for q in (1, 50, 90, 99, 100)], flush=True)
        if rank == 0:
            print(json.dumps(res), flush=True)
            keys = ("period_us", "cta_start_p90_us", "cta_start_spread_us", "scan_end_first_us", "scan_end_median_us", "scan_end_last_us",
                    "ticket_us", "local_topk_us", "pushed_us", "peers_seen_us", "done_us", "transition_wait_median_us",
                    "transition_wait_max_us", "phase_b_offers")
            for label in ("trace_us", "trace_us_stable", "trace_us_stable_burst"):
                t = res.get(label) or {}
                print("#", name, world, args.rows, res["us_per_step"], res["us_per_step_stable"], res["ids_equal_first_config"], label,
                      {k: t.get(k) for k in keys if k in t}, flush=True)
        lines.append(res)
    if args.out and rank == 0:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
