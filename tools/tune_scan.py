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

BASE = dict(cascade=1, dynamic_tail=3, prefetch_iters=6, peek_every=0, phase_a=0, phase_a_stagger=64, transition_mode=1,
            pdl_early=1)
CONFIGS = {
    "r01_like": dict(cascade=0, dynamic_tail=0, prefetch_iters=0, pdl_early=0),
    "dyn+pf": dict(cascade=0, dynamic_tail=3, prefetch_iters=6, pdl_early=0),
    "dyn+pf_early": dict(cascade=0, dynamic_tail=3, prefetch_iters=6, pdl_early=1),
    "casc_late": dict(BASE, pdl_early=0),
    "casc_a8": dict(BASE, phase_a=8),
    "casc_a16": dict(BASE, phase_a=16),
    "casc_pf0": dict(BASE, prefetch_iters=0),
    "casc": dict(BASE),
    "casc_tm0": dict(BASE, transition_mode=0),
    "casc_tm2": dict(BASE, transition_mode=2),
    "casc_dyn2": dict(BASE, dynamic_tail=2),
    "casc_dyn4": dict(BASE, dynamic_tail=4),
    "casc_pf4": dict(BASE, prefetch_iters=4),
    "casc_pf8": dict(BASE, prefetch_iters=8),
    "casc_stag24": dict(BASE, phase_a_stagger=24),
    "casc_peek16": dict(BASE, peek_every=16),
    "casc_tm0_stag24": dict(BASE, transition_mode=0, phase_a_stagger=24),
}


def trace_medians(idx, Q, k, n=24, stable=False):
    idx.set_option("trace", 1)
    rows = []
    for i in range(n):
        # two back-to-back calls so that the traced (second) one starts behind a predecessor, like in the loop
        idx.search_device(Q[i:i + 1], k, stable_queries=stable)
        idx.search_device(Q[i + 1:i + 2], k, stable_queries=stable)
        t = idx.read_trace()
        if t:
            rows.append(t)
    idx.set_option("trace", 0)
    if not rows:
        return None
    return {key: round(float(np.median([r[key] for r in rows if key in r])), 2) for key in rows[0]}


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
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_105_228)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--configs", type=str, default=",".join(CONFIGS))
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    idx = pkg.FlatIPIndex(DIM, metric="inner_product", device=0)
    idx.reserve(args.rows)
    for blk in make_rows(torch, 0, args.rows, dev):
        idx.add(blk)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    Q = torch.randn((1024, DIM), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
    ref = None
    lines = []
    for name in args.configs.split(","):
        for o, v in CONFIGS[name].items():
            idx.set_option(o, v)
        res = {"config": name, "rows": args.rows, "k": args.k, **CONFIGS[name]}
        s, i = idx.search_device(Q[:1], args.k)
        torch.cuda.synchronize()
        if ref is None:
            ref = i.clone()
        res["ids_equal_first_config"] = bool(torch.equal(ref, i))
        for stable in (False, True):
            for w in range(50):
                idx.search_device(Q[w:w + 1], args.k, stable_queries=stable)
            torch.cuda.synchronize()
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for j in range(args.steps):
                    idx.search_device(Q[j % 1024:j % 1024 + 1], args.k, stable_queries=stable)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / args.steps * 1e3)
            res["us_per_step_stable" if stable else "us_per_step"] = round(best, 2)
        res["ideal_us_at_7.2TBs"] = round(args.rows * DIM * 2 / 7.2e12 * 1e6, 2)
        res["trace_us"] = trace_medians(idx, Q, args.k)
        if CONFIGS[name].get("cascade"):
            res["timeline"] = timeline(idx, Q, args.k)
        print(json.dumps(res), flush=True)
        lines.append(res)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines) + "\n")


if __name__ == "__main__":
    main()
