#!/usr/bin/env python3
"""tools/sweep.py -- BASELINE.json configs[4]: query-batch sweep 1 -> 4096 at k in {10,100,1000} over the
8.8M x 384 bf16 corpus on one B200; prints time per batch, queries/s and the fraction of
max(HBM, tensor) roofline (t_roof = max(N*D*2/BW, 2*nq*N*D/PEAK))."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import semantic_search_kd_b200 as pkg  # noqa: E402
from bench import make_rows, N_ROWS, DIM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--nq", type=str, default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--k", type=str, default="10,100,1000")
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--seed", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", type=str, default="")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    args = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    bw = peaks.get("hbm_gbs", 6650.0) * 1e9
    tf = peaks.get("bf16_tflops", 1590.0) * 1e12
    dev = torch.device("cuda", 0)
    idx = pkg.FlatIPIndex(DIM, metric="inner_product", device=0)
    idx.set_option("path", args.path)
    idx.set_option("seed", args.seed)
    for o in args.opt:
        name, val = o.split("=")
        idx.set_option(name, int(val))
    idx.reserve(args.rows)
    for blk in make_rows(torch, 0, args.rows, dev):
        idx.add(blk)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    Q = torch.randn((4096, DIM), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
    rows = []
    for k in [int(x) for x in args.k.split(",")]:
        for nq in [int(x) for x in args.nq.split(",")]:
            q = Q[:nq]
            out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
            for _ in range(2):
                idx.search_device(q, k, out=out)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                idx.search_device(q, k, out=out)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            st = idx.stats()
            t_h = args.rows * DIM * 2 / bw
            t_t = 2.0 * nq * args.rows * DIM / tf
            roof = max(t_h, t_t)
            r = {"k": k, "nq": nq, "ms": round(best, 4), "qps": round(nq / best * 1e3, 1), "path": st["path"],
                 "seeded": st["seeded"], "launches": st["kernel_launches"], "bound": "hbm" if t_h >= t_t else "tensor",
                 "roof_ms": round(roof * 1e3, 4), "frac": round(roof * 1e3 / best, 4)}
            rows.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
