#!/usr/bin/env python3
"""tools/run_configs.py -- BASELINE.json configs[2] and configs[3] on the GPUs of one box (torchrun).

  configs[3]: 100M x 384 bf16 corpus row-sharded across 8 B200, batch 1024, k=100, candidate merge
  configs[2]: ANCE mining sweep, 500k queries x 8.8M corpus, top-200 negatives, batched tensor path
              (corpus-sharded with the candidate exchange, and query-sharded with a replicated corpus)

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py

Every number is device time (CUDA events) max over ranks; rank 0 prints one JSON line per config.
Rows/queries scale with --scale (1.0 = the BASELINE sizes) so that the script can be smoke-tested small."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import semantic_search_kd_b200 as pkg  # noqa: E402
from semantic_search_kd_b200.sharded import ShardedFlatIPIndex, shard_range  # noqa: E402
from bench import make_rows, DIM  # noqa: E402


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0)


def unit_queries(n, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q = torch.randn((n, DIM), generator=g, device=dev, dtype=torch.float32)
    return (q / q.norm(dim=1, keepdim=True)).contiguous()


def timed(fn, reps, world, dev):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, out


def build_sharded(rows, world, rank, local_rank, dev, slot_bytes):
    lo, hi = shard_range(rows, world, rank)
    local = pkg.FlatIPIndex(DIM, metric="inner_product", device=local_rank)
    local.reserve(hi - lo)
    for blk in make_rows(torch, lo, hi, dev):
        local.add(blk)
    torch.cuda.synchronize()
    if world == 1:
        return local, local
    idx = ShardedFlatIPIndex(DIM, metric="inner_product", local_index=local, exchange_slot_bytes=slot_bytes)
    idx.local.set_id_offset(lo)
    idx.n_total, idx.range = rows, (lo, hi)
    return idx, local


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", type=str, default="")
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm, tf_burst, tf_sus = peaks()
    lines = []

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)

    # ---------------- configs[3]: 100M rows over the GPUs, batch 1024, k = 100 ----------------
    if args.only in ("", "cfg3_100m"):
        rows = int(100_000_000 * args.scale)
        nq, k = 1024, 100
        idx, local = build_sharded(rows, world, rank, local_rank, dev, 4 << 20)
        Q = unit_queries(nq, dev, 5)
        ms, (s, i) = timed(lambda: idx.search_device(Q, k), 5, world, dev)
        ms_local, _ = timed(lambda: local.search_device(Q, k), 5, world, dev)
        flops_gpu = 2.0 * nq * local.ntotal * DIM
        emit({"config": "BASELINE configs[3]: 100M x 384 bf16 row-sharded, batch 1024, k=100", "scale": args.scale,
              "n_gpus": world, "rows_total": rows, "rows_per_gpu": local.ntotal, "nq": nq, "k": k,
              "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3, "ms_local_search_only": ms_local,
              "exchange": getattr(idx, "exchange", None), "path": local.stats()["path"],
              "tflops_per_gpu": flops_gpu / (ms * 1e-3) / 1e12, "frac_of_bf16_burst_peak": flops_gpu / (ms * 1e-3) / 1e12 / tf_burst,
              "frac_of_bf16_sustained_peak": flops_gpu / (ms * 1e-3) / 1e12 / tf_sus,
              "hbm_roof_ms": local.ntotal * DIM * 2 / (hbm * 1e9) * 1e3, "hbm_gb_per_gpu": local.ntotal * DIM * 2 / 1e9})
        local.close()
        del idx, local
        torch.cuda.empty_cache()

    # ---------------- configs[2]: ANCE sweep 500k queries x 8.8M, top-200 ----------------
    if args.only in ("", "cfg2_ance"):
        rows = int(8_841_823 * args.scale)
        nq_total, top_k, batch, margin = int(500_000 * args.scale), 200, 4096, 0.1
        for mode in ("corpus_sharded", "query_sharded"):
            if mode == "corpus_sharded":
                idx, local = build_sharded(rows, world, rank, local_rank, dev, 16 << 20)
                my_q = range(0, nq_total, batch)                 # every rank sees every batch
            else:
                local = pkg.FlatIPIndex(DIM, metric="inner_product", device=local_rank)
                local.reserve(rows)
                for blk in make_rows(torch, 0, rows, dev):
                    local.add(blk)
                idx = local
                qlo, qhi = shard_range(nq_total, world, rank)     # each rank mines its own queries
                my_q = range(qlo, qhi, batch)
            torch.cuda.synchronize()
            L = pkg._lib.lib()
            import ctypes
            first_ids = None
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # warm the workspaces with one batch
            Qw = unit_queries(batch, dev, 7)
            idx.search_device(Qw, top_k + 1)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            n_done = 0

            def process(b0):
                nonlocal first_ids, n_done
                hi_q = min(b0 + batch, nq_total if mode == "corpus_sharded" else shard_range(nq_total, world, rank)[1])
                nb = hi_q - b0
                Q = unit_queries(nb, dev, 1000 + b0)              # synthetic student embeddings of this batch
                pos = ((torch.arange(b0, b0 + nb, device=dev, dtype=torch.int64) * 7919) % rows).view(nb, 1)
                s, i = idx.search_device(Q, top_k + 1)            # top-(k + positives) of the whole corpus
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                ps = torch.empty((nb, 1), dtype=torch.float32, device=dev)
                if mode == "query_sharded":
                    rc = L.b2s_score_rows_device(local._h, ctypes.c_void_p(Q.data_ptr()), 0, 1, nb,
                                                 ctypes.c_void_p(pos.data_ptr()), 1, ctypes.c_void_p(ps.data_ptr()), stream)
                    assert rc == 0
                else:
                    # the positive's row lives on one shard: score locally (-FLT_MAX elsewhere), max over ranks
                    rc = L.b2s_score_rows_device(local._h, ctypes.c_void_p(Q.data_ptr()), 0, 1, nb,
                                                 ctypes.c_void_p(pos.data_ptr()), 1, ctypes.c_void_p(ps.data_ptr()), stream)
                    assert rc == 0
                    if world > 1:
                        dist.all_reduce(ps, op=dist.ReduceOp.MAX)
                out_i = torch.empty((nb, top_k), dtype=torch.int64, device=dev)
                out_s = torch.empty((nb, top_k), dtype=torch.float32, device=dev)
                cnt = torch.empty((nb,), dtype=torch.int32, device=dev)
                rc = L.b2s_ance_filter_device(local_rank, ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()), nb,
                                              top_k + 1, ctypes.c_void_p(pos.data_ptr()), ctypes.c_void_p(ps.data_ptr()), 1,
                                              ctypes.c_float(margin), top_k, ctypes.c_void_p(out_i.data_ptr()),
                                              ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(cnt.data_ptr()), stream)
                assert rc == 0
                if first_ids is None and b0 == 0:
                    first_ids = out_i[:64].clone()
                n_done += nb

            process(my_q[0])          # warm every allocation of the per-batch pipeline
            torch.cuda.synchronize()
            first_ids, n_done = None, 0
            if world > 1:
                dist.barrier()
            e0.record()
            for b0 in my_q:
                process(b0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            flops_total = 2.0 * nq_total * rows * DIM
            res = {"config": "BASELINE configs[2]: ANCE sweep 500k queries x 8.8M corpus, top-200", "mode": mode,
                   "scale": args.scale, "n_gpus": world, "rows": rows, "queries": nq_total, "top_k": top_k, "margin": margin,
                   "batch": batch, "seconds": ms * 1e-3, "queries_per_s": nq_total / (ms * 1e-3),
                   "tflops_per_gpu": flops_total / world / (ms * 1e-3) / 1e12,
                   "frac_of_bf16_burst_peak": flops_total / world / (ms * 1e-3) / 1e12 / tf_burst,
                   "frac_of_bf16_sustained_peak": flops_total / world / (ms * 1e-3) / 1e12 / tf_sus,
                   "includes": "query generation, exact top-201 search, positive scoring, ANCE margin filter",
                   "exchange": getattr(idx, "exchange", None)}
            if mode == "corpus_sharded":
                keep_first = first_ids
            else:
                # rank 0 owns queries [0, ...): compare its first 64 with the corpus-sharded answer
                if rank == 0 and first_ids is not None and keep_first is not None:
                    res["modes_agree_first64"] = bool(torch.equal(first_ids, keep_first))
            emit(res)
            local.close()
            del idx, local
            torch.cuda.empty_cache()

    if rank == 0 and args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
