"""tools/bench_extras.py -- the parts of bench.py that are not the headline loop: the parity block
(our answers against a plain PyTorch fp32 reference computed block by block on the same rows), the
kernel's own phase stamps, BASELINE configs[4] (batch x k sweep, one GPU) and bounded versions of
configs[2] (ANCE sweep) and configs[3] (100M rows over the GPUs).  Imported lazily by bench.py; nothing
here touches oracle/ or /root/reference."""
from __future__ import annotations

import ctypes
import json
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
TIE_TOL = 1e-3   # north_star: swaps allowed only between ties whose scores agree within 1e-3 (bf16 storage)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
            "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
            "source": "MEASURED_PEAKS.json" if d else "fallback (B200_PROFILING.md)"}


def unit_queries(torch, n, dim, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q = torch.randn((n, dim), generator=g, device=dev, dtype=torch.float32)
    return (q / q.norm(dim=1, keepdim=True)).contiguous()


# ------------------------------------------------------------------------------------------------
# parity: torch fp32 reference, block by block over the same generator the index was built from
# ------------------------------------------------------------------------------------------------
def torch_fp32_topk_rows(torch, make_rows, lo, hi, q, k, dev):
    """Exact fp32 top-k of q [nq, dim] over global rows [lo, hi): (scores [nq, k'], global ids [nq, k'])."""
    torch.backends.cuda.matmul.allow_tf32 = False
    nq = q.shape[0]
    best_s = torch.empty((nq, 0), device=dev)
    best_i = torch.empty((nq, 0), dtype=torch.int64, device=dev)
    base = lo
    for blk in make_rows(torch, lo, hi, dev):
        s = q @ blk.T
        ts, ti = torch.topk(s, min(k, blk.shape[0]), dim=1)
        best_s = torch.cat([best_s, ts], dim=1)
        best_i = torch.cat([best_i, ti + base], dim=1)
        ts, sel = torch.topk(best_s, min(k, best_s.shape[1]), dim=1)
        best_s, best_i = ts, torch.gather(best_i, 1, sel)
        base += blk.shape[0]
    return best_s, best_i


def check_rule(S, I, ref_s, ref_i, k, n_total):
    """The north_star parity rule for one call: S/I [nq, k] ours, ref_* [nq, k + pad] fp32 reference (descending).
    Returns (violations, tie_swaps, exact_order_rows, max |score - fp32 score| over the common ids)."""
    bad = swaps = exact = 0
    max_err = 0.0
    for r in range(I.shape[0]):
        got, exp = I[r].tolist(), ref_i[r][:k].tolist()
        kth = float(ref_s[r][k - 1])
        where = {int(x): j for j, x in enumerate(ref_i[r].tolist())}
        if got == exp:
            exact += 1
        if len(set(got)) != k or min(got) < 0 or max(got) >= n_total or np.any(np.diff(S[r]) > 0):
            bad += 1
            continue
        gs, es = set(got), set(exp)
        for x in gs - es:      # an id the fp32 reference did not return must be a near-tie of its k-th
            j = where.get(x)
            if j is None or abs(float(ref_s[r][j]) - kth) > TIE_TOL:
                bad += 1
            else:
                swaps += 1
        for x in es - gs:
            if abs(float(ref_s[r][where[x]]) - kth) > TIE_TOL:
                bad += 1
        for j, x in enumerate(got):
            if x in where:
                e = abs(float(S[r][j]) - float(ref_s[r][where[x]]))
                max_err = max(max_err, e)
                if e > TIE_TOL:
                    bad += 1
    return bad, swaps, exact, max_err


def verify_parity(torch, dist, idx, make_rows, n_total, dim, world, rank, dev, shard_lo, shard_hi):
    """Untimed parity block of the bench line.  Calls: 16 batch-1 searches at k=10 (scan kernel; sharded: the
    fused merge + NVLink exchange), one 16-query call at k=100 (tensor kernel), one call of 160 queries
    (> #SMs: the two-kernel push / wait exchange when sharded) -- every answer is compared on rank 0 with a
    plain PyTorch fp32 `q @ X.T` + topk over the same rows, computed shard-locally and merged, under the
    1e-3 near-tie rule; every rank must hold the same ids."""
    pad = 16
    Q = unit_queries(torch, 160, dim, dev, 424242)
    calls = [("batch-1 x16, k=10", [Q[i:i + 1] for i in range(16)], 10),
             ("batch-16, k=100", [Q[:16]], 100),
             ("batch-160, k=10", [Q], 10)]
    kmax = 100 + pad
    ls, li = torch_fp32_topk_rows(torch, make_rows, shard_lo, shard_hi, Q, kmax, dev)
    if ls.shape[1] < kmax:   # tiny shards (smoke sizes): pad so that every rank contributes the same shape
        fill = kmax - ls.shape[1]
        ls = torch.cat([ls, torch.full((Q.shape[0], fill), -3.0e38, device=dev)], dim=1)
        li = torch.cat([li, torch.full((Q.shape[0], fill), -1, dtype=torch.int64, device=dev)], dim=1)
    if world > 1:
        gs = [torch.empty_like(ls) for _ in range(world)]
        gi = [torch.empty_like(li) for _ in range(world)]
        dist.all_gather(gs, ls.contiguous())
        dist.all_gather(gi, li.contiguous())
        cs, ci = torch.cat(gs, dim=1), torch.cat(gi, dim=1)
        ts, sel = torch.topk(cs, kmax, dim=1)
        ref_s, ref_i = ts, torch.gather(ci, 1, sel)
    else:
        ref_s, ref_i = ls, li
    ref_s, ref_i = ref_s.cpu().numpy(), ref_i.cpu().numpy()
    report = {"reference": "torch fp32 q @ X.T + topk, 1 Mi-row blocks, shard-local then merged (TF32 off)",
              "rule": f"ids equal the fp32 reference's; differences only among scores within {TIE_TOL} of its k-th; "
                      f"|score - fp32 score| <= {TIE_TOL}; descending; ids valid and unique", "calls": []}
    ok = True
    for name, qs, k in calls:
        outs_s, outs_i = [], []
        for q in qs:
            s, i = idx.search_device(q, k)
            outs_s.append(s.clone())
            outs_i.append(i.clone())
        torch.cuda.synchronize()
        S, I = torch.cat(outs_s), torch.cat(outs_i)
        ranks_equal = True
        if world > 1:
            all_i = [torch.empty_like(I) for _ in range(world)]
            dist.all_gather(all_i, I.contiguous())
            ranks_equal = all(bool(torch.equal(all_i[0], t)) for t in all_i[1:])
        nq = S.shape[0]
        bad, swaps, exact, err = check_rule(S.cpu().numpy(), I.cpu().numpy(), ref_s[:nq], ref_i[:nq], k, n_total)
        st = (idx.local if hasattr(idx, "local") else idx).stats()
        c = {"call": name, "queries": nq, "k": k, "violations": bad, "tie_swaps": swaps, "exact_order_rows": exact,
             "max_abs_score_err": err, "all_ranks_same_ids": ranks_equal, "path": st.get("path")}
        ok = ok and bad == 0 and ranks_equal
        report["calls"].append(c)
    report["ok"] = bool(ok)
    return report


# ------------------------------------------------------------------------------------------------
# phase stamps of the scan kernel (option "trace")
# ------------------------------------------------------------------------------------------------
def tail_breakdown(torch, dist, idx, local, Qd, k, world, rank, n=24):
    """Medians (this rank) of the batch-1 scan kernel's %globaltimer stamps over n traced searches, in
    microseconds from the earliest CTA start of the traced launch; gathered over the ranks."""
    local.set_option("trace", 1)
    rows = []
    for i in range(n):
        idx.search_device(Qd[i:i + 1], k)
        idx.search_device(Qd[i + 1:i + 2], k)   # the traced launch starts behind a predecessor, as in the loop
        t = local.read_trace()
        if t:
            rows.append(t)
    local.set_option("trace", 0)
    mine = None
    if rows:
        keys = [key for key in rows[0] if all(key in r for r in rows)]
        mine = {key: round(float(np.median([r[key] for r in rows])), 2) for key in keys}
        m = mine
        m["after_last_scan_us"] = round(m["done_us"] - m["scan_end_last_us"], 2)
        m["ragged_end_us"] = round(m["scan_end_last_us"] - m["scan_end_first_us"], 2)
        if "pushed_us" in m:
            m["split_us"] = {"last scan end -> ticket": round(m["ticket_us"] - m["scan_end_last_us"], 2),
                             "ticket -> local top-k": round(m["local_topk_us"] - m["ticket_us"], 2),
                             "local top-k -> pushed to peers": round(m["pushed_us"] - m["local_topk_us"], 2),
                             "pushed -> every peer's candidates seen (NVLink + rank skew)": round(m["peers_seen_us"] - m["pushed_us"], 2),
                             "seen -> outputs written": round(m["done_us"] - m["peers_seen_us"], 2)}
        else:
            m["split_us"] = {"last scan end -> ticket": round(m["ticket_us"] - m["scan_end_last_us"], 2),
                             "ticket -> local top-k": round(m["local_topk_us"] - m["ticket_us"], 2),
                             "local top-k -> outputs written": round(m["done_us"] - m["local_topk_us"], 2)}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        return {"per_rank": allr, "note": "medians of %d traced launches per rank; microseconds from the launch's earliest CTA start" % n}
    return {"per_rank": [mine], "note": "medians of %d traced launches; microseconds from the launch's earliest CTA start" % n}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: batch x k sweep on one GPU
# ------------------------------------------------------------------------------------------------
def sweep_report(torch, index, dev, dim, nqs=(1, 16, 128, 256, 1024, 4096), ks=(10, 100, 1000), reps=3, budget_s=20.0):
    pk = measured_peaks()
    bw, tf = pk["hbm_gbs"] * 1e9, pk["bf16_tflops"] * 1e12
    rows_n = index.ntotal
    Q = unit_queries(torch, max(nqs), dim, dev, 7)
    t_start = time.perf_counter()
    points = []
    for k in ks:
        for nq in nqs:
            if time.perf_counter() - t_start > budget_s:
                points.append({"k": k, "nq": nq, "skipped": "time budget"})
                continue
            q = Q[:nq]
            out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
            for _ in range(2):
                index.search_device(q, k, out=out)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                index.search_device(q, k, out=out)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            st = index.stats()
            t_h = rows_n * dim * 2 / bw
            t_t = 2.0 * nq * rows_n * dim / tf
            roof = max(t_h, t_t)
            points.append({"k": k, "nq": nq, "ms": round(best, 4), "queries_per_s": round(nq / best * 1e3, 1),
                           "path": "scan" if st["path"] == 1 else "tensor", "launches": st["kernel_launches"],
                           "bound": "hbm" if t_h >= t_t else "tensor", "roof_ms": round(roof * 1e3, 4),
                           "frac": round(roof * 1e3 / best, 4)})
    cross = t_h * tf / (2.0 * rows_n * dim)
    return {"workload": "BASELINE configs[4]: query-batch sweep at k in {10,100,1000} over the same corpus (bounded: "
                        f"{len(nqs)} batch sizes, best of {reps} whole calls, CUDA events)",
            "roofline": "t_roof = max(rows*dim*2 / measured HBM copy peak, 2*nq*rows*dim / measured dense bf16 burst peak)",
            "peaks": pk, "roofline_crossover_nq": round(cross, 1),
            "routing": "nq <= 2 -> scan kernel (K1); nq >= 3 -> TMA + tcgen05 kernel (K2)", "points": points}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[3] (100M rows over the GPUs, batch 1024, k=100) and configs[2] (ANCE sweep), bounded
# ------------------------------------------------------------------------------------------------
def _timed_max(torch, dist, fn, reps, world, dev):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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


def cfg_100m_report(torch, dist, pkg, make_rows, dim, world, rank, local_rank, dev, rows=100_000_000, nq=1024, k=100):
    from semantic_search_kd_b200.sharded import ShardedFlatIPIndex, shard_range
    pk = measured_peaks()
    lo, hi = shard_range(rows, world, rank)
    local = pkg.FlatIPIndex(dim, metric="inner_product", device=local_rank)
    local.reserve(hi - lo)
    t0 = time.perf_counter()
    for blk in make_rows(torch, lo, hi, dev):
        local.add(blk)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    if world > 1:
        idx = ShardedFlatIPIndex(dim, metric="inner_product", local_index=local, exchange_slot_bytes=4 << 20)
        idx.local.set_id_offset(lo)
        idx.n_total, idx.range = rows, (lo, hi)
    else:
        idx = local
    Q = unit_queries(torch, nq, dim, dev, 5)
    ms, (s, i) = _timed_max(torch, dist, lambda: idx.search_device(Q, k), 5, world, dev)
    # parity of this very call against the torch fp32 reference (16 of the queries)
    ls, li = torch_fp32_topk_rows(torch, make_rows, lo, hi, Q[:16], k + 16, dev)
    if world > 1:
        gs = [torch.empty_like(ls) for _ in range(world)]
        gi = [torch.empty_like(li) for _ in range(world)]
        dist.all_gather(gs, ls.contiguous())
        dist.all_gather(gi, li.contiguous())
        cs, ci = torch.cat(gs, dim=1), torch.cat(gi, dim=1)
        ts, sel = torch.topk(cs, k + 16, dim=1)
        ls, li = ts, torch.gather(ci, 1, sel)
    bad, swaps, exact, err = check_rule(s[:16].cpu().numpy(), i[:16].cpu().numpy(), ls.cpu().numpy(), li.cpu().numpy(), k, rows)
    flops_gpu = 2.0 * nq * local.ntotal * dim
    tfs = flops_gpu / (ms * 1e-3) / 1e12
    res = {"workload": f"BASELINE configs[3]: {rows:,} x {dim} bf16 rows row-sharded over {world} GPU(s), batch {nq}, k={k}, "
                       "candidates exchanged over NVLink peer memory inside the merge kernel",
           "rows_per_gpu": local.ntotal, "hbm_gb_per_gpu": round(local.ntotal * dim * 2 / 1e9, 2), "build_s": round(build_s, 2),
           "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3, "exchange": getattr(idx, "exchange", None),
           "tflops_per_gpu": tfs, "frac_of_bf16_sustained_peak": tfs / pk["bf16_tflops_sustained"],
           "frac_of_bf16_burst_peak": tfs / pk["bf16_tflops"],
           "denominator": "sustained dense bf16 peak (a 7 ms call repeated back to back), burst also given",
           "parity_16_queries": {"violations": bad, "tie_swaps": swaps, "exact_order_rows": exact, "max_abs_score_err": err,
                                 "ok": bad == 0}}
    local.close()
    del idx, local
    torch.cuda.empty_cache()
    return res


def cfg_ance_report(torch, dist, pkg, idx, local, dim, rows, world, rank, local_rank, dev,
                    nq_total=32768, top_k=200, batch=4096, margin=0.1, mode="corpus_sharded"):
    """BASELINE configs[2], bounded: nq_total (of the 500k) queries x the 8.8M corpus, exact top-(200 + 1 positive),
    positive scoring, ANCE margin filter (csrc/ance_filter.cuh).
      mode "corpus_sharded": the corpus this run already holds row-sharded; every rank sees every query batch, the
                             candidates are exchanged over NVLink inside the merge kernels;
      mode "query_sharded":  `idx` = `local` holds the WHOLE corpus on every GPU (6.8 GB of 180), every rank mines its
                             own 1/world of the queries -- no exchange at all (the better layout for this workload)."""
    pk = measured_peaks()
    L = pkg._lib.lib()
    from semantic_search_kd_b200.sharded import shard_range
    if mode == "query_sharded":
        qlo, qhi = shard_range(nq_total, world, rank)
    else:
        qlo, qhi = 0, nq_total
    phases = {"queries": 0.0, "search": 0.0, "positives": 0.0, "filter": 0.0}

    def process(b0, ev=None):
        nb = min(batch, qhi - b0)
        if ev:
            ev[0].record()
        Q = unit_queries(torch, nb, dim, dev, 1000 + b0)
        pos = ((torch.arange(b0, b0 + nb, device=dev, dtype=torch.int64) * 7919) % rows).view(nb, 1)
        if ev:
            ev[1].record()
        s, i = idx.search_device(Q, top_k + 1)
        if ev:
            ev[2].record()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ps = torch.empty((nb, 1), dtype=torch.float32, device=dev)
        rc = L.b2s_score_rows_device(local._h, ctypes.c_void_p(Q.data_ptr()), 0, 1, nb, ctypes.c_void_p(pos.data_ptr()), 1,
                                     ctypes.c_void_p(ps.data_ptr()), stream)
        assert rc == 0, pkg._lib.last_error()
        if world > 1 and mode == "corpus_sharded":   # the positive's row lives on one shard (-FLT_MAX elsewhere)
            dist.all_reduce(ps, op=dist.ReduceOp.MAX)
        if ev:
            ev[3].record()
        out_i = torch.empty((nb, top_k), dtype=torch.int64, device=dev)
        out_s = torch.empty((nb, top_k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nb,), dtype=torch.int32, device=dev)
        rc = L.b2s_ance_filter_device(local_rank, ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()), nb, top_k + 1,
                                      ctypes.c_void_p(pos.data_ptr()), ctypes.c_void_p(ps.data_ptr()), 1, ctypes.c_float(margin),
                                      top_k, ctypes.c_void_p(out_i.data_ptr()), ctypes.c_void_p(out_s.data_ptr()),
                                      ctypes.c_void_p(cnt.data_ptr()), stream)
        assert rc == 0, pkg._lib.last_error()
        if ev:
            ev[4].record()
        return cnt

    process(qlo)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b0 in range(qlo, qhi, batch):
        cnt = process(b0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the same loop once more with events between the phases (untimed: where a batch's time goes)
    n_b = 0
    for b0 in range(qlo, qhi, batch):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        process(b0, ev)
        torch.cuda.synchronize()
        for name, a, b in (("queries", 0, 1), ("search", 1, 2), ("positives", 2, 3), ("filter", 3, 4)):
            phases[name] += ev[a].elapsed_time(ev[b])
        n_b += 1
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    flops_gpu = 2.0 * nq_total * rows * dim / world
    tfs = flops_gpu / (ms * 1e-3) / 1e12
    return {"workload": f"BASELINE configs[2] (bounded sample): ANCE sweep, {nq_total} of the 500k queries x {rows:,} rows, "
                        f"{mode.replace('_', '-')} over {world} GPU(s), top-{top_k} negatives, margin {margin}, batches of {batch}",
            "mode": mode, "seconds": ms * 1e-3, "queries_per_s": nq_total / (ms * 1e-3),
            "full_500k_sweep_s_at_this_rate": 500_000 / (nq_total / (ms * 1e-3)),
            "tflops_per_gpu": tfs, "frac_of_bf16_sustained_peak": tfs / pk["bf16_tflops_sustained"],
            "frac_of_bf16_burst_peak": tfs / pk["bf16_tflops"],
            "includes": "query generation, exact top-201 search" + (" + candidate exchange" if mode == "corpus_sharded" and world > 1 else "")
                        + ", positive scoring, ANCE margin filter",
            "ms_per_batch_by_phase_rank0": {k: round(v / max(1, n_b), 4) for k, v in phases.items()},
            "negatives_last_batch_mean": float(cnt.float().mean().item()), "exchange": getattr(idx, "exchange", None)}
