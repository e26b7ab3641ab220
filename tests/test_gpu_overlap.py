"""GPU: consecutive searches of one handle that are allowed to overlap on the device.

  * B2S_SEARCH_STABLE_QUERIES: the scan of call i+1 runs while the tail of call i (top-k read-out, on a sharded
    index the NVLink exchange) is still in flight -- the two control sets of the scan kernel, their epoch
    words and the one remaining grid dependency (scan_topk.cuh) must give bit-identical answers to the same
    calls issued one at a time, for hundreds of calls in a row, with k switching between the cascade select
    (k <= 16), the per-CTA lists (k > 16) and the tensor kernel in between.
  * searches of one handle issued on DIFFERENT streams never overlap (the handle's workspaces are shared): the
    event guard of b2s_search_device orders them.
  * a search captured into a CUDA graph takes no part in the epoch protocol and may be replayed between live
    overlapped calls.
The shard is large enough (700k rows >= cascade_min_units iterations per warp) for the cascade select.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, D = 700_000, 384


@pytest.fixture(scope="module")
def big():
    import torch
    import semantic_search_kd_b200 as pkg
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    X = torch.randn((N, D), generator=g, device=dev)
    X = X / X.norm(dim=1, keepdim=True)
    idx = pkg.FlatIPIndex(D, metric="inner_product", device=0)
    idx.add(X)
    Q = torch.randn((512, D), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).contiguous()
    torch.cuda.synchronize()
    yield pkg, idx, X, Q, dev
    idx.close()


def _serial(idx, Q, ks):
    import torch
    outs = []
    for i, k in enumerate(ks):
        s, ids = idx.search_device(Q[i:i + 1], k)
        torch.cuda.synchronize()
        outs.append((s.clone(), ids.clone()))
    return outs


def test_cascade_select_matches_torch_reference(big):
    """The cascade select (k <= 16 on a large shard) against torch fp32 on the bf16 rows the index holds."""
    import torch
    pkg, idx, X, Q, dev = big
    Xb = X.to(torch.bfloat16).float()
    for k in (1, 10, 16):
        s, ids = idx.search_device(Q[:1], k)
        torch.cuda.synchronize()
        ref_s, ref_i = torch.topk(Q[:1] @ Xb.T, k, dim=1)
        assert torch.equal(ids, ref_i), (k, ids, ref_i)
        assert torch.allclose(s, ref_s, atol=2e-5)
    s2, ids2 = idx.search_device(Q[:2], 10)           # two queries in one launch
    torch.cuda.synchronize()
    ref_s, ref_i = torch.topk(Q[:2] @ Xb.T, 10, dim=1)
    assert torch.equal(ids2, ref_i) and torch.allclose(s2, ref_s, atol=2e-5)


@pytest.mark.parametrize("pattern", ["k10", "mixed"])
def test_stable_queries_overlap_is_bit_identical(big, pattern):
    import torch
    pkg, idx, X, Q, dev = big
    n = 300
    ks = [10] * n if pattern == "k10" else [(10, 100, 3, 16, 17, 1)[i % 6] for i in range(n)]
    ref = _serial(idx, Q, ks)
    outs = []
    for i, k in enumerate(ks):                       # back to back, no synchronisation in between
        outs.append(idx.search_device(Q[i:i + 1], k, stable_queries=True))
        if pattern == "mixed" and i % 50 == 25:      # a tensor-path call of the same handle in the chain
            idx.search_device(Q[:8], 10)
    torch.cuda.synchronize()
    for i, ((s, ids), (rs, ri)) in enumerate(zip(outs, ref)):
        assert torch.equal(ids, ri), (i, ks[i])
        assert torch.equal(s, rs), (i, ks[i])


def test_stable_queries_option_pdl2_and_host_calls_interleaved(big):
    """Option pdl=2 (every device call treated as stable) with host-buffer calls in between."""
    import torch
    pkg, idx, X, Q, dev = big
    Qh = Q.cpu().numpy()
    ref = _serial(idx, Q, [10] * 40)
    idx.set_option("pdl", 2)
    try:
        outs = []
        for i in range(40):
            outs.append(idx.search_device(Q[i:i + 1], 10))
            if i % 7 == 3:
                sh, ih = idx.search(Qh[i:i + 1], 10)   # private stream + stream guard
                assert np.array_equal(ih, ref[i][1].cpu().numpy())
        torch.cuda.synchronize()
        for (s, ids), (rs, ri) in zip(outs, ref):
            assert torch.equal(ids, ri) and torch.equal(s, rs)
    finally:
        idx.set_option("pdl", 1)


def test_searches_on_different_streams_do_not_overlap(big):
    """Two torch streams alternate on one handle, one of them behind a long-running kernel: without the event
    guard the second stream's search would run concurrently with the first one's and share its workspaces."""
    import torch
    pkg, idx, X, Q, dev = big
    ref = _serial(idx, Q, [10] * 64)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    outs = []
    for i in range(64):
        st = s1 if i % 2 == 0 else s2
        with torch.cuda.stream(st):
            outs.append(idx.search_device(Q[i:i + 1], 10, stable_queries=(i % 3 == 0)))
    torch.cuda.synchronize()
    for i, ((s, ids), (rs, ri)) in enumerate(zip(outs, ref)):
        assert torch.equal(ids, ri) and torch.equal(s, rs), i


def test_graph_replay_between_overlapped_calls(big):
    import torch
    pkg, idx, X, Q, dev = big
    ref = _serial(idx, Q, [10] * 30)
    q = Q[100:101].clone()
    out = (torch.empty((1, 10), dtype=torch.float32, device=dev), torch.empty((1, 10), dtype=torch.int64, device=dev))
    idx.search_device(q, 10, out=out)
    torch.cuda.synchronize()
    want = (out[0].clone(), out[1].clone())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            idx.search_device(q, 10, out=out)
    outs = []
    for i in range(30):
        outs.append(idx.search_device(Q[i:i + 1], 10, stable_queries=True))
        if i % 10 == 5:
            torch.cuda.synchronize()                 # the graph is ordered by whoever replays it
            out[0].zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(out[1], want[1]) and torch.equal(out[0], want[0])
    torch.cuda.synchronize()
    for (s, ids), (rs, ri) in zip(outs, ref):
        assert torch.equal(ids, ri) and torch.equal(s, rs)


def test_trace_reports_both_launches(big):
    import torch
    pkg, idx, X, Q, dev = big
    idx.set_option("trace", 1)
    try:
        for i in range(6):
            idx.search_device(Q[i:i + 1], 10, stable_queries=True)
        t = idx.read_trace(previous=True)
    finally:
        idx.set_option("trace", 0)
    assert t is not None and t["grid"] > 0 and t["done_us"] > t["scan_end_first_us"] > 0
    assert "previous" in t and 0 < t["period_us"] < 10_000


def test_every_kernel_family_small_shapes():
    """tools/sanitize_smoke.py (written for compute-sanitizer, which this pool keeps closed) as a plain run: every
    kernel family on tiny shapes, including the cascade select forced onto a 140k-row shard (cascade_min_units = 4),
    both transition variants, overlapped launches with the trace on."""
    import importlib.util
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("sanitize_smoke", root / "tools" / "sanitize_smoke.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_host_completion_flag_and_stream_wait_agree(big):
    """Host-buffer calls of 1-2 queries: waiting on the kernel's completion word (default) and waiting on the stream
    (option host_spin = 0) return the same bytes, call after call, also right after device-path calls."""
    import torch
    pkg, idx, X, Q, dev = big
    Qh = Q.cpu().numpy()
    want = []
    idx.set_option("host_spin", 0)
    try:
        for i in range(24):
            want.append(idx.search(Qh[i:i + 1 + (i % 2)], (10, 3, 16, 50)[i % 4]))
    finally:
        idx.set_option("host_spin", 1)
    for i in range(24):
        if i % 5 == 0:
            idx.search_device(Q[i:i + 1], 10, stable_queries=True)       # something in flight on another stream
        D, I = idx.search(Qh[i:i + 1 + (i % 2)], (10, 3, 16, 50)[i % 4])
        assert np.array_equal(I, want[i][1]) and np.array_equal(D, want[i][0]), i


@pytest.mark.parametrize("case", ["duplicates", "ascending", "descending", "all_equal", "two_values", "ascending_large"])
def test_cascade_select_on_adversarial_data(case):
    """The cascade select forced onto a 150k-row shard (cascade_min_units = 4) with the inputs that stress it: exact
    ties (the lower id must win, as in faiss' strict-'>' heap and the oracle), scores that rise with the row index
    (every row beats the running k-th: an insertion storm into the global slots), scores that fall, all rows equal,
    and two distinct score values.  Reference: torch fp32 over the bf16 rows the index holds, ties by lower id."""
    import torch
    import semantic_search_kd_b200 as pkg
    dev = torch.device("cuda", 0)
    n, d = (700_000 if case == "ascending_large" else 150_000), 384   # large: dozens of phase B iterations -> the storm guard
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    X = torch.randn((n, d), generator=g, device=dev)
    X = X / X.norm(dim=1, keepdim=True)
    q = torch.randn((2, d), generator=g, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    if case == "duplicates":
        X[n // 2:] = X[: n - n // 2]
    elif case in ("ascending", "descending", "ascending_large"):
        order = torch.argsort((X.to(torch.bfloat16).float() @ q[0]), descending=(case == "descending"))
        X = X[order].contiguous()
    elif case == "all_equal":
        X[:] = X[0]
    elif case == "two_values":
        X[:] = X[0]
        X[1::3] = X[1]
    idx = pkg.FlatIPIndex(d, metric="inner_product", device=0)
    if case != "ascending_large":
        idx.set_option("cascade_min_units", 4)
    idx.add(X)
    Xb = X.to(torch.bfloat16).float()
    for k in (1, 10, 16):
        for qq in (q[:1], q):
            s, ids = idx.search_device(qq, k)
            torch.cuda.synchronize()
            sc = qq @ Xb.T                                                  # fp32 reference scores
            for r in range(qq.shape[0]):
                got_i, got_s = ids[r].cpu().numpy(), s[r].cpu().numpy()
                assert len(set(got_i.tolist())) == k and got_i.min() >= 0 and got_i.max() < n
                assert np.all(np.diff(got_s) <= 0)
                # our scores: fp32 accumulation of the same products in another order -> compare through a tolerance,
                # ids through the rule "every returned row scores at least the (k+1)-th best minus the tolerance"
                ref_sorted, _ = torch.sort(sc[r], descending=True)
                kth = float(ref_sorted[k - 1])
                assert np.all(sc[r][torch.from_numpy(got_i).to(dev)].cpu().numpy() >= kth - 2e-5), (case, k, r)
                assert np.allclose(got_s, ref_sorted[:k].cpu().numpy(), atol=2e-5), (case, k, r)
                if case in ("all_equal", "duplicates", "two_values"):
                    # exact ties (identical rows -> identical score bits): the lower id wins, in order
                    order = torch.argsort(-sc[r], stable=True)[:k]               # stable: equal scores keep id order
                    assert got_i.tolist() == order.cpu().tolist(), (case, k, r)
    st = idx.stats()
    assert st["path"] == 1
    idx.close()


def test_completion_word_is_not_trusted_across_handles(oracle):
    """Regression (found by test_ragged_sizes_scan during round 2): the completion word of a host call lives in
    pinned memory behind the outputs, so it moves with (nq, k), and pinned memory is recycled between handles.  A
    new handle must never see the value an earlier handle left there and return before its kernel has written."""
    import semantic_search_kd_b200 as pkg
    rng = np.random.default_rng(11)
    for gen in range(6):
        X = rng.standard_normal((37 + gen, 384)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        q = X[gen:gen + 1] + 0.01 * rng.standard_normal((1, 384)).astype(np.float32)
        idx = pkg.FlatIPIndex(384, metric="inner_product")
        idx.set_option("path", 1)
        idx.add(X)
        for k in (1, 10, 1, 10):                      # the same call sequence on every handle: same flag values
            D, I = idx.search(q, k)
            Dr, Ir = oracle.flat_ip_topk(X, q, k)
            rep = oracle.compare_topk(D, I, Dr, Ir, X, q, tie_tol=1e-3)
            assert I[0, 0] == gen and rep["ok"], (gen, k, I, rep)
        idx.close()
