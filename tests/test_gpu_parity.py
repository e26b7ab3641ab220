"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): ids identical to the fp32 flat oracle, swaps allowed only between
ties whose fp32 scores agree within 1e-3 under bf16 storage.  Tighter checks where available:
against the oracle run on the SAME bf16-rounded corpus the device holds, the scan path (fp32
query, fp32 accumulate) must agree to fp32 summation-order noise (1e-5).
"""
import numpy as np
import pytest

from conftest import unit_rows

pytestmark = pytest.mark.gpu

TIE_TOL_BF16 = 1e-3     # north_star tolerance for bf16 storage
TIE_TOL_F32 = 2e-5      # fp32 summation-order noise


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import semantic_search_kd_b200 as m
    return m


def build(pkg, X, metric="inner_product", **opts):
    idx = pkg.FlatIPIndex(X.shape[1], metric=metric)
    for k, v in opts.items():
        idx.set_option(k, v)
    idx.add(X)
    return idx


def check(oracle, idx, X, Q, k, tight=True, q_bf16=False):
    """q_bf16: the tensor path rounds the queries to bf16 as well (tcgen05 operands are bf16)."""
    D, I = idx.search(Q, k)
    q_bf16 = q_bf16 or (Q.shape[0] > 0 and k > 0 and idx.ntotal > 0 and idx.stats()["path"] == 2)
    assert D.shape == (Q.shape[0], k) and I.shape == (Q.shape[0], k)
    assert D.dtype == np.float32 and I.dtype == np.int64
    Dr, Ir = oracle.flat_ip_topk(X, Q, k)
    rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL_BF16)
    assert rep["ok"], rep
    assert rep["max_abs_score_err"] <= TIE_TOL_BF16, rep
    valid = I >= 0
    assert np.all(np.diff(np.where(valid, D, -np.inf), axis=1)[valid[:, 1:]] <= 0), "scores not descending"
    if tight:
        Xb = oracle.round_bf16(X)
        Qt = oracle.round_bf16(Q) if q_bf16 else Q
        Db, Ib = oracle.flat_ip_topk(Xb, Qt, k)
        rep2 = oracle.compare_topk(D, I, Db, Ib, Xb, Qt, tie_tol=TIE_TOL_F32)
        assert rep2["ok"], rep2
        assert rep2["max_abs_score_err"] <= TIE_TOL_F32, rep2
    return D, I


def test_golden_vectors_scan_path(pkg, oracle, golden_cases):
    for name, (X, Q, z) in golden_cases.items():
        idx = build(pkg, X, path=1)
        for k in sorted(int(f[5:]) for f in z.files if f.startswith("ids_k")):
            D, I = idx.search(Q, k)
            ref_I, ref_D = z[f"ids_k{k}"], z[f"scores_k{k}"]
            rep = oracle.compare_topk(D, I, ref_D, ref_I, X, Q, tie_tol=TIE_TOL_BF16)
            assert rep["ok"], (name, k, rep)
            if f"ids_bf16corpus_k{k}" in z.files:   # same rounding as the device: must be identical
                rep = oracle.compare_topk(D, I, z[f"scores_bf16corpus_k{k}"], z[f"ids_bf16corpus_k{k}"],
                                          oracle.round_bf16(X), Q, tie_tol=TIE_TOL_F32)
                assert rep["ok"], (name, k, rep)
                if name == "rand2000":
                    assert np.array_equal(I, z[f"ids_bf16corpus_k{k}"]), (name, k)
        idx.close()


def test_conftest_fixture_identity(pkg, golden_cases):
    """The reference fixture (tests/conftest.py:65-73): query = corpus row -> top-1 is itself, score 1."""
    X, Q, z = golden_cases["conftest_fixture"]
    idx = build(pkg, X)
    D, I = idx.search(X, 1)
    assert list(I[:, 0]) == list(range(10))
    np.testing.assert_allclose(D[:, 0], 1.0, atol=4e-3)
    D, I = idx.search(Q, 12)                       # k > ntotal
    assert np.all(I[:, 10:] == -1) and np.all(D[:, 10:] == np.float32(-3.4028234663852886e38))
    assert np.array_equal(np.sort(I[:, :10], axis=1), np.tile(np.arange(10), (Q.shape[0], 1)))
    idx.close()


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 1000, 4097, 50000])
@pytest.mark.parametrize("nq", [1, 2, 3, 4, 7])
def test_ragged_sizes_scan(pkg, oracle, n, nq):
    X, Q = unit_rows(n, 384, n), unit_rows(nq, 384, 1000 + nq)
    idx = build(pkg, X, path=1)
    for k in (1, 10):
        check(oracle, idx, X, Q, k)
    idx.close()


@pytest.mark.parametrize("k", [1, 10, 32, 100, 200, 1000, 2048])
def test_k_sweep_scan(pkg, oracle, k):
    X, Q = unit_rows(30000, 384, 5), unit_rows(3, 384, 6)
    for seed in (0, 1):
        idx = build(pkg, X, path=1, seed=seed)
        check(oracle, idx, X, Q, k)
        assert idx.stats()["seeded"] == seed
        idx.close()


def test_exact_ties_prefer_lower_id(pkg, oracle):
    X = unit_rows(5000, 384, 9)
    X[100:140] = X[7]            # 41 identical rows -> identical scores
    X[4990:] = X[7]
    Q = np.concatenate([X[[7]], unit_rows(2, 384, 10)])
    idx = build(pkg, X, path=1)
    D, I = idx.search(Q, 20)
    assert list(I[0]) == [7] + list(range(100, 119)), I[0]
    check(oracle, idx, X, Q, 20)
    idx.close()


def test_adversarial_ascending_scores(pkg, oracle):
    """Every row beats all earlier ones: the candidate lists overflow and compact continuously."""
    rng = np.random.default_rng(3)
    q = unit_rows(1, 384, 77)[0]
    noise = unit_rows(20000, 384, 78)
    t = np.linspace(-0.9, 0.9, 20000, dtype=np.float32)[:, None]
    X = t * q[None, :] + 0.05 * noise
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = X.astype(np.float32)
    Q = np.stack([q, -q]).astype(np.float32)
    idx = build(pkg, X, path=1)
    for k in (10, 100):
        check(oracle, idx, X, Q, k)
    idx.close()


def test_all_equal_scores(pkg):
    X = np.tile(unit_rows(1, 384, 1), (3000, 1))
    idx = build(pkg, X, path=1)
    D, I = idx.search(X[:1], 50)
    assert list(I[0]) == list(range(50))
    idx.close()


def test_cosine_metric_normalises(pkg, oracle):
    rng = np.random.default_rng(4)
    Xr = rng.standard_normal((4000, 384)).astype(np.float32) * rng.uniform(0.1, 9, (4000, 1)).astype(np.float32)
    Qr = rng.standard_normal((5, 384)).astype(np.float32) * 3
    X = Xr / np.linalg.norm(Xr, axis=1, keepdims=True)
    Q = Qr / np.linalg.norm(Qr, axis=1, keepdims=True)
    idx = build(pkg, Xr, metric="cosine", path=1)
    D, I = idx.search(Qr, 10)
    Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
    rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL_BF16)
    assert rep["ok"], rep
    idx.close()


def test_incremental_add_and_id_offset(pkg, oracle):
    X, Q = unit_rows(9000, 384, 31), unit_rows(4, 384, 32)
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    for s in range(0, 9000, 2500):
        idx.add(X[s:s + 2500])
    assert idx.ntotal == 9000
    check(oracle, idx, X, Q, 10)
    idx.set_id_offset(1_000_000_000_000)
    D, I = idx.search(Q, 10)
    Dr, Ir = oracle.flat_ip_topk(oracle.round_bf16(X), Q, 10)
    assert np.array_equal(I - 1_000_000_000_000, Ir)
    idx.close()


def test_torch_device_path_and_bf16_queries(pkg, oracle):
    import torch
    X, Q = unit_rows(20000, 384, 41), unit_rows(6, 384, 42)
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    idx.set_option("path", 1)
    idx.add(torch.from_numpy(X).cuda())
    s, i = idx.search(torch.from_numpy(Q).cuda(), 10)
    assert s.is_cuda and i.is_cuda and s.dtype == torch.float32 and i.dtype == torch.int64
    Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
    rep = oracle.compare_topk(s.cpu().numpy(), i.cpu().numpy(), Dr, Ir, X, Q, tie_tol=TIE_TOL_BF16)
    assert rep["ok"], rep
    s2, i2 = idx.search(torch.from_numpy(Q).cuda().bfloat16(), 10)
    Qb = oracle.round_bf16(Q)
    Db, Ib = oracle.flat_ip_topk(oracle.round_bf16(X), Qb, 10)
    rep = oracle.compare_topk(s2.cpu().numpy(), i2.cpu().numpy(), Db, Ib, oracle.round_bf16(X), Qb,
                              tie_tol=TIE_TOL_F32)
    assert rep["ok"], rep
    idx.close()


def test_empty_index_and_zero_k(pkg):
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    idx.add(np.zeros((0, 384), np.float32))
    D, I = idx.search(unit_rows(2, 384, 1), 5)
    assert np.all(I == -1)
    D, I = idx.search(unit_rows(2, 384, 1), 0)
    assert D.shape == (2, 0)
    D, I = idx.search(np.zeros((0, 384), np.float32), 5)
    assert D.shape == (0, 5)
    idx.close()


def test_other_dims(pkg, oracle):
    for d in (128, 256, 512, 768, 1024):
        X, Q = unit_rows(3000, d, d), unit_rows(3, d, d + 1)
        idx = build(pkg, X, path=1)
        D, I = idx.search(Q, 10)
        Dr, Ir = oracle.flat_ip_topk(oracle.round_bf16(X), Q, 10)
        rep = oracle.compare_topk(D, I, Dr, Ir, oracle.round_bf16(X), Q, tie_tol=TIE_TOL_F32)
        assert rep["ok"], (d, rep)
        idx.close()


def test_save_load_roundtrip(pkg, oracle, tmp_path):
    X, Q = unit_rows(3000, 384, 51), unit_rows(3, 384, 52)
    idx = build(pkg, X)
    idx.doc_ids = [f"doc_{i}" for i in range(3000)]
    D0, I0 = idx.search(Q, 10)
    idx.save(tmp_path / "index")
    assert (tmp_path / "index" / "index.faiss").stat().st_size == 45 + 3000 * 384 * 4
    for drop_sidecar in (False, True):
        if drop_sidecar:
            (tmp_path / "index" / "rows.bf16").unlink()
        j = pkg.FAISSIndexBuilder(embedding_dim=384, metric="inner_product")
        j.load(tmp_path / "index")          # src/serve/app.py:430-433
        assert j.ntotal == 3000 and j.doc_ids[5] == "doc_5"
        D1, I1 = j.search(Q, 10)
        assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
        j.close()
    idx.close()


def test_similarity_matches_oracle(pkg, oracle):
    """compute_similarity (tests/test_student_model.py:104-124): shape (2,3), range, values."""
    import ctypes
    L = pkg._lib.lib()
    q, d = unit_rows(2, 384, 1), unit_rows(3, 384, 2)
    out = np.empty((2, 3), np.float32)
    rc = L.b2s_similarity(0, q.ctypes.data_as(ctypes.c_void_p), 2, d.ctypes.data_as(ctypes.c_void_p), 3, 384,
                          out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, pkg._lib.last_error()
    assert out.shape == (2, 3) and np.all(out >= -1.01) and np.all(out <= 1.01)
    np.testing.assert_allclose(out, oracle.similarity_np(q, d), atol=1e-6)


def test_merge_device_matches_global_topk(pkg, oracle):
    """Row-sharded search emulated on one GPU: G shards searched one after another, candidates
    stacked as the all-gather would, merged by b2s_merge_device."""
    import ctypes
    import torch
    X, Q = unit_rows(40000, 384, 61), unit_rows(5, 384, 62)
    X[30000:30010] = X[5]     # ties across shards
    Q[0] = X[5]
    G, k = 4, 10
    parts = np.array_split(np.arange(40000), G)
    sc, ids = [], []
    for p in parts:
        idx = build(pkg, X[p], path=1)
        idx.set_id_offset(int(p[0]))
        s, i = idx.search(torch.from_numpy(Q).cuda(), k)
        sc.append(s)
        ids.append(i)
        idx.close()
    S = torch.stack(sc).contiguous()
    I = torch.stack(ids).contiguous()
    outS = torch.empty((5, k), dtype=torch.float32, device="cuda")
    outI = torch.empty((5, k), dtype=torch.int64, device="cuda")
    rc = pkg._lib.lib().b2s_merge_device(0, ctypes.c_void_p(S.data_ptr()), ctypes.c_void_p(I.data_ptr()), G, 5, k,
                                         ctypes.c_void_p(outS.data_ptr()), ctypes.c_void_p(outI.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, pkg._lib.last_error()
    torch.cuda.synchronize()
    Xb = oracle.round_bf16(X)
    Dr, Ir = oracle.flat_ip_topk(Xb, Q, k)
    rep = oracle.compare_topk(outS.cpu().numpy(), outI.cpu().numpy(), Dr, Ir, Xb, Q, tie_tol=TIE_TOL_F32)
    assert rep["ok"], rep
    assert list(outI[0].cpu().numpy()[:3]) == [5, 30000, 30001]


# ------------------------------------------------------------------------------------------------
# tensor path (K2: TMA + tcgen05 + TMEM epilogue), forced with path=2
# ------------------------------------------------------------------------------------------------

def test_golden_vectors_tensor_path(pkg, oracle, golden_cases):
    for name, (X, Q, z) in golden_cases.items():
        idx = build(pkg, X, path=2)
        for k in sorted(int(f[5:]) for f in z.files if f.startswith("ids_k")):
            D, I = idx.search(Q, k)
            assert idx.stats()["path"] == 2
            rep = oracle.compare_topk(D, I, z[f"scores_k{k}"], z[f"ids_k{k}"], X, Q, tie_tol=TIE_TOL_BF16)
            assert rep["ok"], (name, k, rep)
        idx.close()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 256, 257, 1000, 20000])
@pytest.mark.parametrize("nq", [1, 5, 31, 32, 33, 64, 100, 129, 256, 257, 300, 600])
def test_ragged_sizes_tensor(pkg, oracle, n, nq):
    X, Q = unit_rows(n, 384, n + 7), unit_rows(nq, 384, 2000 + nq)
    idx = build(pkg, X, path=2)
    check(oracle, idx, X, Q, 10, q_bf16=True)
    idx.close()


@pytest.mark.parametrize("k", [1, 10, 100, 200, 1000, 2048])
def test_k_sweep_tensor(pkg, oracle, k):
    X, Q = unit_rows(40000, 384, 15), unit_rows(70, 384, 16)
    for seed, shared in ((0, 1), (1, 1), (1, 0)):
        idx = build(pkg, X, path=2, seed=seed, tc_shared_thr=shared)
        check(oracle, idx, X, Q, k, q_bf16=True)
        idx.close()


@pytest.mark.parametrize("k", [10, 100, 1000])
def test_tensor_shared_threshold_large(pkg, oracle, k):
    """The shared survivor histogram tightens thresholds across CTA pairs: sparse sample (1/64),
    1M rows, several query blocks; must stay exact, also with near-duplicate-heavy data."""
    X, Q = unit_rows(1000000, 384, 81), unit_rows(300, 384, 82)
    X[5000:5400] = X[4999] + 1e-3 * unit_rows(400, 384, 83)      # a dense cluster of near ties
    X[5000:5400] /= np.linalg.norm(X[5000:5400], axis=1, keepdims=True)
    Q[0] = X[4999]
    idx = build(pkg, X, path=2)
    check(oracle, idx, X, Q, k, q_bf16=True)
    idx.close()


def test_tensor_ties_and_adversarial(pkg, oracle):
    X = unit_rows(6000, 384, 19)
    X[200:260] = X[9]
    X[5990:] = X[9]
    Q = np.concatenate([X[[9]], unit_rows(40, 384, 20)])
    idx = build(pkg, X, path=2)
    D, I = idx.search(Q, 20)
    assert list(I[0]) == [9] + list(range(200, 219)), I[0]
    check(oracle, idx, X, Q, 20, q_bf16=True)
    idx.close()
    # ascending scores: every tile raises every threshold
    q = unit_rows(1, 384, 77)[0]
    t = np.linspace(-0.9, 0.9, 30000, dtype=np.float32)[:, None]
    Xa = t * q[None, :] + 0.05 * unit_rows(30000, 384, 78)
    Xa = (Xa / np.linalg.norm(Xa, axis=1, keepdims=True)).astype(np.float32)
    Qa = np.concatenate([np.stack([q, -q]), unit_rows(30, 384, 79)]).astype(np.float32)
    idx = build(pkg, Xa, path=2)
    for k in (10, 100):
        check(oracle, idx, Xa, Qa, k, q_bf16=True)
    idx.close()


@pytest.mark.parametrize("k", [10, 100])
def test_tensor_many_query_blocks(pkg, oracle, k):
    """Several 256-query blocks share every corpus chunk (work items = chunk x query block), with
    small chunks so that each CTA pair switches query blocks many times."""
    X, Q = unit_rows(150000, 384, 31), unit_rows(700, 384, 32)
    for chunk in (0, 3):
        idx = build(pkg, X, path=2)
        if chunk:
            idx.set_option("tc_chunk_tiles", chunk)
        check(oracle, idx, X, Q, k, q_bf16=True)
        st = idx.stats()
        assert st["path"] == 2 and st["seeded"] == 1 and st["passes"] == 3, st
        idx.close()


def test_tensor_more_queries_than_one_round(pkg, oracle):
    """nq above the 4096-query workspace round."""
    X, Q = unit_rows(5000, 384, 41), unit_rows(4200, 384, 42)
    idx = build(pkg, X, path=2)
    check(oracle, idx, X, Q, 10, q_bf16=True)
    idx.close()


def test_tensor_list_overflow_without_seed(pkg, oracle):
    """No sampled thresholds: every thread-private list fills and is compacted repeatedly."""
    X, Q = unit_rows(60000, 384, 51), unit_rows(50, 384, 52)
    idx = build(pkg, X, path=2, seed=0)
    for k in (1, 10, 64, 300):
        check(oracle, idx, X, Q, k, q_bf16=True)
        assert idx.stats()["seeded"] == 0
    idx.close()


def test_scan_and_tensor_paths_agree(pkg, oracle):
    X, Q = unit_rows(100000, 384, 91), unit_rows(16, 384, 92)
    a = build(pkg, X, path=1)
    b = build(pkg, X, path=2)
    Da, Ia = a.search(Q, 10)
    Db, Ib = b.search(Q, 10)
    Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
    for D, I in ((Da, Ia), (Db, Ib)):
        rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL_BF16)
        assert rep["ok"], rep
    a.close()
    b.close()


def test_tensor_other_dims(pkg, oracle):
    for d in (128, 256, 512, 768, 1024):
        X, Q = unit_rows(5000, d, d + 3), unit_rows(40, d, d + 4)
        idx = build(pkg, X, path=2)
        D, I = idx.search(Q, 10)
        # dims above 512 are served by the scan kernel, which keeps the query in fp32
        Xb, Qb = oracle.round_bf16(X), (oracle.round_bf16(Q) if idx.stats()["path"] == 2 else Q)
        Dr, Ir = oracle.flat_ip_topk(Xb, Qb, 10)
        rep = oracle.compare_topk(D, I, Dr, Ir, Xb, Qb, tie_tol=TIE_TOL_F32)
        assert rep["ok"], (d, rep)
        idx.close()


@pytest.mark.parametrize("k", [10, 100])
def test_pdl_overlap_mode_back_to_back(pkg, oracle, k):
    """Option pdl=2: the scan of call i+1 runs while the merge of call i still reads the shared
    candidate workspace; 200 back-to-back device calls must give exactly the serialised answers."""
    import torch
    X = unit_rows(300000, 384, 71)
    dev = torch.device("cuda", 0)
    idx = build(pkg, X, path=1)
    Q = torch.from_numpy(unit_rows(200, 384, 72)).to(dev)
    ref_s, ref_i = [], []
    for i in range(200):
        s, ids = idx.search_device(Q[i:i + 1], k)
        ref_s.append(s.clone())
        ref_i.append(ids.clone())
    torch.cuda.synchronize()
    for mode in (2, 0):
        idx.set_option("pdl", mode)
        outs = [(torch.empty((1, k), dtype=torch.float32, device=dev), torch.empty((1, k), dtype=torch.int64, device=dev))
                for _ in range(200)]
        for i in range(200):
            idx.search_device(Q[i:i + 1], k, out=outs[i])
        torch.cuda.synchronize()
        for i in range(200):
            assert torch.equal(outs[i][1], ref_i[i]), (mode, i)
            assert torch.equal(outs[i][0], ref_s[i]), (mode, i)
    idx.set_option("pdl", 1)
    D, I = idx.search(Q[:4].cpu().numpy(), k)
    Dr, Ir = oracle.flat_ip_topk(X, Q[:4].cpu().numpy(), k)
    assert oracle.compare_topk(D, I, Dr, Ir, X, Q[:4].cpu().numpy(), tie_tol=TIE_TOL_BF16)["ok"]
    idx.close()


@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("k", [10, 100])
def test_keep_fp32_rescoring_gives_flat_ip_ids_exactly(pkg, oracle, path, k):
    """keep_fp32: k + pad bf16 candidates are re-ranked with the fp32 rows, so ids are IDENTICAL to
    the fp32 flat oracle (faiss.IndexFlatIP semantics), not merely equal up to bf16 near-ties."""
    X, Q = unit_rows(120000, 384, 131), unit_rows(96, 384, 132)
    X[70000] = X[5]                                  # exact duplicate: tie must resolve to the lower id
    Q[0] = X[5]
    idx = pkg.FlatIPIndex(384, metric="inner_product", keep_fp32=True)
    idx.set_option("path", path)
    idx.add(X)
    D, I = idx.search(Q, k)
    Dr, Ir = oracle.flat_ip_topk(X, Q, k)            # fp64-accumulated adjudicator on the fp32 rows
    agree = (I == Ir).all(axis=1).mean()
    assert agree >= 0.98, agree                      # fp32 vs fp64 summation may still flip a 1e-7 tie
    rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=2e-6)
    assert rep["ok"] and rep["max_abs_score_err"] < 2e-6, rep
    assert list(I[0][:2]) == [5, 70000]
    # without the fp32 copy the same search shows bf16 near-tie swaps on this data
    plain = build(pkg, X, path=path)
    Dp, Ip = plain.search(Q, k)
    assert (Ip == Ir).all(axis=1).mean() <= agree
    plain.close()
    idx.close()


def test_keep_fp32_cosine_unnormalised_inputs(pkg, oracle):
    rng = np.random.default_rng(9)
    Xr = (unit_rows(30000, 384, 141) * rng.uniform(0.5, 3.0, (30000, 1))).astype(np.float32)
    Qr = (unit_rows(20, 384, 142) * rng.uniform(0.5, 3.0, (20, 1))).astype(np.float32)
    idx = pkg.FlatIPIndex(384, metric="cosine", keep_fp32=True)
    idx.add(Xr)
    Xn = Xr / np.linalg.norm(Xr, axis=1, keepdims=True)
    Qn = Qr / np.linalg.norm(Qr, axis=1, keepdims=True)
    Dr, Ir = oracle.flat_ip_topk(Xn, Qn, 10)
    for q in (Qr[:1], Qr):                            # scan path and tensor path
        D, I = idx.search(q, 10)
        rep = oracle.compare_topk(D, I, Dr[:len(q)], Ir[:len(q)], Xn, Qn[:len(q)], tie_tol=5e-6)
        assert rep["ok"] and rep["max_abs_score_err"] < 5e-6, rep
    idx.close()


@pytest.mark.parametrize("single", [0, 1])
def test_tensor_small_batch_both_mma_shapes(pkg, oracle, single):
    """<= 128 queries run as single-CTA MMAs (M = 128, option tc_single_cta=1, default) or as CTA pairs
    with padding queries (0); both must agree with the oracle and with each other."""
    X, Q = unit_rows(200000, 384, 151), unit_rows(100, 384, 152)
    idx = build(pkg, X, path=2, tc_single_cta=single)
    for nq, k in ((3, 10), (100, 10), (64, 100), (128, 300)):
        q = np.concatenate([Q, Q[:28]])[:nq]
        check(oracle, idx, X, q, k, q_bf16=True)
        assert idx.stats()["path"] == 2
    idx.close()


def test_randomised_shapes_auto_path(pkg, oracle):
    """Seeded sweep over (rows, queries, k, metric, duplicates) with the default kernel choice: whatever
    combination of scan / single-CTA / pair kernels, seeding, shared thresholds and merges a shape
    selects must satisfy the parity rule."""
    rng = np.random.default_rng(20260101)
    for case in range(36):
        n = int(rng.choice([1, 7, 255, 256, 257, 1000, 4097, 20000, 65536, 150001]))
        nq = int(rng.choice([1, 2, 3, 4, 17, 128, 129, 255, 256, 257, 513]))
        k = int(rng.choice([1, 2, 10, 31, 32, 33, 100, 257, 1000]))
        cosine = bool(rng.integers(0, 2))
        X, Q = unit_rows(n, 384, 1000 + case), unit_rows(nq, 384, 2000 + case)
        if n > 600 and rng.integers(0, 2):
            X[n // 3:n // 3 + 40] = X[5]                      # a block of exact duplicates
            Q[0] = X[5]
        if cosine:                                            # un-normalised inputs, cosine metric
            Xr = (X * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
            Qr = (Q * rng.uniform(0.5, 2.0, (nq, 1))).astype(np.float32)
            idx = build(pkg, Xr, metric="cosine")
            D, I = idx.search(Qr, k)
        else:
            idx = build(pkg, X)
            D, I = idx.search(Q, k)
        Dr, Ir = oracle.flat_ip_topk(X, Q, k)
        rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL_BF16)
        assert rep["ok"] and rep["max_abs_score_err"] <= TIE_TOL_BF16, (case, n, nq, k, cosine, rep)
        valid = I >= 0
        assert (valid.sum(axis=1) == min(k, n)).all(), (case, n, nq, k)
        idx.close()


@pytest.mark.parametrize("nq,k", [(1, 10), (64, 10), (300, 100)])
def test_search_is_cuda_graph_capturable(pkg, oracle, nq, k):
    """Once the workspace is warm a device search allocates nothing and only enqueues work on the
    caller's stream, so it can be captured in a CUDA graph and replayed on new query contents."""
    import torch
    X = unit_rows(60000, 384, 161)
    dev = torch.device("cuda", 0)
    idx = build(pkg, X)
    q = torch.from_numpy(unit_rows(nq, 384, 162)).to(dev)
    out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
    idx.search_device(q, k, out=out)                     # warm: workspaces, function attributes, tensor maps
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            idx.search_device(q, k, out=out)
    for rep in range(3):
        Qn = unit_rows(nq, 384, 170 + rep)
        q.copy_(torch.from_numpy(Qn).to(dev))
        g.replay()
        torch.cuda.synchronize()
        Dr, Ir = oracle.flat_ip_topk(X, Qn, k)
        repo = oracle.compare_topk(out[0].cpu().numpy(), out[1].cpu().numpy(), Dr, Ir, X, Qn, tie_tol=TIE_TOL_BF16)
        assert repo["ok"], (rep, repo)
    idx.close()


def test_search_handler_consumer_loop(pkg, oracle, tmp_path):
    """The reference's /search handler consumes the index like this (/root/reference/src/serve/app.py:293-317):
    `distances, indices = index_builder.search(query_emb, k=k)`, then for (dist, idx) in zip(distances[0],
    indices[0]): skip `idx < 0`, `doc_ids[idx]`, `float(dist)`.  Run that loop on our index: after a
    save()/load() round trip (app.py:427-433), with k > ntotal (the -1 padding the guard exists for)."""
    X = unit_rows(50, 384, 171)
    doc_ids = [f"doc_{i:03d}" for i in range(50)]
    built = pkg.FAISSIndexBuilder(embedding_dim=384, index_type="HNSW", metric="cosine")
    built.add(X, doc_ids)
    built.save(tmp_path / "index")
    index_builder = pkg.FAISSIndexBuilder(embedding_dim=384)          # app.py:427-429
    index_builder.load(tmp_path / "index")                            # app.py:430-433
    app_doc_ids = index_builder.doc_ids
    query_emb = unit_rows(1, 384, 172)                                # encode_queries([...]) -> [1, 384] fp32
    Dr, Ir = oracle.flat_ip_topk(X, query_emb, 50)
    for k in (10, 50, 100):                                           # schemas.py:12 allows k up to 100
        distances, indices = index_builder.search(query_emb, k=k)
        results = []
        for rank, (dist, idx) in enumerate(zip(distances[0], indices[0]), 1):
            if idx < 0 or (app_doc_ids and idx >= len(app_doc_ids)):
                continue
            doc_id = app_doc_ids[idx] if app_doc_ids else f"doc_{idx}"
            results.append({"doc_id": doc_id, "score": float(dist), "rank": rank})
        assert len(results) == min(k, 50)
        assert [r["rank"] for r in results] == list(range(1, len(results) + 1))
        assert all(isinstance(r["score"], float) and -1.01 <= r["score"] <= 1.01 for r in results)
        assert results[0]["doc_id"] == doc_ids[int(Ir[0, 0])]
        assert sorted(r["score"] for r in results) == [r["score"] for r in results][::-1]
        if k > 50:
            assert list(indices[0][50:]) == [-1] * (k - 50)
    built.close()
    index_builder.close()


def test_packed_merge_single_gpu(pkg, oracle):
    """b2s_merge_packed_device (the NCCL all-gather variant's K4) on one GPU: G shard results written into
    the packed [ids | scores] blocks exactly as ShardedFlatIPIndex lays them out, merged, compared."""
    import ctypes
    import torch
    from semantic_search_kd_b200.sharded import packed_bytes, shard_range
    L = pkg._lib.lib()
    dev = torch.device("cuda", 0)
    n, nq, k, G = 30011, 9, 20, 4
    X, Q = unit_rows(n, 384, 181), unit_rows(nq, 384, 182)
    X[20000] = X[3]                                              # a tie across shards
    Q[0] = X[3]
    q = torch.from_numpy(Q).to(dev)
    per = packed_bytes(nq, k)
    assert per == L.b2s_packed_bytes(nq, k)
    gathered = torch.zeros((G, per), dtype=torch.uint8, device=dev)
    shards = []
    for r in range(G):
        lo, hi = shard_range(n, G, r)
        sh = build(pkg, X[lo:hi])
        sh.set_id_offset(lo)
        ids = gathered[r][: nq * k * 8].view(torch.int64).view(nq, k)
        scores = gathered[r][nq * k * 8: nq * k * 12].view(torch.float32).view(nq, k)
        sh.search_device(q, k, out=(scores, ids))
        shards.append(sh)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    rc = L.b2s_merge_packed_device(0, ctypes.c_void_p(gathered.data_ptr()), G, nq, k, ctypes.c_void_p(out_s.data_ptr()),
                                   ctypes.c_void_p(out_i.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    assert rc == 0, pkg._lib.last_error()
    torch.cuda.synchronize()
    whole = build(pkg, X)
    D, I = whole.search(Q, k)
    assert np.array_equal(out_i.cpu().numpy(), I) and np.allclose(out_s.cpu().numpy(), D, atol=1e-6)
    assert list(I[0][:2]) == [3, 20000]
    for sh in shards + [whole]:
        sh.close()


def test_small_api_surface(pkg):
    import ctypes
    L = pkg._lib.lib()
    idx = build(pkg, unit_rows(100, 384, 1))
    h = idx._h
    assert L.b2s_dim(h) == 384 and L.b2s_ntotal(h) == 100 and L.b2s_device_count() >= 1
    assert L.b2s_get_option(h, b"path") == 0 and L.b2s_get_option(h, b"num_sms") > 100
    assert L.b2s_get_option(h, b"no_such_option") == -1
    assert L.b2s_set_option(h, b"no_such_option", 1) == pkg._lib.B2S_ERR_INVALID
    assert L.b2s_set_option(h, b"path", 7) == pkg._lib.B2S_ERR_INVALID
    assert L.b2s_reserve(h, 5000) == 0 and L.b2s_ntotal(h) == 100
    D0, I0 = idx.search(unit_rows(2, 384, 2), 5)
    assert L.b2s_reset(h) == 0 and L.b2s_ntotal(h) == 0
    D, I = idx.search(unit_rows(2, 384, 2), 5)
    assert (I == -1).all()                                      # empty index: padding only
    idx.add(unit_rows(100, 384, 1))
    D1, I1 = idx.search(unit_rows(2, 384, 2), 5)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    idx.close()


def test_nan_rows_and_queries_do_not_crash(pkg, oracle):
    """A NaN score fails every threshold test: NaN rows never appear, a NaN query returns only padding."""
    X, Q = unit_rows(5000, 384, 191), unit_rows(6, 384, 192)
    Xn = X.copy()
    Xn[17, 5] = np.nan
    Xn[4000] = np.nan
    for path in (1, 2):
        idx = build(pkg, Xn, path=path)
        D, I = idx.search(Q, 10)
        assert not np.isin(I, [17, 4000]).any() and np.isfinite(D).all()
        Xc = X.copy()
        Xc[[17, 4000]] = 0.0                              # the oracle sees them as rows that cannot win
        Dr, Ir = oracle.flat_ip_topk(Xc, Q, 10)
        assert oracle.compare_topk(D, I, Dr, Ir, Xc, Q, tie_tol=TIE_TOL_BF16)["ok"]
        qn = Q.copy()
        qn[0, 0] = np.nan
        D, I = idx.search(qn, 10)
        assert (I[0] == -1).all() and (I[1:] >= 0).all()
        idx.close()
