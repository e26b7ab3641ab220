"""CPU: the oracle against the committed golden vectors and against itself (numpy vs C)."""
import numpy as np
import pytest

from conftest import unit_rows


def test_oracle_matches_golden(oracle, golden_cases):
    for name, (X, Q, z) in golden_cases.items():
        ks = sorted(int(k[5:]) for k in z.files if k.startswith("ids_k"))
        assert ks, name
        for k in ks:
            D, I = oracle.flat_ip_topk(X, Q, k)            # C oracle, fp64 accumulate
            assert np.array_equal(I, z[f"ids_k{k}"]), (name, k)
            np.testing.assert_allclose(D, z[f"scores_k{k}"], rtol=0, atol=2e-7)
            Dn, In = oracle.flat_ip_topk_np(X, Q, k)        # numpy restatement
            assert np.array_equal(In, z[f"ids_k{k}"]), (name, k)


def test_golden_agrees_with_reference_expression(golden_cases):
    """ids from the reference's own `argsort(q @ c.T)[::-1][:k]` (src/kd/eval.py:75,86) equal the
    oracle's wherever there is no exact score tie (argsort's tie order is unspecified)."""
    for name, (X, Q, z) in golden_cases.items():
        for key in [k for k in z.files if k.startswith("ref_expr_k")]:
            k = int(key[len("ref_expr_k"):])
            ref, ours, sc = z[key], z[f"ids_k{k}"], z[f"scores_k{k}"]
            for i in range(ref.shape[0]):
                valid = ours[i] >= 0
                if len(np.unique(sc[i][valid])) == valid.sum():   # no ties in this row
                    assert np.array_equal(ref[i][valid], ours[i][valid]), (name, k, i)
                else:
                    assert set(ref[i][ref[i] >= 0]) == set(ours[i][valid]) or True
            # score multiset must agree regardless of ties
            for i in range(ref.shape[0]):
                r = ref[i][ref[i] >= 0]
                s_ref = np.sort((Q[i].astype(np.float64) @ X[r].astype(np.float64).T))[::-1]
                s_our = sc[i][ours[i] >= 0].astype(np.float64)
                np.testing.assert_allclose(s_ref, s_our, atol=1e-6)


def test_bf16_corpus_golden(oracle, golden_cases):
    for name in ("rand2000", "dups3000"):
        X, Q, z = golden_cases[name]
        Xb = oracle.round_bf16(X)
        assert np.array_equal(Xb, oracle.round_bf16_np(X))
        for k in (10, 100):
            D, I = oracle.flat_ip_topk(Xb, Q, k)
            assert np.array_equal(I, z[f"ids_bf16corpus_k{k}"])


def test_tie_order_and_padding(oracle):
    X = unit_rows(7, 384, 3)
    X[5] = X[1]
    X[6] = X[1]
    Q = X[[1]]
    for fn in (oracle.flat_ip_topk, oracle.flat_ip_topk_np):
        D, I = fn(X, Q, 10)
        assert list(I[0][:3]) == [1, 5, 6]                 # equal scores: ascending id
        assert list(I[0][7:]) == [-1, -1, -1]              # k > n: -1 padding (app.py:300)
        assert np.all(D[0][7:] == oracle.FLT_LOWEST)
        assert np.all(np.diff(D[0][:7]) <= 0)               # descending


def test_empty_inputs(oracle):
    X = unit_rows(5, 384, 0)
    D, I = oracle.flat_ip_topk(X, np.zeros((0, 384), np.float32), 4)
    assert D.shape == (0, 4) and I.shape == (0, 4)
    D, I = oracle.flat_ip_topk(np.zeros((0, 384), np.float32), X[:2], 3)
    assert np.all(I == -1)
    D, I = oracle.flat_ip_topk(X, X[:2], 0)
    assert D.shape == (2, 0)


def test_streaming_merge_equals_one_shot(oracle):
    X = unit_rows(5000, 384, 11)
    Q = unit_rows(9, 384, 12)
    D, I = oracle.flat_ip_topk(X, Q, 25)
    acc = None
    for s in range(0, 5000, 1300):
        acc = oracle.flat_ip_topk(X[s:s + 1300], Q, 25, id_offset=s, merge_into=acc)
    assert np.array_equal(acc[1], I)
    assert np.array_equal(acc[0], D)


def test_f32_and_f64_accumulate_agree_within_rounding(oracle):
    X = unit_rows(3000, 384, 5)
    Q = unit_rows(4, 384, 6)
    D32, I32 = oracle.flat_ip_topk(X, Q, 10, acc="f32")
    D64, I64 = oracle.flat_ip_topk(X, Q, 10, acc="f64")
    np.testing.assert_allclose(D32, D64, atol=5e-7)
    rep = oracle.compare_topk(D32, I32, D64, I64, X, Q, tie_tol=1e-6)
    assert rep["ok"], rep


def test_similarity_shape_and_range(oracle):
    """tests/test_student_model.py:104-124 of the reference: shape (2,3), values in [-1.01, 1.01]."""
    q, d = unit_rows(2, 384, 1), unit_rows(3, 384, 2)
    S = oracle.similarity_np(q, d)
    assert S.shape == (2, 3)
    assert np.all(S >= -1.01) and np.all(S <= 1.01)


def test_ance_filter_ref(oracle):
    ids = ["a", "b", "c", "d"]
    sc = [0.5, 0.9, 0.75, 0.2]
    assert oracle.ance_filter_ref(ids, sc, [0.8], 0.1, 5) == ["b", "c"]
    assert oracle.ance_filter_ref(ids, sc, [], 0.1, 2) == ["b", "c"]   # no positives: max_pos = 0.0
    assert oracle.ance_filter_ref(ids, sc, [2.0], 0.1, 5) == []


def test_parity_rule_detects_violations(oracle):
    X = unit_rows(500, 384, 21)
    Q = unit_rows(3, 384, 22)
    D, I = oracle.flat_ip_topk(X, Q, 10)
    ok = oracle.compare_topk(D, I, D, I, X, Q)
    assert ok["ok"] and ok["exact_order"] == 3 and ok["tie_swaps"] == 0
    bad_I = I.copy()
    worst = int(np.argmin(Q[0] @ X.T))
    bad_I[0, -1] = worst
    rep = oracle.compare_topk(D, bad_I, D, I, X, Q)
    assert not rep["ok"] and rep["violations"] >= 1


def test_hnsw_restatement_recall_small(oracle):
    """Low-dimensional sanity check of the HNSW baseline (clustered data, recall should be high)."""
    rng = np.random.default_rng(0)
    centers = rng.standard_normal((50, 32)).astype(np.float32)
    X = centers[rng.integers(0, 50, 4000)] + 0.3 * rng.standard_normal((4000, 32)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Q = X[rng.integers(0, 4000, 100)] + 0.05 * rng.standard_normal((100, 32)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    h = oracle.HnswRef(X, M=16, ef_construction=100, nthreads=2)
    _, Ih = h.search(Q, 10, ef_search=64)
    _, If = oracle.flat_ip_topk(X, Q, 10)
    assert oracle.recall_at_k(Ih, If) > 0.9


def test_hnsw_restatement_recall_grows_with_ef_search(oracle):
    """On isotropic 384-d data (the worst case for a graph index, and what BASELINE config 0 prescribes)
    the restated HNSW's recall@10 is low at efSearch=64 but rises monotonically to ~1 with efSearch:
    the graph is sound, the data is hard."""
    X, Q = oracle.gen_unit_rows(6000, 384, 0), oracle.gen_unit_rows(100, 384, 1)
    h = oracle.HnswRef(X, M=32, ef_construction=200, nthreads=4)
    _, If = oracle.flat_ip_topk(X, Q, 10, acc="f32")
    rec = []
    for ef in (16, 64, 256, 2048):
        _, Ih = h.search(Q, 10, ef_search=ef, nthreads=4)
        rec.append(oracle.recall_at_k(Ih, If))
    assert rec == sorted(rec) and rec[-1] >= 0.99 and rec[0] < rec[-1], rec


@pytest.mark.parametrize("case", ["conftest_fixture", "dups3000"])
def test_faiss_recorded_answers(oracle, case):
    """The oracle against answers RECORDED from faiss itself (tools/make_faiss_golden.py).  faiss is not in
    this image, so the fixtures do not exist yet and this skips; it pins the faiss half of the oracle on any
    checkout where someone with faiss has run the tool."""
    from conftest import GOLDEN
    p = GOLDEN / f"faiss_{case}.npz"
    if not p.exists():
        pytest.skip(f"{p.name} not recorded yet (needs a machine with faiss: tools/make_faiss_golden.py)")
    z = np.load(p)
    for k in (1, 10, 100):
        D, I = oracle.flat_ip_topk(z["X"], z["Q"], k, acc="f32")
        assert np.array_equal(I, z[f"I_k{k}"])
        ok = I >= 0
        assert np.allclose(D[ok], z[f"D_k{k}"][ok], atol=2e-6)
        assert np.all(D[~ok] == z[f"D_k{k}"][~ok])
