"""Parity against vectors produced by RUNNING the reference's own code (tests/golden/make_ref_golden.py:
/root/reference/src/kd/eval.py KDEvaluator.evaluate_retrieval, scripts/simple_eval.py evaluate_model,
src/mining/miners.py ANCEMiner.mine -- imported unmodified, model stubbed).  Nothing here reads
/root/reference: inputs are regenerated from the seeds and checked by SHA-256.

CPU part: the oracle must retrieve what the reference retrieved.  GPU part (-m gpu): so must the
CUDA path, through the C ABI.  np.argsort leaves the order of EXACT ties unspecified (the
reference reverses an ascending sort, so equal scores come out by descending id; faiss and this
library keep the lower id first), so rows are compared position by position on scores and as sets
on ids, and ids must be identical wherever the scores of the row are distinct."""
import hashlib
import json
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, str(GOLDEN))


@pytest.fixture(scope="module")
def ref_case():
    import make_ref_golden as mr
    meta = json.loads((GOLDEN / "ref_ance.json").read_text())
    X, Q = mr.unit_rows(meta["n"], meta["seed_x"]), mr.unit_rows(meta["nq"], meta["seed_q"])
    X[100:104] = X[7]
    X[4000] = Q[3]
    assert hashlib.sha256(X.tobytes()).hexdigest() == meta["sha_X"], "corpus bytes differ from the generator's"
    assert hashlib.sha256(Q.tobytes()).hexdigest() == meta["sha_Q"], "query bytes differ from the generator's"
    return X, Q, np.load(GOLDEN / "ref_eval.npz"), meta


def assert_same_retrieval(I, ref_ids, X, Q, tol, what):
    """ids equal where the row has no exact tie among the reference's scores; score sequence equal always."""
    Xd, Qd = X.astype(np.float64), Q.astype(np.float64)
    for i in range(ref_ids.shape[0]):
        s_ref = np.array([Qd[i] @ Xd[j] for j in ref_ids[i]])
        s_our = np.array([Qd[i] @ Xd[j] for j in I[i]])
        np.testing.assert_allclose(s_our, s_ref, rtol=0, atol=tol, err_msg=f"{what} row {i}")
        gaps = np.abs(np.diff(np.sort(s_ref))) if len(s_ref) > 1 else np.array([np.inf])
        if gaps.min() > 2 * tol:   # no (near-)tie in this row: the ids themselves must be identical
            assert np.array_equal(I[i], ref_ids[i]), (what, i, I[i], ref_ids[i])
        else:
            # (near-)ties may swap; every id must still be one the reference could have returned
            kth = s_ref[-1]
            for j, s in zip(I[i], s_our):
                assert j in ref_ids[i] or abs(s - kth) <= max(tol, 1e-12), (what, i, j, s, kth)


def test_oracle_retrieves_what_the_reference_retrieved(oracle, ref_case):
    X, Q, z, _ = ref_case
    for key in [f for f in z.files if f.startswith(("eval_ids_k", "simple_ids_k"))]:
        k = int(key.split("_k")[1])
        for fn in (oracle.flat_ip_topk, oracle.flat_ip_topk_np):
            D, I = fn(X, Q, k)
            assert_same_retrieval(I, z[key], X, Q, 0, key)
    # the two reference entry points agree with each other
    for k in (1, 5, 10):
        assert np.array_equal(z[f"eval_ids_k{k}"], z[f"simple_ids_k{k}"])


def test_oracle_ance_filter_matches_reference(oracle, ref_case):
    X, Q, _, meta = ref_case
    for name, ref_negs in meta["negatives"].items():
        margin, top_k = float(name.split("_")[0][len("margin"):]), int(name.split("_top")[1])
        for i in range(meta["nq"]):
            pos = [int(s[1:]) for s in meta["positives"][i]]
            cand = [int(s[1:]) for s in meta["candidates"][i]]
            ps = np.matmul(Q[i:i + 1], X[pos].T)[0]
            cs = np.matmul(Q[i:i + 1], X[cand].T)[0]
            got = oracle.ance_filter_ref(meta["candidates"][i], cs, ps, margin, top_k)
            assert got == ref_negs[i], (name, i)


@pytest.mark.gpu
def test_cuda_path_retrieves_what_the_reference_retrieved(ref_case):
    import semantic_search_kd_b200 as pkg
    X, Q, z, _ = ref_case
    for path in (1, 2):
        idx = pkg.FlatIPIndex(384, metric="inner_product")
        idx.set_option("path", path)
        idx.add(X)
        for key in [f for f in z.files if f.startswith("eval_ids_k")]:
            k = int(key.split("_k")[1])
            D, I = idx.search(Q, k)
            assert idx.stats()["path"] == path
            # north-star parity rule against the ids the REFERENCE retrieved: an id only one side
            # returned must score within 1e-3 of the k-th score (bf16 storage may swap near-ties)
            ref_ids = z[key]
            ref_scores = np.array([[Q[i].astype(np.float64) @ X[j].astype(np.float64) for j in ref_ids[i]]
                                   for i in range(len(Q))], dtype=np.float32)
            from oracle import oracle as orc
            rep = orc.compare_topk(D, I, ref_scores, ref_ids, X, Q, tie_tol=1e-3)
            assert rep["ok"] and rep["max_abs_score_err"] <= 1e-3, (key, path, rep)
            assert rep["set_match"] >= len(Q) // 2, (key, path, rep)   # swaps only where the k-th score is a near-tie
        idx.close()


@pytest.mark.gpu
def test_cuda_ance_miner_matches_reference(ref_case):
    """ANCEMiner.mine mirror (similarity on the GPU through b2s_similarity) returns exactly the hard
    negatives the reference's ANCEMiner.mine returned on the same stubbed model."""
    import make_ref_golden as mr
    import semantic_search_kd_b200 as pkg
    X, Q, _, meta = ref_case
    stub = mr.StubStudent(X, Q)
    queries = [f"q{i}" for i in range(meta["nq"])]
    texts = {f"d{i}": f"d{i}" for i in range(meta["n"])}
    for name, ref_negs in meta["negatives"].items():
        margin, top_k = float(name.split("_")[0][len("margin"):]), int(name.split("_top")[1])
        got = pkg.ANCEMiner(stub, margin=margin).mine(queries, meta["positives"], meta["candidates"], texts, texts,
                                                       top_k=top_k)
        assert got == ref_negs, name


@pytest.mark.gpu
@pytest.mark.parametrize("nq,top_k", [(40, 20), (300, 200)])
def test_cuda_corpus_wide_ance(oracle, nq, top_k):
    """mine_corpus = exact search + the reference's selection rule (miners.py:237-247) with the
    query's positives removed, checked against the oracle restatement on the same data."""
    import semantic_search_kd_b200 as pkg
    from conftest import unit_rows
    n, margin = 30000, 0.05
    X, Q = unit_rows(n, 384, 61), unit_rows(nq, 384, 62)
    Xb = oracle.round_bf16(X)
    rng = np.random.default_rng(5)
    D0, I0 = oracle.flat_ip_topk(Xb, Q, 4)
    positives = []
    for i in range(nq):
        p = [int(I0[i, 1])] if i % 3 else []                      # a strong positive (rank 2) or none
        if i % 5 == 0:
            p.append(int(rng.integers(0, n)))                     # plus a random (weak) one
        positives.append(p)
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    idx.add(X)
    ids, scores, counts = pkg.ANCEMiner(None, margin=margin).mine_corpus(idx, Q, positives, top_k=top_k)
    path = idx.stats()["path"]
    Qe = oracle.round_bf16(Q) if path == 2 else Q
    n_pos = max(1, max(len(p) for p in positives))
    Dr, Ir = oracle.flat_ip_topk(Xb, Qe, top_k + n_pos)
    for i in range(nq):
        ps = [float(Qe[i].astype(np.float64) @ Xb[p].astype(np.float64)) for p in positives[i]]
        thr = (max(ps) if ps else 0.0) - margin
        sure, maybe = [], set()
        for s, d in zip(Dr[i], Ir[i]):
            if d < 0 or int(d) in positives[i]:
                continue
            if s >= thr + 1e-4:
                sure.append(int(d))
            elif s >= thr - 1e-4:
                maybe.add(int(d))
        got = [int(x) for x in ids[i][:counts[i]]]
        assert all(x == -1 for x in ids[i][counts[i]:])
        assert not (set(got) & set(positives[i]))
        exp = sure[:top_k]
        if len(sure) >= top_k:
            # ids may swap only between scores within fp32 summation noise of each other
            assert set(got) - set(exp) <= set(sure) | maybe and len(got) == top_k, i
            s_exp = [float(Qe[i].astype(np.float64) @ Xb[d].astype(np.float64)) for d in exp]
            s_got = [float(Qe[i].astype(np.float64) @ Xb[d].astype(np.float64)) for d in got]
            np.testing.assert_allclose(s_got, s_exp, atol=2e-5)
        else:
            assert set(exp) <= set(got) <= set(exp) | maybe, (i, got, exp, maybe)
        assert np.all(np.diff(scores[i][:counts[i]]) <= 0)
    idx.close()


def test_maxsim_oracle_matches_reference(ref_case, oracle):
    """The oracle's MaxSim restatement == the reference's own maxsim_aggregation output (tests/golden/ref_maxsim.json);
    the product is the device kernel checked below, there is no host copy of the function in the package."""
    X, Q, z, meta = ref_case
    ref = json.loads((GOLDEN / "ref_maxsim.json").read_text())
    Xd, Qd = X.astype(np.float64), Q.astype(np.float64)
    for i in range(meta["nq"]):
        hits = z["eval_ids_k20"][i]
        chunk_scores = [(f"doc{int(r) // 3}_{int(r) % 3}", float(Qd[i] @ Xd[int(r)])) for r in hits]
        assert oracle.maxsim_ref(chunk_scores) == ref[i]


@pytest.mark.gpu
def test_cuda_maxsim_matches_reference(ref_case):
    """Device MaxSim over the CUDA search result == the reference's maxsim_aggregation over the chunks
    the reference retrieved (same documents, scores within the bf16 tolerance)."""
    import semantic_search_kd_b200 as pkg
    X, Q, z, meta = ref_case
    ref = json.loads((GOLDEN / "ref_maxsim.json").read_text())
    idx = pkg.FlatIPIndex(384, metric="inner_product")
    idx.add(X)
    chunk_to_doc = np.arange(len(X), dtype=np.int64) // 3
    S, D = pkg.maxsim_topk(idx, Q, 20, chunk_to_doc, k_chunks=20)
    for i in range(meta["nq"]):
        got = {f"doc{int(d)}": float(s) for s, d in zip(S[i], D[i]) if d >= 0}
        exp = ref[i]
        kth = min(exp.values())
        for doc in set(got) ^ set(exp):                      # only near-ties of the 20th chunk may differ
            s = got.get(doc, exp.get(doc))
            assert abs(s - kth) <= 2e-3, (i, doc, s, kth)
        for doc in set(got) & set(exp):
            assert abs(got[doc] - exp[doc]) <= 1e-3, (i, doc)
        assert list(S[i][:len(got)]) == sorted(S[i][:len(got)], reverse=True)
    idx.close()
