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
        if len(np.unique(s_ref)) == len(s_ref) and tol == 0:
            assert np.array_equal(I[i], ref_ids[i]), (what, i, I[i], ref_ids[i])
        elif len(np.unique(np.round(s_ref / max(tol, 1e-12)))) == len(s_ref):
            assert np.array_equal(I[i], ref_ids[i]), (what, i, I[i], ref_ids[i])


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
            # bf16 storage: ids may swap only between scores closer than the north-star tolerance
            assert_same_retrieval(I, z[key], X, Q, 1e-3, f"{key} path {path}")
        idx.close()
