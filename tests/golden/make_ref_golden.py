#!/usr/bin/env python3
"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE in this container.

Run from the repo root, CPU only, with the reference mounted at /root/reference:

    python tests/golden/make_ref_golden.py

What is executed (unmodified, imported from /root/reference):

* ``src.kd.eval.KDEvaluator.evaluate_retrieval``  (src/kd/eval.py:42-101) -- the exact
  brute-force retrieval: one score row per query, ``np.argsort(scores)[::-1][:k]``.
* ``scripts.simple_eval.evaluate_model``           (scripts/simple_eval.py:16-49) -- the same
  search with the full ``np.matmul(query_embs, corpus_embs.T)`` matrix.
* ``src.mining.miners.ANCEMiner.mine``             (src/mining/miners.py:184-253) -- margin filter
  + descending sort + top-k over the candidate scores.
* ``src.utils.chunk.maxsim_aggregation``           (src/utils/chunk.py:123-148) -- max chunk score per
  document, applied to the chunks the reference retrieved (row r = chunk r % 3 of document r // 3).

What is NOT the reference's: ``src/models/student.py`` is absent from the tree,
so the model is a stub that returns seeded unit-norm 384-d embeddings for the strings it is given
and implements ``compute_similarity(q, d) = q @ d.T`` -- the behaviour pinned by
/root/reference/tests/test_student_model.py:104-124 (shape [nq, nd], cosine of unit vectors).
``rank_bm25`` (imported by src/data/bm25.py at module load, unused on this path) is stubbed too.

How the ids are captured: the reference only returns metrics, but it reads the relevance label of
each retrieved id in rank order (``labels[i] for i in top_k_indices``, eval.py:87;
simple_eval.py:36), so a list subclass that records ``__getitem__`` calls yields exactly the ids
the reference retrieved, in order, for every (query, k).

Outputs: tests/golden/ref_eval.npz, tests/golden/ref_ance.json, tests/golden/ref_maxsim.json (inputs are regenerated from the
seeds by the tests and checked by SHA-256).
"""
from __future__ import annotations

import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
DIM = 384


def unit_rows(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, DIM))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    return a.astype(np.float32)


class StubStudent:
    """Stands in for the absent src/models/student.py; texts are "q<i>" / "d<i>" keys into seeded tables."""

    def __init__(self, doc_embs: np.ndarray, query_embs: np.ndarray):
        self.docs, self.queries = doc_embs, query_embs
        self.embedding_dim = DIM

    @staticmethod
    def _ids(texts, prefix):
        if isinstance(texts, str):
            texts = [texts]
        return [int(t[len(prefix):]) for t in texts]

    def encode_queries(self, texts, **kw):
        return self.queries[self._ids(texts, "q")]

    def encode_documents(self, texts, **kw):
        ids = self._ids(texts, "d")
        return self.docs[ids] if ids else np.zeros((0, DIM), np.float32)

    def compute_similarity(self, q, d):
        return np.matmul(q, d.T)


class Recorder(list):
    """Relevance labels that remember which ids were looked up, in order."""

    def __init__(self, n):
        super().__init__([0] * n)
        self.seen = []

    def __getitem__(self, i):
        self.seen.append(int(i))
        return 1 if (int(i) % 7 == 0) else 0


def install_stubs():
    for name in ("src.models", "src.models.student", "src.models.teacher", "rank_bm25"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["src.models.student"].StudentModel = StubStudent
    sys.modules["src.models.teacher"].TeacherModel = object
    sys.modules["rank_bm25"].BM25Okapi = object
    sys.path.insert(0, str(REF))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    install_stubs()
    from src.kd.eval import KDEvaluator                       # /root/reference/src/kd/eval.py
    from src.mining.miners import ANCEMiner                   # /root/reference/src/mining/miners.py
    import scripts.simple_eval as simple_eval                 # /root/reference/scripts/simple_eval.py

    n, nq, seed_x, seed_q = 5000, 24, 2024, 2025
    X, Q = unit_rows(n, seed_x), unit_rows(nq, seed_q)
    # a few exact duplicates (ties) and near-duplicates of queries (scores ~ 1)
    X[100:104] = X[7]
    X[4000] = Q[3]
    stub = StubStudent(X, Q)
    queries = [f"q{i}" for i in range(nq)]
    corpus = [f"d{i}" for i in range(n)]
    k_values = [1, 5, 10, 20]

    payload = {}
    # --- KDEvaluator.evaluate_retrieval -------------------------------------------------
    labels = [Recorder(n) for _ in range(nq)]
    metrics = KDEvaluator(student=stub).evaluate_retrieval(queries, corpus, labels, k_values=k_values)
    for ki, k in enumerate(k_values):
        off = sum(k_values[:ki])
        payload[f"eval_ids_k{k}"] = np.array([lab.seen[off:off + k] for lab in labels], dtype=np.int64)
    payload["eval_metrics"] = np.array([[metrics[f"ndcg@{k}"], metrics[f"mrr@{k}"]] for k in k_values])
    # --- scripts/simple_eval.evaluate_model ---------------------------------------------
    labels2 = [Recorder(n) for _ in range(nq)]
    m2 = simple_eval.evaluate_model(stub, queries, corpus, labels2, k_values=[1, 5, 10])
    for ki, k in enumerate([1, 5, 10]):
        off = sum([1, 5, 10][:ki])
        payload[f"simple_ids_k{k}"] = np.array([lab.seen[off:off + k] for lab in labels2], dtype=np.int64)
    payload["simple_metrics"] = np.array([[m2[f"ndcg@{k}"], m2[f"mrr@{k}"]] for k in (1, 5, 10)])
    payload["sha_X"] = np.frombuffer(bytes.fromhex(sha(X)), dtype=np.uint8)
    payload["sha_Q"] = np.frombuffer(bytes.fromhex(sha(Q)), dtype=np.uint8)
    np.savez_compressed(OUT / "ref_eval.npz", **payload)

    # --- ANCEMiner.mine -------------------------------------------------------------------
    rng = np.random.default_rng(77)
    positives, candidates = [], []
    for i in range(nq):
        # positives: the query's true best doc and one random doc; candidates: a mix of near and far docs
        best = np.argsort(np.matmul(Q[i:i + 1], X.T)[0])[::-1]
        pos = [int(best[0]), int(rng.integers(0, n))]
        cand = [int(x) for x in best[1:13]] + [int(x) for x in rng.integers(0, n, 8)]
        positives.append([f"d{p}" for p in pos])
        candidates.append([f"d{c}" for c in cand])
    texts = {f"d{i}": f"d{i}" for i in range(n)}
    out = {}
    for margin, top_k in ((0.1, 5), (0.02, 5), (0.3, 20)):
        negs = ANCEMiner(stub, margin=margin).mine(queries, positives, candidates, texts, texts, top_k=top_k)
        out[f"margin{margin}_top{top_k}"] = negs
    # --- maxsim_aggregation (src/utils/chunk.py:123-148) on the reference's own retrieved chunks ----
    from src.utils.chunk import maxsim_aggregation          # /root/reference/src/utils/chunk.py
    Xd, Qd = X.astype(np.float64), Q.astype(np.float64)
    maxsim = []
    for i in range(nq):
        hits = payload["eval_ids_k20"][i]
        # corpus row r is chunk r % 3 of document r // 3
        chunk_scores = [(f"doc{int(r) // 3}_{int(r) % 3}", float(Qd[i] @ Xd[int(r)])) for r in hits]
        maxsim.append(maxsim_aggregation(chunk_scores))
    (OUT / "ref_maxsim.json").write_text(json.dumps(maxsim) + "\n")

    (OUT / "ref_ance.json").write_text(json.dumps(
        {"n": n, "nq": nq, "seed_x": seed_x, "seed_q": seed_q, "positives": positives, "candidates": candidates,
         "negatives": out, "sha_X": sha(X), "sha_Q": sha(Q)}, indent=0) + "\n")
    print("wrote ref_eval.npz, ref_ance.json;", {k: v for k, v in metrics.items()})


if __name__ == "__main__":
    main()
