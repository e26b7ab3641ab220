#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ (run from the repo root, CPU only).

The reference holds no golden vector for this path (the
only faiss call sites are an unused fixture, ``/root/reference/tests/conftest.py:175-200``),
so the vectors are produced here from

* the reference's OWN expression for exact search, executed verbatim:
  ``scores = q @ c.T`` (``src/kd/eval.py:75`` via ``compute_similarity``) and
  ``np.argsort(scores)[::-1][:k]`` (``src/kd/eval.py:86``,
  ``scripts/simple_eval.py:35``)                       -> keys ``ref_expr_*``
* the numpy oracle ``oracle.oracle.flat_ip_topk_np`` (fp64 accumulate, ties by
  ascending id, (-FLT_MAX,-1) padding)                  -> keys ``ids_* / scores_*``

on the reference's fixture recipe (``tests/conftest.py:65-73``:
``np.random.seed(42); randn(10,384)``, unit-normalised) and on seeded larger
cases whose inputs are regenerated from the seed and checked by SHA-256.
"""
from __future__ import annotations

import hashlib
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent


def unit(a: np.ndarray) -> np.ndarray:
    """L2-normalise rows; the norm uses math.fsum (exactly rounded, order-independent) so the
    bytes -- and the SHA-256 in manifest.json -- do not depend on the CPU's SIMD reduction."""
    a = np.asarray(a, dtype=np.float64)
    nrm = np.array([math.sqrt(math.fsum((row * row).tolist())) for row in a])
    return (a / nrm[:, None]).astype(np.float32)


def conftest_embeddings() -> np.ndarray:
    # /root/reference/tests/conftest.py:65-73, verbatim recipe
    np.random.seed(42)
    e = np.random.randn(10, 384).astype(np.float32)
    return e / np.linalg.norm(e, axis=1, keepdims=True)


def seeded_case(n: int, nq: int, seed: int, dup: bool = False):
    rng = np.random.default_rng(seed)
    X = unit(rng.standard_normal((n, 384)))
    Q = unit(rng.standard_normal((nq, 384)))
    if dup:
        # exact duplicates -> exact score ties; queries near corpus rows -> scores ~ 1
        X[n // 2:n // 2 + 8] = X[3]
        X[n - 5:] = X[7]
        Q[0] = X[3]
        Q[1] = X[7]
        Q[2] = unit((X[11] + 0.05 * rng.standard_normal(384)).reshape(1, -1))[0]
    return X, Q


def ref_expr(X: np.ndarray, Q: np.ndarray, k: int) -> np.ndarray:
    """The reference's expression, as written in src/kd/eval.py:75,86."""
    out = np.full((Q.shape[0], k), -1, dtype=np.int64)
    for i in range(Q.shape[0]):
        scores = np.matmul(Q[i:i + 1], X.T)[0]
        top_k_indices = np.argsort(scores)[::-1][:k]
        out[i, :len(top_k_indices)] = top_k_indices
    return out


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    # --- case A: the reference fixture (inputs stored, 15 KB) -------------------
    X = conftest_embeddings()
    rng = np.random.default_rng(7)
    Q = np.concatenate([X[[0, 4, 9]], unit(rng.standard_normal((5, 384)))]).astype(np.float32)
    payload = {"X": X, "Q": Q}
    for k in (1, 5, 10, 12):
        D, I = orc.flat_ip_topk_np(X, Q, k)
        payload[f"ids_k{k}"] = I
        payload[f"scores_k{k}"] = D
        payload[f"ref_expr_k{k}"] = ref_expr(X, Q, k)
    np.savez_compressed(OUT / "conftest_fixture.npz", **payload)

    # --- seeded cases: outputs + input hashes only ------------------------------
    manifest = {}
    for name, (n, nq, seed, dup) in {
        "rand2000": (2000, 16, 1234, False),
        "dups3000": (3000, 12, 99, True),
    }.items():
        X, Q = seeded_case(n, nq, seed, dup)
        Xb = orc.round_bf16_np(X)
        payload = {}
        for k in (10, 100):
            D, I = orc.flat_ip_topk_np(X, Q, k)
            Db, Ib = orc.flat_ip_topk_np(Xb, Q, k)
            payload[f"ids_k{k}"] = I
            payload[f"scores_k{k}"] = D
            payload[f"ids_bf16corpus_k{k}"] = Ib
            payload[f"scores_bf16corpus_k{k}"] = Db
            payload[f"ref_expr_k{k}"] = ref_expr(X, Q, k)
        np.savez_compressed(OUT / f"{name}.npz", **payload)
        manifest[name] = {"n": n, "nq": nq, "seed": seed, "dup": dup, "d": 384,
                          "sha256_X": sha(X), "sha256_Q": sha(Q)}
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")), "manifest.json")


if __name__ == "__main__":
    main()
