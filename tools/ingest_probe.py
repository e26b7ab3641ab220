#!/usr/bin/env python3
"""tools/ingest_probe.py -- rate of building an index from HOST rows (FlatIPIndex.add(np.ndarray), fp32 and bf16):
rows/s and GB/s through b2s_add_f32 / b2s_add_bf16 (pinned double-buffered staging, csrc/b2s_api.cu add_impl)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import semantic_search_kd_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(0)
X = rng.standard_normal((n, 384), dtype=np.float32)
X /= np.linalg.norm(X, axis=1, keepdims=True)
for metric in ("inner_product", "cosine"):
    idx = pkg.FlatIPIndex(384, metric=metric, device=0)
    idx.reserve(n)
    idx.add(X[:1000])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx.add(X)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    D, I = idx.search(X[12345:12346], 3)
    print(f"{metric}: add({n} x 384 fp32 host rows) {dt:.3f} s = {n / dt / 1e6:.2f} M rows/s = {X.nbytes / dt / 1e9:.2f} GB/s; "
          f"self-retrieval id {I[0, 0] - 1000} score {D[0, 0]:.4f}")
    idx.close()
