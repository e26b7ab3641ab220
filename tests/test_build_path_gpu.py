"""GPU: the index BUILD path -- `FAISSIndexBuilder.build_from_parquet`
(/root/reference/scripts/build_faiss_index.py:49-72) and the CLI twin `tools/build_index.py`.

The parquet has the reference's corpus schema (`chunk_id`, `text`, `doc_id`:
/root/reference/tests/conftest.py:203-219); the encoder is a stub with the interface the reference's
StudentModel offers (`encode_documents(list[str]) -> [n, 384] float32`, unit rows), deterministic in the text.
Checked: ntotal, the `doc_ids[idx]` mapping the /search handler relies on, `texts.json`, `max_docs` and
`batch_size` handling, error behaviour, and search parity against the oracle after a save()/load() round trip
under the default metric ("cosine": what /index/load constructs, src/serve/app.py:427).
"""
import json
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu
TIE_TOL = 1e-3


class StubStudent:
    """encode_documents like the reference's StudentModel: unit-norm fp32 rows, a function of the text only."""
    embedding_dim = 384

    def __init__(self):
        self.calls = []

    def encode_documents(self, texts, **kw):
        self.calls.append(len(texts))
        out = np.empty((len(texts), 384), dtype=np.float32)
        for i, t in enumerate(texts):
            rng = np.random.default_rng(zlib.crc32(t.encode()))
            v = rng.standard_normal(384).astype(np.float32)
            out[i] = v / np.linalg.norm(v)
        return out


def _corpus(tmp_path, n=257):
    import pandas as pd
    texts = [f"Document {i}: passage about topic {i % 17} and entity {i * 7919 % 1000}" for i in range(n)]
    df = pd.DataFrame({"chunk_id": [f"chunk_{i}" for i in range(n)], "text": texts,
                       "doc_id": [f"doc_{i}" for i in range(n)]})
    p = tmp_path / "corpus.parquet"
    df.to_parquet(p)
    return p, texts


def test_build_from_parquet_save_load_search(tmp_path, oracle):
    import semantic_search_kd_b200 as pkg
    p, texts = _corpus(tmp_path)
    model = StubStudent()
    builder = pkg.FAISSIndexBuilder(embedding_dim=384, index_type="HNSW", metric="cosine")
    index = builder.build_from_parquet(model=model, parquet_path=p, batch_size=32, max_docs=200,
                                       hnsw_m=32, hnsw_ef_construction=200)
    assert index.ntotal == 200 and builder.ntotal == 200
    assert model.calls == [32] * 6 + [8]                       # batch_size respected, max_docs truncates
    assert builder.doc_ids == [f"doc_{i}" for i in range(200)]
    out = tmp_path / "index"
    builder.save(out)
    assert (out / "index.faiss").exists() and (out / "doc_ids.json").exists() and (out / "texts.json").exists()
    assert json.loads((out / "doc_ids.json").read_text()) == builder.doc_ids
    saved_texts = json.loads((out / "texts.json").read_text())
    assert saved_texts["doc_7"] == texts[7] and len(saved_texts) == 200

    # /index/load constructs FAISSIndexBuilder(embedding_dim=...) with the default metric and calls load()
    served = pkg.FAISSIndexBuilder(embedding_dim=384)
    served.load(out)
    assert served.ntotal == 200 and served.doc_ids == builder.doc_ids
    X = StubStudent().encode_documents(texts[:200])
    Q = StubStudent().encode_documents([texts[3], texts[150], "an unseen query about topic 5"])
    for idx in (builder, served):
        D, I = idx.search(Q, 10)
        Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
        rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL)
        assert rep["ok"], rep
        assert I[0, 0] == 3 and I[1, 0] == 150 and idx.doc_ids[int(I[1, 0])] == "doc_150"   # the handler's mapping
    Da, Ia = builder.search(Q, 10)
    Db, Ib = served.search(Q, 10)
    assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db)   # rows came back verbatim: bit-equal scores
    builder.close()
    served.close()


def test_cosine_save_load_cycles_are_bit_stable(tmp_path):
    """under metric="cosine" load() must not normalise (and re-round) rows a second time."""
    import semantic_search_kd_b200 as pkg
    rng = np.random.default_rng(5)
    X = (rng.standard_normal((3000, 384)) * rng.uniform(0.2, 5.0, (3000, 1))).astype(np.float32)   # NOT unit rows
    Q = rng.standard_normal((4, 384)).astype(np.float32)
    a = pkg.FlatIPIndex(384, metric="cosine")
    a.add(X)
    D0, I0 = a.search(Q, 20)
    d = tmp_path / "gen0"
    a.save(d)
    for gen in range(1, 4):                                     # three save / load generations
        b = pkg.FlatIPIndex(384)                                # default metric = cosine
        b.load(d)
        D, I = b.search(Q, 20)
        assert np.array_equal(I, I0) and np.array_equal(D, D0), gen
        d = tmp_path / f"gen{gen}"
        b.save(d)
        b.close()
    # the faiss file alone (no side-car) gives the same ids; scores equal up to the fp32 -> bf16 re-rounding of the file
    for f in ("rows.bf16", "b200_meta.json"):
        (d / f).unlink()
    c = pkg.FlatIPIndex(384)
    c.load(d)
    D, I = c.search(Q, 20)
    assert np.array_equal(I, I0) and np.array_equal(D, D0)      # bf16 -> fp32 -> bf16 is exact
    c.close()
    # an inner_product index must not be loaded as cosine silently, and vice versa
    ip = pkg.FlatIPIndex(384, metric="inner_product")
    ip.add(X)
    ip.save(tmp_path / "ip")
    with pytest.raises(pkg.IndexBuildError):
        pkg.FlatIPIndex(384, metric="cosine").load(tmp_path / "ip")
    back = pkg.FlatIPIndex(384, metric="inner_product")
    back.load(tmp_path / "ip")
    Dip, Iip = ip.search(Q, 20)
    Db, Ib = back.search(Q, 20)
    assert np.array_equal(Iip, Ib) and np.array_equal(Dip, Db)
    a.close(), ip.close(), back.close()


def test_stale_sidecar_does_not_shadow_a_newer_faiss_file(tmp_path):
    import os
    import time
    import semantic_search_kd_b200 as pkg
    from semantic_search_kd_b200 import faiss_io
    rng = np.random.default_rng(6)
    X1 = rng.standard_normal((500, 384)).astype(np.float32)
    X1 /= np.linalg.norm(X1, axis=1, keepdims=True)
    X2 = rng.standard_normal((500, 384)).astype(np.float32)
    X2 /= np.linalg.norm(X2, axis=1, keepdims=True)
    a = pkg.FlatIPIndex(384)
    a.add(X1)
    d = tmp_path / "idx"
    a.save(d)
    a.close()
    # someone replaces index.faiss (same shape) later, e.g. the reference's own builder
    faiss_io.write_flat_ip(d / "index.faiss", [X2], 500, 384)
    future = time.time() + 60
    os.utime(d / "index.faiss", (future, future))
    b = pkg.FlatIPIndex(384)
    b.load(d)
    D, I = b.search(X2[:3], 1)
    assert I[:, 0].tolist() == [0, 1, 2] and np.all(D[:, 0] > 0.99)
    b.close()


def test_build_errors(tmp_path):
    import pandas as pd
    import semantic_search_kd_b200 as pkg
    b = pkg.FAISSIndexBuilder(embedding_dim=384)
    with pytest.raises(pkg.IndexNotFoundError):
        b.build_from_parquet(StubStudent(), tmp_path / "missing.parquet")
    pd.DataFrame({"body": ["x"]}).to_parquet(tmp_path / "bad.parquet")
    with pytest.raises(pkg.IndexBuildError):
        b.build_from_parquet(StubStudent(), tmp_path / "bad.parquet")

    class Broken(StubStudent):
        def encode_documents(self, texts, **kw):
            if len(self.calls) == 2:
                raise RuntimeError("encoder out of memory")
            return super().encode_documents(texts, **kw)

    p, _ = _corpus(tmp_path, 100)
    with pytest.raises(pkg.IndexBuildError) as ei:
        b.build_from_parquet(Broken(), p, batch_size=16)
    assert "32 documents" in str(ei.value)
    b.close()


def test_cli_embeddings_build(tmp_path, oracle):
    """tools/build_index.py --embeddings: the north star's "build from an embedding array", end to end in a subprocess."""
    import semantic_search_kd_b200 as pkg
    rng = np.random.default_rng(9)
    X = rng.standard_normal((4096, 384)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    np.save(tmp_path / "emb.npy", X)
    ids = [f"p{i}" for i in range(4096)]
    (tmp_path / "ids.json").write_text(json.dumps(ids))
    out = tmp_path / "artifacts"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "build_index.py"), "--embeddings", str(tmp_path / "emb.npy"),
                        "--doc-ids", str(tmp_path / "ids.json"), "--output-dir", str(out), "--max-docs", "4000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Total vectors: 4000" in r.stdout
    idx = pkg.FAISSIndexBuilder(embedding_dim=384)
    idx.load(out)
    assert idx.ntotal == 4000 and idx.doc_ids == ids[:4000]
    Q = X[[5, 3999]]
    D, I = idx.search(Q, 5)
    Dr, Ir = oracle.flat_ip_topk(X[:4000], Q, 5)
    rep = oracle.compare_topk(D, I, Dr, Ir, X[:4000], Q, tie_tol=TIE_TOL)
    assert rep["ok"] and I[:, 0].tolist() == [5, 3999]
    idx.close()


def test_load_reference_hnsw_directory(tmp_path, oracle):
    """The reference builds `FAISSIndexBuilder(index_type="HNSW")` and saves with faiss.write_index: index.faiss is an
    IndexHNSWFlat file.  Our load() takes the stored vectors out of it (the graph is of no use to an exact index):
    a directory written by the reference's own build script is servable as it is."""
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    from test_host_logic import _write_hnsw_flat
    import semantic_search_kd_b200 as pkg
    rng = np.random.default_rng(21)
    X = rng.standard_normal((2000, 384)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    d = tmp_path / "ref_index"
    d.mkdir()
    _write_hnsw_flat(d / "index.faiss", X)
    (d / "doc_ids.json").write_text(json.dumps([f"doc_{i}" for i in range(len(X))]))
    idx = pkg.FAISSIndexBuilder(embedding_dim=384)            # what /index/load constructs
    idx.load(d)
    assert idx.ntotal == 2000 and idx.doc_ids[7] == "doc_7"
    Q = X[[7, 1999]] + 0.01 * rng.standard_normal((2, 384)).astype(np.float32)
    D, I = idx.search(Q, 10)
    Dr, Ir = oracle.flat_ip_topk(X, Q, 10)
    rep = oracle.compare_topk(D, I, Dr, Ir, X, Q, tie_tol=TIE_TOL)
    assert rep["ok"] and I[:, 0].tolist() == [7, 1999], rep
    idx.close()
