"""The faiss pin that switches itself on.

faiss-cpu (the reference pins faiss-cpu 1.7.4: pyproject.toml:15) is NOT in this image, so the faiss half of
the oracle is restated from faiss' published sources and marked "unpinned" (oracle/flat_ip.c header,
DESIGN.md).  Everything below is skipped here and runs wherever `import faiss` works:

  CPU (no GPU needed): the oracle's IndexFlatIP restatement == faiss.IndexFlatIP.search on the reference's
      conftest fixture (tests/conftest.py:66-73) and on the duplicate-heavy golden case: ids, tie order,
      -1 / -FLT_MAX padding for k > ntotal; our index.faiss writer is readable by faiss.read_index and
      byte-identical to faiss.write_index; our reader parses a faiss-written file.
  GPU: FlatIPIndex.load of a faiss-written directory (the reference's `temp_index_dir` fixture,
      tests/conftest.py:180-200) and search parity against faiss itself.

tools/make_faiss_golden.py writes the same comparisons as fixtures (tests/golden/faiss_*.npz) on any
machine that has faiss, so that the pin can then travel to machines that have not.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

faiss = pytest.importorskip("faiss", reason="faiss is not installed in this image: the faiss half of the oracle stays unpinned")

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def _fixture_rows():
    np.random.seed(42)                      # /root/reference/tests/conftest.py:66-73
    e = np.random.randn(10, 384).astype(np.float32)
    return e / np.linalg.norm(e, axis=1, keepdims=True)


def _cases():
    from conftest import unit_rows
    X = unit_rows(3000, 384, 7)
    X[1000:2000] = X[:1000]                 # exact duplicates: every score tie is between ids i and i + 1000
    return {"conftest_fixture": (_fixture_rows(), unit_rows(4, 384, 8)), "dups3000": (X, unit_rows(16, 384, 9))}


@pytest.mark.parametrize("case", ["conftest_fixture", "dups3000"])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_oracle_equals_faiss_flat_ip(oracle, case, k):
    X, Q = _cases()[case]
    index = faiss.IndexFlatIP(384)
    index.add(X)
    Df, If = index.search(Q, k)
    Do, Io = oracle.flat_ip_topk(X, Q, k, acc="f32")
    assert np.array_equal(If, Io), "ids / tie order / -1 padding differ from faiss"
    pad = If < 0
    assert np.array_equal(pad, Io < 0)
    assert np.allclose(Df[~pad], Do[~pad], atol=2e-6)
    if pad.any():
        assert np.all(Df[pad] == Do[pad]), "padding score differs from faiss (-FLT_MAX expected)"


def test_index_file_is_byte_identical_and_readable(tmp_path):
    from semantic_search_kd_b200 import faiss_io
    X = _fixture_rows()
    index = faiss.IndexFlatIP(384)
    index.add(X)
    faiss.write_index(index, str(tmp_path / "theirs.faiss"))
    faiss_io.write_flat_ip(tmp_path / "ours.faiss", [X], len(X), 384)
    assert (tmp_path / "ours.faiss").read_bytes() == (tmp_path / "theirs.faiss").read_bytes()
    back = faiss.read_index(str(tmp_path / "ours.faiss"))
    assert back.ntotal == len(X) and back.d == 384 and back.metric_type == faiss.METRIC_INNER_PRODUCT
    rows, metric = faiss_io.read_flat(tmp_path / "theirs.faiss")
    assert metric == 0 and np.array_equal(np.asarray(rows), X)


@pytest.mark.gpu
def test_load_faiss_written_directory_and_match_faiss(tmp_path, oracle):
    """The reference's temp_index_dir fixture, read by our index; answers vs faiss under the 1e-3 rule."""
    import semantic_search_kd_b200 as pkg
    for name, (X, Q) in _cases().items():
        d = tmp_path / name
        d.mkdir()
        index = faiss.IndexFlatIP(384)
        index.add(X)
        faiss.write_index(index, str(d / "index.faiss"))
        (d / "doc_ids.json").write_text(json.dumps([f"doc_{i}" for i in range(len(X))]))
        ours = pkg.FAISSIndexBuilder(embedding_dim=384)
        ours.load(d)
        assert ours.ntotal == len(X)
        for k in (1, 10, 100):
            Df, If = index.search(Q, k)
            D, I = ours.search(Q, k)
            rep = oracle.compare_topk(D, I, Df, If, X, Q, tie_tol=1e-3)
            assert rep["ok"], (name, k, rep)
            assert np.array_equal(I < 0, If < 0)
        ours.save(tmp_path / (name + "_resaved"))
        again = faiss.read_index(str(tmp_path / (name + "_resaved") / "index.faiss"))
        assert again.ntotal == len(X)
        ours.close()


def test_our_reader_takes_the_vectors_out_of_a_faiss_hnsw_file(tmp_path):
    """faiss.write_index(IndexHNSWFlat) -- the reference's default index_type -- read by faiss_io.read_flat."""
    from semantic_search_kd_b200 import faiss_io
    X = _cases()["dups3000"][0]
    index = faiss.IndexHNSWFlat(384, 32, faiss.METRIC_INNER_PRODUCT)
    index.hnsw.efConstruction = 40
    index.add(X)
    faiss.write_index(index, str(tmp_path / "index.faiss"))
    rows, metric = faiss_io.read_flat(tmp_path / "index.faiss")
    assert metric == 0 and np.array_equal(np.asarray(rows), X)
