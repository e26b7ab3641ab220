#!/usr/bin/env python3
"""tools/make_faiss_golden.py -- run on ANY machine where `import faiss` works (faiss-cpu 1.7.4 is what the
reference pins: pyproject.toml:15) to turn the faiss half of the parity claim into committed fixtures:

    tests/golden/faiss_<case>.npz : X, Q, and faiss.IndexFlatIP(384).search(Q, k) -> D_k*, I_k* for k in 1, 10, 100
    tests/golden/faiss_index_10x384.faiss : faiss.write_index of the reference's conftest fixture

tests/test_faiss_compat.py compares live against faiss when it is importable; with these files present the
CPU suite can check the oracle against faiss' recorded answers even where faiss is absent (the loader is in
tests/test_oracle.py::test_faiss_recorded_answers, which skips while the files do not exist).  This image has
no faiss wheel, so the files are not in the repository yet -- stated in DESIGN.md ("parity unpinned: faiss half")."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))


def main():
    try:
        import faiss
    except ImportError:
        raise SystemExit("faiss is not importable here: nothing written")
    from test_faiss_compat import _cases, _fixture_rows
    out = ROOT / "tests" / "golden"
    for name, (X, Q) in _cases().items():
        index = faiss.IndexFlatIP(384)
        index.add(X)
        rec = {"X": X, "Q": Q, "faiss_version": np.array(faiss.__version__)}
        for k in (1, 10, 100):
            D, I = index.search(Q, k)
            rec[f"D_k{k}"], rec[f"I_k{k}"] = D, I
        np.savez_compressed(out / f"faiss_{name}.npz", **rec)
        print("wrote", out / f"faiss_{name}.npz")
    index = faiss.IndexFlatIP(384)
    index.add(_fixture_rows())
    faiss.write_index(index, str(out / "faiss_index_10x384.faiss"))
    print("wrote", out / "faiss_index_10x384.faiss")


if __name__ == "__main__":
    main()
