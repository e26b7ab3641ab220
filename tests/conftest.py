"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu`` and run on a B200 via gpurun;
everything else runs on CPU only.  Nothing here reads /root/reference."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def unit_rows(n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, d)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    return a.astype(np.float32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def golden_cases():
    """name -> (X, Q, npz) for the committed golden vectors (inputs regenerated + hash-checked)."""
    import hashlib
    import json
    sys.path.insert(0, str(GOLDEN))
    import make_golden as mg
    out = {}
    z = np.load(GOLDEN / "conftest_fixture.npz")
    out["conftest_fixture"] = (z["X"], z["Q"], z)
    manifest = json.loads((GOLDEN / "manifest.json").read_text())
    for name, m in manifest.items():
        X, Q = mg.seeded_case(m["n"], m["nq"], m["seed"], m["dup"])
        assert hashlib.sha256(X.tobytes()).hexdigest() == m["sha256_X"], f"{name}: corpus bytes differ"
        assert hashlib.sha256(Q.tobytes()).hexdigest() == m["sha256_Q"], f"{name}: query bytes differ"
        out[name] = (X, Q, np.load(GOLDEN / f"{name}.npz"))
    return out
