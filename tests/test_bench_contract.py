"""CPU: the reference arm of bench.py honours the driver's JSON contract (the CUDA arm needs a B200
and is exercised by the driver itself)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))


def test_reference_arm_line():
    r = _run(None, "--steps", "2", "--warmup", "1", "--no-hnsw")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "queries/s"
    assert d["metric"].startswith("queries/sec exact top-10") and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and d["config"]["rows"] == 8841823
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
