"""The reference arm of bench.py (`--impl reference`, CPU only) prints exactly one JSON line on stdout with the keys
the driver reads; ranks other than 0 of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.update(extra_env or {})
    cmd = [sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1",
           "--cpu-sample-queries", "4", "--cpu-workers", "2"]
    return subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
