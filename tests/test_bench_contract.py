"""bench.py's reference arm runs without a GPU: its JSON line must carry the keys the driver reads, and under a
multi-rank launch only rank 0 may print."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "600", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["unit"] == "eigenpairs/s" and b["higher_is_better"] is True
    assert b["metric"].startswith("SymArrow n=600") and b["n_gpus"] == 1 and b["steps"] == 2 and b["warmup"] == 1
    assert b["value"] > 0 and b["ms_per_step"] > 0 and b["dtype"] == "f64" and b["data"] == "synthetic"
    assert b["cpu_baseline"]["kind"] == "port" and b["cpu_baseline"]["cores"] >= 1 and b["cpu_baseline"]["value"] == b["value"]
    assert b["e2e"] == {"value": b["value"], "unit": "eigenpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["vs_baseline"] is None and "workload" in b["config"]


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
