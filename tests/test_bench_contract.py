"""bench.py's reference arm (CPU, the oracle port) honours the JSON contract the driver parses."""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("tokens/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["config"]["workload"].startswith("Qwen3-8B bf16 seq_len 4096")
    # the timed step is the bounded sample itself, really executed: value and ms_per_step describe the same measurement
    assert abs(d["value"] - d["config"]["sample_tokens"] / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-2
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank0_only():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, cwd=REPO, env=env)
    assert res.returncode == 0 and not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
