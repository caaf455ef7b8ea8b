"""bench.py's CPU arm (`--impl reference`) runs without a GPU and prints the contract's JSON line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_contract_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 1e5
    assert d["config"]["workload"] == "c2" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_ours_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)


def test_non_rank0_reference_arm_is_silent(monkeypatch):
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
