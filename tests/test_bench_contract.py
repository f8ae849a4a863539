"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
keys, runnable without a GPU. Uses the small C1 config so it takes seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line(tmp_path):
    env = dict(os.environ, PGCN_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["unit"] == "edges/s" and r["higher_is_better"] is True
    assert r["metric"].startswith("aggregated edges/sec") and r["steps"] == 3 and r["warmup"] == 1
    assert r["value"] > 0 and abs(r["value"] - (10556 + 2708) / (r["ms_per_step"] * 1e-3)) / r["value"] < 1e-6
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["vs_baseline"] is None and r["dtype"] == "f32" and r["data"] == "synthetic" and "workload" in r["config"]


def test_reference_arm_other_ranks_exit_quietly(tmp_path):
    env = dict(os.environ, PGCN_CACHE=str(tmp_path), RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1",
                          "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
