"""bench.py's reference arm (`--impl reference`): the JSON line the driver parses, produced on the CPU (the arm times the
oracle port of the reference's path on the host cores), and the multi-rank rule that only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "4k_10bit_frames_per_sec_full_cascade" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) <= 1e-6 * d["ms_per_step"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "frames/s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
