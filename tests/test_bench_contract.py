"""bench.py's reference arm (`--impl reference`: the oracle port timed on host cores) produces
the JSON line the driver parses.  Runs on CPU with a tiny bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--dim", "40", "--ref-seconds", "0.2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                  # exactly one JSON line on stdout
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["metric"] == "leapfrog_steps_per_sec" and r["unit"] == "leapfrog-steps/s"
    assert r["higher_is_better"] is True and r["n_gpus"] == 1 and r["steps"] == 2 and r["warmup"] == 1
    assert r["dtype"] == "f64" and r["data"] == "synthetic" and r["vs_baseline"] is None
    assert r["value"] > 0 and r["ms_per_step"] > 0 and "workload" in r["config"]
    cb = r["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == r["value"] and cb["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["gpu_launches"] == 0                           # nothing of the CUDA path runs in this arm


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--dim", "40", "--ref-seconds", "0.1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
