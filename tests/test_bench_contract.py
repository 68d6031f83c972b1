"""bench.py's reference arm (CPU only, so it runs here): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def _have_ref():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "stable_baselines3", "__init__.py"))


def test_reference_arm_line():
    """kind "reference" (the unmodified reference from baseline/_ref) when that copy exists, the oracle port otherwise."""
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-steps", "12", "--ref-budget", "6"])
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] and d["unit"] and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("HalfCheetah")
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if _have_ref() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "K4" in cb["sample"]
    assert set(d["config"]) >= {"workload", "batch_size", "n_epochs", "rollouts", "n_steps", "parallelism"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-steps", "12"], env).strip() == ""
