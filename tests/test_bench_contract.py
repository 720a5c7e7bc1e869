"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (`--impl reference`) prints ONE JSON
line with the agreed keys, a non-zero rank of a torchrun launch prints nothing, and the GPU arm refuses to run without a
device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "3", "--no-ref-full-step"])    # (the full-K step takes ~40 s here)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rollouts/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("MPPI rollouts/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "K=65536 T=64" in d["config"]["workload"]
    assert d["config"]["K"] == 65536 and d["config"]["T"] == 64 and d["config"]["sample_K"] <= 65536
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    from motion_planning_b200 import _capi
    if _capi.load().mppi_device_count() > 0:
        return      # on a GPU box the arm runs for real (covered by the driver)
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in (r.stderr + r.stdout)
