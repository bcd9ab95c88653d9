"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys, and
the thread count of the CPU arm honours the container's CPU quota."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "3"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["unit"] == "props/s" and out["dtype"] == "f64" and out["warmup"] >= 3
    assert out["e2e"] == {"value": out["value"], "unit": "props/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == out["value"] and cb["cores"] == cb["host"]["threads"] >= 1
    assert "workload" in out["config"] and "model" not in out["config"]


def test_usable_cpus_never_exceeds_the_affinity_mask():
    sys.path.insert(0, ROOT)
    import bench
    info = bench.usable_cpus()
    assert 1 <= info["threads"] <= min(info["logical"], info["affinity"])
    if info["cgroup_cpu_max"] and not info["cgroup_cpu_max"].startswith("max"):
        quota, period = info["cgroup_cpu_max"].split()
        assert info["threads"] <= max(1, round(float(quota) / float(period)))
