"""bench.py's reference arm runs without a GPU (it times the reference's CPU path): check the JSON
line it prints against the driver's contract on the smallest workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "small_96x128_L16_trws_linear", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fusion_move_sweeps_per_sec" and d["unit"] == "sweeps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["config"]["workload"].startswith("small_")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a device the product arm must fail loudly, not print a number."""
    import stereo_b200._lib as L
    if L.lib().sb_device_count() > 0:
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small_96x128_L16_trws_linear",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
