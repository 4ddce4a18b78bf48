"""CPU: `bench.py --impl reference` (the reference's own CPU path on this box's cores) prints the one JSON line the driver reads,
with the same metric / unit / config as the GPU arm.  No GPU, no oracle import outside the allowed leg."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    rec = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert rec["impl"] == "reference" and rec["metric"] == "pages/sec (resize+PNG+base64)" and rec["unit"] == "pages/s"
    assert rec["higher_is_better"] is True and rec["n_gpus"] == 1 and rec["steps"] == 1 and rec["value"] > 0
    assert rec["config"]["workload"] == bench.WORKLOAD and rec["config"]["pages_per_gpu"] == bench.PAGES_PER_GPU
    cb = rec["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == rec["value"] and "64 pages" in cb["sample"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(rec["ms_per_step"] * rec["value"] / 1e3 - bench.PAGES_PER_GPU) < 1e-6 * bench.PAGES_PER_GPU + 1e-3
