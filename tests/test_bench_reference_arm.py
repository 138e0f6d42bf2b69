"""`bench.py --impl reference` (the driver's reference arm) on this machine's CPU: it must run the UNMODIFIED
reference functions staged in oracle/_ref when they are available (kind "reference"), on the b200 arm's workload
object, and print one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_contract():
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    from oracle import stage_ref
    stage_ref.stage()
    assert line["config"] == bench.workload_config(1)                      # the same workload object as the b200 arm
    assert line["cpu_baseline"]["kind"] == ("reference" if stage_ref.available() else "port")
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["sample_images_per_step"] == bench.WORKLOAD["n"]           # the full 16-image batch
