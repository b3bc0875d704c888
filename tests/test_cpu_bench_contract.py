"""bench.py's reference arm on the smoke-sized workload, no GPU: the JSON line carries the keys the measurement contract names, and its `config`
is the dict the GPU arm would print for the same workload (what makes the driver's ratio a same-configuration claim)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"], cwd=ROOT, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["dtype"] == "f64" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["unit"] == "residuals/s"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # the same `config` as the GPU arm of this workload
    sys.path.insert(0, ROOT)
    import bench
    c = line["config"]
    assert c == bench.job_config("tiny", c["n_cams"], c["n_points"], c["n_obs"], 1, "strong")
