"""bench.py's reference arm runs on the CPU: check the JSON contract of the line it prints (one line
on stdout, the keys the driver reads) on a tiny workload.  The GPU arm's line has the same keys plus
`roofline`, `clocks`, `gpu_launches` (checked on the GPU box by the driver itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--mesh", "24", "--states", "8", "--ref-sample-states", "4"],
                         capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    return out.stdout


def test_reference_arm_prints_one_json_line():
    lines = [l for l in _run().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("band-FFTs/s") and d["unit"] == "band-FFTs/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 3 * 4 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "4 of 8 states" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
