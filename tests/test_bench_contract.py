"""bench.py's reference arm runs on the CPU alone: its JSON line is checked here against the
contract (one line on stdout; metric / unit / config of the GPU arm; `impl`, `cpu_baseline`,
`e2e` with zero copy bytes), on a small raster so that the CPU suite stays short."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0")
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "2048",
         "--steps", "2", "--warmup", "1", "--legs", "chain"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["unit"] == "Gpixel/s" and line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["ms_per_step"] > 0 and abs(line["value"] - 2048 * 2048 / line["ms_per_step"] / 1e6) < 1e-6 * line["value"] + 1e-9
    assert line["config"]["same_config"] is True and "cfg2" in line["config"]["workload"]
    base = line["cpu_baseline"]
    assert base["kind"] in ("stub harness", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "2048"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""
