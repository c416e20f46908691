"""bench.py prints ONE JSON line with the keys the driver's contract names — checked for the
reference arm on CPU and for the B200 arm on a GPU (small workload, few steps)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
             "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
             "gpu_launches"}


def run_bench(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line: %r" % lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["flow_value"] == 100  # config 1: F* = M


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--workload", "c1", "--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    e = d["e2e"]
    # the headline end-to-end figure starts from the C ABI's 32-bit host columns and does whatever
    # narrowing it uses inside the timed region; the other encodings are reported next to it
    assert e["encode_in_timed_region"] is True and "uint32 start/end" in e["host_input"]
    assert e["h2d_bytes_per_step"] in (2 * 1_000_000, 8 * 1_000_000)
    assert e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    others = {k for k in d if k.startswith("e2e_")}
    assert "e2e_preencoded" in others and d["e2e_preencoded"]["encode_in_timed_region"] is False
    assert d["e2e_preencoded"]["h2d_bytes_per_step"] == 2 * 1_000_000
    assert ("e2e_u32" in others) != (e["transport"] == "start u32, end u32")
    assert d["result"]["quality"]["kept_over_lower_bound"] < 1.06
    assert d["result"]["bundle_path"] == "histogram in shared memory"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["result"]["fstar"] == d["result"]["flow_value"] == 100
    assert d["config"]["workload"].startswith("c1")
    # both arms describe the workload with the same dict (the driver compares them)
    ref = run_bench("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0")
    assert ref["config"] == d["config"] and ref["cpu_baseline"]["cores"] >= 1
