"""CPU: the bench.py output contract -- the committed B200 lines under profiles/ carry every key the driver
and the judge read, and the reference arm (`--impl reference`: the oracle port on the host cores) runs here."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config")


def _lines():
    out = []
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_*gpu*.json"))):
        with open(path) as f:
            out.append((os.path.basename(path), json.loads(f.read().strip().splitlines()[-1])))
    return out


def test_profiles_hold_bench_lines():
    names = [n for n, _ in _lines()]
    assert any("c3_1gpu" in n for n in names) and any("8gpu" in n for n in names)


@pytest.mark.parametrize("name,line", _lines())
def test_committed_bench_line_matches_contract(name, line):
    for k in BASE_KEYS:
        assert k in line, k
    assert line["unit"] == "edge-updates/s" and line["higher_is_better"] is True and line["data"] == "synthetic"
    assert line["dtype"] == "f32" and line["scaling"] in ("weak", "strong") and line["warmup"] >= 3
    assert "workload" in line["config"] and "model" not in line["config"]
    # value = E / time per step, whole job
    assert line["value"] == pytest.approx(line["config"]["E"] / (line["ms_per_step"] * 1e-3), rel=1e-9)
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] < line["value"]                         # host copies are inside its timed region
    assert line["gpu_launches"] >= (8 if name.startswith("r01") else 6) * line["steps"]      # round 2: fused preparation
    clk = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clk)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clk["reasons"])
    if line["n_gpus"] == 1:
        roof = line["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
        assert roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"], rel=1e-6)
        if line.get("cpu_baseline") is not None:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
            assert line["cpu_baseline"]["kind"] in ("port", "reference")


def test_timing_rule_is_a_function_of_the_workload():
    """`config.l2` names the timing rule (L2 flushed before every step, or inputs larger than L2 and K iterations back
    to back) from the workload alone, so both arms print the same `config`; a line timed back to back also carries
    the flushed single-iteration figure."""
    sys.path.insert(0, ROOT)
    import bench
    c3, c2 = bench.WORKLOADS["c3"], bench.WORKLOADS["c2"]
    assert bench.iteration_bytes(c3, 1_000_000, 3_999_789) == pytest.approx(212e6, rel=0.01)
    assert bench.back_to_back(c3, 1_000_000, 3_999_789) and not bench.back_to_back(c2, 100_000, 399_984)
    assert "NOT flushed" in bench.workload_config(c3, 1_000_000, 3_999_789)["l2"]
    assert "flushed before every timed step" in bench.workload_config(c2, 100_000, 399_984)["l2"]
    for name, line in _lines():
        if "NOT flushed" in line["config"].get("l2", ""):
            fl = line["flushed_single_replays"]
            assert fl is not None and fl["ms_per_step"] > 0, name
            assert line["ms_per_step"] <= 1.05 * fl["ms_per_step"], name      # launch latency of one graph per iteration


def test_reference_arm_runs_on_host_cores():
    """`bench.py --impl reference`: one JSON line, impl = reference, e2e repeats the line's own value with zero
    transfer bytes, cpu_baseline describes the run."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for k in BASE_KEYS:
        assert k in line, k
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference")
