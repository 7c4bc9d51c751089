"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and
the last committed native line (profiles/) carries every key the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.strip().splitlines()) == 1, "stdout must carry the JSON line and nothing else"
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "projections/s" and line["higher_is_better"] is True
    assert BASE_KEYS <= set(line)
    assert line["value"] > 0 and line["gpu_launches"] == 0
    # "reference" when oracle/_ref holds the reference modules (oracle/build_ref.py), else the oracle's port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_committed_native_line_has_contract_keys():
    path = os.path.join(ROOT, "profiles", "r02_bench_cfg3.json")
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(line)
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
    assert line["config"]["Q"] == 64 and line["config"]["N"] == 3000
    # round 2: the headline fraction is the STEP (one pass of algorithmic bytes / ms_per_step), kernels listed under it
    assert abs(roof["achieved"] - roof["algorithmic_bytes"] / (line["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * roof["achieved"]
    assert {"p50", "p95", "max"} <= set(line["step_ms"]) and line["parity_check"]["ok"] is True
    assert roof["kernels"] and all("avg_ms" in v for v in roof["kernels"].values())
