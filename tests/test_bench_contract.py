"""bench.py's JSON contract, checked without a GPU: the product arm runs with host stand-ins for the device-facing pieces
(tests/_mock_bench.py), so its control flow, argument handling and the keys the driver reads are exercised on every CPU run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_mock_bench.py"), "--steps", "4", "--warmup", "3", "--no-profile",
                        "--no-cpu-baseline", "--sustained-seconds", "0.05", *(flags if "--extras" in flags else ("--no-extras", *flags))], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout                                   # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def test_product_arm_line_has_the_contract_keys():
    d = _line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "bf16" and d["data"] == "synthetic" and d["unit"] == "images/s"
    assert "EdgeLine-YOLO-n" in d["metric"] and "640x640" in d["metric"] and "workload" in d["config"] and "model" not in d["config"]
    assert "predict defaults" in d["config"]["workload"] and "l2_policy" in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 64 * 640 * 640 * 3 and e["d2h_bytes_per_step"] == 64 * 300 * 6 * 4 + 64 * 4
    assert e["value"] > 0 and e["ms_per_step"] > 0 and e["wall_ms_per_step"] > 0
    assert d["gpu_launches"] == 7 * 4 and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_product_arm_names_a_non_default_configuration():
    d = _line("--scale", "s", "--imgsz", "1280", "--batch", "32", "--nc", "10", "--conf", "0.001", "--multi-label")
    assert "EdgeLine-YOLO-s" in d["metric"] and "1280x1280" in d["metric"]
    w = d["config"]["workload"]
    assert "conf 0.001" in w and "multi_label" in w and "33600 anchors" in w and d["config"]["batch_per_gpu"] == 32


def test_product_arm_extra_inference_configs_ride_in_the_same_line():
    d = _line("--extras", "c2,c3")
    x = d["extra_configs"]
    assert set(x) == {"configs[2]", "configs[3]"}
    assert "1280x1280" in x["configs[2]"]["metric"] and "EdgeLine-YOLO-s" in x["configs[2]"]["metric"] and x["configs[2]"]["scaling"] == "weak"
    assert x["configs[3]"]["scaling"] == "strong" and x["configs[3]"]["global_batch"] == 512 and x["configs[3]"]["batch_per_gpu"] == 512
    assert x["configs[3]"]["e2e"]["h2d_bytes_per_step"] == 512 * 640 * 640 * 3
    assert d["sustained"]["seconds"] >= 0.05 and d["sustained"]["steps"] >= 4 and d["sustained"]["value"] > 0
    assert d["e2e"]["h2d_gbs_per_rank"] > 0
