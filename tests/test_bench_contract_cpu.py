"""The committed bench lines (profiles/r02_bench_*gpu_final*.json, produced by bench.py on the B200 boxes) carry every key of the
measurement contract -- a format check that needs no GPU, so a change to bench.py's JSON line that drops a key is caught here."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*gpu_final*.json")))

TOP = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
       "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "sections_ms"]


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_bench_line_has_the_contract_keys(path):
    d = json.load(open(path))
    for k in TOP:
        assert k in d, k
    assert d["metric"].startswith("SVGP-Gibbs ELBO steps/s") and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    assert abs(d["value"] - d["steps"] / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-6 * d["value"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:  # the CPU baseline and the headline-size parity run on rank 0 of the single-GPU line only
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in d["cpu_baseline"], k
        assert d["cpu_baseline"]["kind"] in ("port", "reference")
        assert d["parity"]["ok"] is True and d["parity"]["elbo_rel"] <= 1e-6 and d["parity"]["grad_rel_max"] <= 1e-6


def test_there_are_lines_for_1_2_4_8_gpus():
    n = sorted({json.load(open(f))["n_gpus"] for f in FILES})
    assert n == [1, 2, 4, 8], n
