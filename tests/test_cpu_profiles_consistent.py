"""CPU tier: the committed bench evidence (profiles/r02_bench_*.json, written by bench.py on the GPU box) carries the keys the
contract asks for and its derived figures follow from its own raw ones (a hand-edited or stale file fails here)."""
import json
import os

import pytest

from util import ROOT

FILES = ["r02_bench_c2.json", "r02_bench_c3_hifigan_b8.json", "r02_bench_c4_t2000.json", "r02_bench_c2_2gpu.json"]


def _load(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read()
    return json.loads([l for l in txt.splitlines() if l.startswith("{")][0])


@pytest.mark.parametrize("name", FILES)
def test_bench_line_is_self_consistent(name):
    d = _load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["warmup"] >= 3
    cfg = d["config"]
    samples = cfg["global_batch"] * cfg["samples_per_utt"]
    assert abs(d["value"] - samples / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]          # whole-job samples / max-over-ranks time
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 4 * samples // d["n_gpus"] and 0.5 * d["value"] < e["value"] < 1.05 * d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} and c["sm_mhz"] > 0.7 * c["sm_max_mhz"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and abs(r["frac_executed"] - r["executed_tflops"] / r["peak"]) < 1e-9
    assert r["frac_executed"] <= 2.0 * r["frac"] * 1.51 and r["frac"] < 0.5            # 2 (3 in the F0/N predictor) MMAs per algorithmic product
    assert r["traffic"] is None or 0.9 < r["traffic"] / 3019948032.0 < 1.2            # ncu DRAM bytes of the K=11 C128 launch vs algorithmic
    assert abs(r["achieved"] - r["algorithmic_gflop_per_launch"] * r["launches_per_step"] / r["kernel_ms_per_step_eager_events"]) < 1e-6 * r["achieved"] * 1e3


def test_single_gpu_line_has_the_cpu_baseline():
    d = _load("r02_bench_c2.json")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "samples/s" and "utterance" in cb["sample"]
    assert d["value"] / cb["value"] > 100        # a reported baseline, not the target
