"""The bench line's contract (task statement, section 4 of the tier framing) checked on the committed records of the round:
every key the driver reads is there, with the meaning the contract gives it."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _line(name):
    return json.loads(open(os.path.join(PROF, name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n_gpus", [("r02_bench_default_1m.json", 1), ("r02_bench_2gpu_1m.json", 2), ("r02_bench_8gpu_1m.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n_gpus):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == n_gpus and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "synthetic" in d["data"]
    assert "LM it" in d["unit"] and "workload" in d["config"] and "model" not in d["config"]
    assert "LM iterations/s" in d["metric"] and "1M" in d["metric"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] * 1.001
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.5 < r["frac"] < 1.0
    if n_gpus == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and len(cb["sample"]) > 20
        assert r["traffic"] is None or r["traffic"] >= 0.9 * r["algorithmic_bytes"]
        for sub in ("fp32_mode", "reference_api_e2e", "classic_ba", "config5"):
            assert isinstance(d.get(sub), dict) and "error" not in d[sub], sub
    else:
        assert d["point_sharded"]["lm_it_per_s"] > d["value"] / n_gpus          # the same pair is faster on N GPUs than on one
        assert d["config5"]["problems"] == 64 * n_gpus


def test_committed_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == pytest.approx(d["value"])
    assert d["e2e"]["value"] == pytest.approx(d["value"]) and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["correspondences"] == 1_000_000 and d["config"]["lm_iters_timed"] == 30       # like for like: the same pair, all 30 iterations
