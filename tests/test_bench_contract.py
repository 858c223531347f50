"""The bench lines committed under profiles/ (real B200 runs of `bench.py`) carry every key the driver's contract names."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parent.parent / "profiles"
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches"}


def _load(name):
    return json.loads((PROFILES / name).read_text().strip().splitlines()[-1])


def test_full_line_has_the_contract_keys():
    d = _load("r1_s2_bench_full.json")  # `python bench.py`, defaults
    assert BASE <= set(d) and {"clocks", "roofline", "cpu_baseline", "train_iter"} <= set(d)
    assert d["unit"] == "MPix/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - 1920 * 1080 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 < e["d2h_bytes_per_step"]
    assert e["value"] < d["value"]  # host copies inside the timed region cost something
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    t = d["train_iter"]
    assert "error" not in t and t["deform_roofline"]["bound"] == "tensor" and 0 < t["deform_roofline"]["frac"] < 1
    assert t["ms"] > t["ms_without_deform"] > 0


def test_reference_arm_line():
    d = _load("r1_s2_bench_reference.json")  # `python bench.py --impl reference`
    assert d["impl"] == "reference" and BASE <= set(d) and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"


@pytest.mark.parametrize("name", ["r1_s2_bench_2gpu.json"])
def test_multi_gpu_line(name):
    d = _load(name)
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and BASE <= set(d)
    assert "view-sharded" in d["config"]["parallelism"] and d["config"]["views_per_step"] == 2


# ---- round 2 lines -------------------------------------------------------------------------------------------------------
def test_round2_full_line():
    d = _load("r2_final_bench1.json")  # `python bench.py --steps 20 --warmup 5`
    assert BASE <= set(d) and {"clocks", "roofline", "cpu_baseline", "train_iter", "configs"} <= set(d)
    assert d["unit"] == "MPix/s" and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert r["kernel"] == "rasterize_bwd" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["frac"] > 0.55
    c = d["configs"]
    assert "error" not in c and {"cfg1", "cfg2", "knn"} <= set(c)
    k = c["knn"]
    assert k["points"] == 3_000_000 and k["k"] == 16 and k["bit_exact_distances_on_sample"] is True
    assert {"sklearn_1core", "sklearn_all_cores"} <= set(k["cpu"]) and k["roofline"]["bound"] == "hbm"
    t = d["train_iter"]
    assert "error" not in t and {"p10", "p50", "p90", "max"} <= set(t["ms_quantiles"])
    assert t["ms_quantiles"]["p90"] / t["ms_quantiles"]["p50"] < 1.1  # a reproducible number
    assert 0 < t["deform_roofline"]["frac_algorithmic"] < t["deform_roofline"]["frac"] < 1


def test_round2_reference_arm_is_like_for_like():
    d, r = _load("r2_final_bench1.json"), _load("r2_final_bench_reference.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    assert r["config"]["workload"] == d["config"]["workload"]          # same workload string on both arms
    assert r["steps"] == d["steps"] and r["warmup"] == d["warmup"]      # same step / warm-up counts
    assert "crop-extrapolated" in r["cpu_baseline"]["sample"] and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("name,n", [("r2_final_bench2.json", 2), ("r2_final_bench4.json", 4), ("r2_final_bench8.json", 8)])
def test_round2_multi_gpu_lines(name, n):
    d = _load(name)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and BASE <= set(d) and d["config"]["views_per_step"] == n
    e = d["exchange"]
    assert e["mode"].startswith("peer") and e["ms"] > 0 and e["allreduce_bytes"] == 68_000_000
    assert e["multicast"] is (n >= 4)
