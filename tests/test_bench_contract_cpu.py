"""The bench lines kept under profiles/ carry every key of the measurement contract (the driver parses the same line from
`python bench.py` / `python bench.py --impl reference`)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [x for x in f.read().strip().splitlines() if x.startswith("{")]
    assert len(lines) == 1, "bench.py prints exactly one JSON line"
    return json.loads(lines[0])


def test_bench_line_has_the_contract_keys():
    d = _line("r02_bench_n1_final3.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["steps"] * d["ms_per_step"] >= 2000.0            # >= 2 s timed region
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    # studies/s follows from the step time and the per-GPU batch
    assert abs(d["value"] - 256 / d["ms_per_step"] * 1e3) / d["value"] < 1e-6


def test_reference_arm_line_has_the_contract_keys():
    d = _line("r02_reference_arm_final.json")
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "studies/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_multi_gpu_lines_verify_the_gather():
    for n in (2, 4, 8):
        d = _line(f"r02_bench_n{n}_final.json")
        assert d["n_gpus"] == n and d["gather_check"] is True and d["strong"]["global_batch"] == 256
        assert abs(d["value"] - n * 256 / d["ms_per_step"] * 1e3) / d["value"] < 1e-6
