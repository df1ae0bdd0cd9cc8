"""The committed bench lines (profiles/bench_r1_*.json, produced by bench.py on a B200) carry every key the
measurement contract names; a cheap guard against dropping one while editing bench.py."""
import json
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_our_arm_line_has_the_contract_keys():
    l = _line("bench_r1_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in l, k
    assert l["warmup"] >= 3 and l["n_gpus"] == 1 and l["higher_is_better"] is True and l["vs_baseline"] is None
    r = l["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is not None and r["traffic"] > r["algorithmic_bytes_per_launch"] * 0.9
    e = l["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < l["value"]
    c = l["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["unit"] == l["unit"]
    assert l["gpu_launches"] == 2 * 5 * l["steps"]
    assert not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "workload" in l["config"]


def test_reference_arm_line():
    l = _line("bench_r1_reference.json")
    assert l["impl"] == "reference" and l["cpu_baseline"]["kind"] == "reference"
    assert l["e2e"]["h2d_bytes_per_step"] == 0 and l["e2e"]["d2h_bytes_per_step"] == 0 and l["e2e"]["value"] == l["value"]
    ours = _line("bench_r1_final.json")
    assert (l["metric"], l["unit"], l["higher_is_better"]) == (ours["metric"], ours["unit"], ours["higher_is_better"])
