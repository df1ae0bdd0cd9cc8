"""The committed bench lines (profiles/bench_r2_*.json, produced by bench.py on a B200) carry every key the
measurement contract names; a cheap guard against dropping one while editing bench.py."""
import json
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_our_arm_line_has_the_contract_keys():
    l = _line("bench_r2_final.json")
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
    assert e["steps"] >= 10 and e["copy_only_value"] > 0            # e2e over >= 10 steps, with its roof
    assert l["config"]["distinct_images_per_gpu"] == l["config"]["batch_per_gpu"]
    c3 = l["configs3"]                                                # BASELINE configs[3], measured in every run
    assert c3["scaling"] == "strong" and c3["stage_mpix_s"] > 0
    assert c3["files_device_entropy"]["file_equals_host_entropy_path"] is True
    assert c3["files_host_entropy"]["round_trip_equals_decode_stage"] is True
    assert l["class_api"]["transform_quant_mpix_s"] > 0
    assert r["thread_inst_per_sample"] > 0


def test_scaling_lines_share_the_config():
    base = _line("scale_r2_n1.json")
    for n in (2, 4, 8):
        l = _line("scale_r2_n%d.json" % n)
        assert l["n_gpus"] == n and l["config"] == base["config"] and l["scaling"] == "weak"
        assert l["value"] > 0.9 * n * base["value"]                   # independent images per rank: no collective
        assert l["configs3"]["images_per_gpu"] * n >= 4096


def test_reference_arm_line():
    l = _line("bench_r2_reference.json")
    assert l["impl"] == "reference" and l["cpu_baseline"]["kind"] == "reference"
    assert l["e2e"]["h2d_bytes_per_step"] == 0 and l["e2e"]["d2h_bytes_per_step"] == 0 and l["e2e"]["value"] == l["value"]
    ours = _line("bench_r2_final.json")
    assert (l["metric"], l["unit"], l["higher_is_better"]) == (ours["metric"], ours["unit"], ours["higher_is_better"])
    assert l["config"] == ours["config"]                              # identical config object in both arms
