"""The one JSON line bench.py prints is a contract with the driver.  These CPU tests check the shape of the
committed records of the final commit (profiles/bench_r2.json and its reference arm) and the host logic that
builds the `config` object, so that a change to bench.py that drops a key is caught without a GPU."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _last_json_line(path):
    with open(path) as fh:
        lines = [ln for ln in fh.read().splitlines() if ln.strip().startswith("{")]
    return json.loads(lines[-1])


def test_committed_bench_record_has_every_contract_key():
    d = _last_json_line(os.path.join(ROOT, "profiles", "bench_r2.json"))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"].startswith("spectra/s") and d["unit"] == "spectra/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["n_gpus"] == 1
    assert abs(d["value"] - d["config"]["total_spectra"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == "spectra/s" and e["h2d_bytes_per_step"] > 10_000 * 131072 * 8 and e["d2h_bytes_per_step"] > 0
    assert 0.9 * d["value"] < e["value"] <= 1.01 * d["value"] and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not any("slowdown" in r for r in c["reasons"])
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    assert 0.5 < r["fp64_pipe_util"] < 1.0 and 0.5 < d["pipeline_fp64"]["fp64_pipe_util"] < 1.0
    b = d["cpu_baseline"]
    assert b["kind"] == "port" and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    p = d["parity_sample"]
    assert p["pass"] and p["float_valued"]["peak_sets_and_lorentzians_bit_exact_vs_oracle"] \
        and p["integer_valued"]["peak_sets_and_lorentzians_bit_exact_vs_oracle"]
    assert p["float_valued"]["mse_max_rel_err_vs_oracle"] <= p["mse_tolerance"] == 1e-9


def test_reference_arm_record_describes_the_same_workload():
    ours = _last_json_line(os.path.join(ROOT, "profiles", "bench_r2.json"))
    ref = _last_json_line(os.path.join(ROOT, "profiles", "bench_reference_arm_r2.json"))
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["higher_is_better"] == ours["higher_is_better"] and ref["steps"] == ours["steps"]
    assert ref["config"]["workload"] == ours["config"]["workload"] and ref["config"]["total_spectra"] == ours["config"]["total_spectra"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["kind"] == "port" and ref["cpu_baseline"]["value"] == ref["value"]


def test_both_arms_build_the_identical_config_object():
    import bench
    args = types.SimpleNamespace(total_spectra=10000, workload="config5", superposition="fast")
    for world in (1, 2, 4, 8):
        a, b = bench.workload_config(args, world), bench.workload_config(args, world)
        assert a == b and a["total_spectra"] == 10000 and a["spectra_per_gpu_per_step"] == 10000 // world
        assert str(world) in a["parallelism"] and "larger than the 126 MB L2" in a["l2"]
    assert bench.sample_indices(10000, 64)[0] == 0 and bench.sample_indices(10000, 64)[-1] == 9999
    assert len(bench.sample_indices(10, 64)) == 10
