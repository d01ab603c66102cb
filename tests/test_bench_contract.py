"""The bench line contract (driver README): the last line measured on the B200 and kept under profiles/ must
carry every key the driver and the judge read, with consistent arithmetic.  CPU-only: parses stored JSON."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_single_gpu_line_has_the_contract_keys():
    d = _load("r1_bench_reddit_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and r["traffic"] < r["algorithmic_bytes"]            # measured DRAM bytes << gathered bytes
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                           # copies are inside the e2e timed region
    assert d["gpu_launches"] == d["steps"] * d["launches_per_step"] > 0
    # value = 2 * nnz * dim / t
    flops = 2.0 * d["config"]["stored_entries"] * d["config"]["dim"]
    assert abs(d["value"] - flops / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_multi_gpu_lines_are_whole_job_aggregates():
    one = _load("r1_bench_reddit_final.json")
    eight = _load("r1d_scale_bench_reddit_8.json")
    assert eight["n_gpus"] == 8 and eight["scaling"] == "strong" and eight["metric"] == one["metric"]
    assert eight["config"]["stored_entries"] == one["config"]["stored_entries"]       # same total work
    assert 4.0 < eight["value"] / one["value"] < 8.0
    ph = eight["config"]["phases"]
    assert ph["exchange_rows_vs_allgather"] < 1.0 and ph["exchange_only_ms"] + ph["kernel_only_ms"] <= 1.05 * eight["ms_per_step"] + 0.1


def test_reference_arm_line():
    d = _load("r1_bench_reference_cpu_v2.json")
    assert d["impl"] == "reference" and d["metric"] == "spmm_gflops" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
