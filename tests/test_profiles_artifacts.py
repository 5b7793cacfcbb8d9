"""The bench lines committed under profiles/ (copied from the GPU validation runs of the round) carry every key of the bench.py
contract, and the numbers the documents quote are the ones in the files. CPU-only: reads JSON."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _line(name):
    path = os.path.join(PROF, name)
    if not os.path.exists(path):
        pytest.skip("%s not committed" % name)
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name, n_gpus", [("r2_final_bench.json", 1), ("r2_final_bench_c4.json", 1), ("r2_n2_bench.json", 2),
                                          ("r2_n4_bench.json", 4), ("r2_n8_bench.json", 8)])
def test_committed_bench_lines_carry_the_contract(name, n_gpus):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == n_gpus and d["unit"] == "videos/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # whole-job throughput: clips of all ranks / max-over-ranks step time
    clips = n_gpus * (4 if name.endswith("_c4.json") else 1)
    assert abs(d["value"] - clips / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_multi_gpu_runs_computed_the_single_gpu_tokens():
    hashes = {n: _line(n)["extra"]["tokens_clip0_sha256_16"] for n in ("r2_final_bench.json", "r2_n2_bench.json", "r2_n4_bench.json",
                                                                       "r2_n8_bench.json")}
    assert len(set(hashes.values())) == 1, hashes
    for n in ("r2_n2_bench.json", "r2_n4_bench.json", "r2_n8_bench.json"):
        ex = _line(n)["extra"]
        assert ex["strong_scaling_tokens_sha256_16"] == ex["tokens_clip0_sha256_16"]
        assert ex["strong_scaling_1_clip_ms"] < _line("r2_final_bench.json")["ms_per_step"]


def test_reference_arm_line_and_cpu_baseline():
    r = _line("r2_final_bench_reference_arm.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    assert r["cpu_baseline"]["kind"] == "reference" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["extra"]["cfg1_cpu_end_to_end"]["seconds_per_video"] > 0
    d = _line("r2_final_bench.json")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["unit"] == d["unit"] and 0 < cb["value"] < d["value"]


def test_decode_roofline_is_the_launch_alone_median():
    dec = _line("r2_final_bench.json")["extra"]["decode"]
    s = sorted(dec["samples_ms_per_step"])
    assert dec["ms_per_step"] == s[len(s) // 2]
    assert abs(dec["frac"] - dec["bytes_per_step"] / (dec["ms_per_step"] * 1e-3) / 1e9 / dec["peak_gbs"]) < 1e-9
    assert 0.55 < dec["frac"] < 0.80
