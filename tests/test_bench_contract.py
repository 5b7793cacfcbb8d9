"""bench.py contract pieces that can be checked without a GPU: the reference arm's JSON line and the loud failure of the
product arm on a box without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    from oracle import ref_shims
    if not ref_shims.available():
        import pytest
        pytest.skip("oracle/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
    # GVL_REF_CFG1=0: skip the full-depth configs[0] end-to-end leg here (22 GB of host memory, minutes on 8 cores)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, GVL_REF_CFG1="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "videos/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "Phi-3.5-3.8B" in d["config"]["workload"] and d["config"]["samples_timed"] == 1


def test_product_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
