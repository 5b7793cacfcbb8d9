#!/bin/bash
# round 2, session 2: reference parity at full depth (flash-attention switch fixed), L10 / CLI / ABI-bounds tests, whole GPU suite,
# GPU reference timing, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q -s > gpurun_out/r2s2_ref.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -q -x -k "longrope_switch or cap_chunks or cli_synthetic" > gpurun_out/r2s2_l10.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_vs_reference.py > gpurun_out/r2s2_pytest.log 2>&1
timeout 600 python tools/gpu_reference.py --out gpurun_out/r2s2_gpu_reference.json > gpurun_out/r2s2_gpuref.log 2>&1
timeout 600 python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err
grep -v Warn gpurun_out/r2s2_ref.log | tail -15
tail -5 gpurun_out/r2s2_l10.log
tail -3 gpurun_out/r2s2_pytest.log
tail -2 gpurun_out/r2s2_gpuref.log
