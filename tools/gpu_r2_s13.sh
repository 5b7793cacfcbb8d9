#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2s13_pytest.log 2>&1; tail -4 gpurun_out/r2s13_pytest.log
timeout 900 python bench.py > gpurun_out/r2s13_bench.json 2> gpurun_out/r2s13_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2s13_bench.json; tail -3 gpurun_out/r2s13_bench.err
timeout 600 python tools/bench_cfg4.py > gpurun_out/r2s13_cfg4.json 2> gpurun_out/r2s13_cfg4.err; cat gpurun_out/r2s13_cfg4.json; tail -2 gpurun_out/r2s13_cfg4.err
bash tools/gpu_sanitize.sh
