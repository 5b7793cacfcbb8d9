#!/bin/bash
# round 2, session 12: whole GPU suite + bench (N=1) with the round-2 bench.py + reference arm on this box's host cores
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2s12_pytest.log 2>&1; tail -4 gpurun_out/r2s12_pytest.log
timeout 900 python bench.py > gpurun_out/r2s12_bench.json 2> gpurun_out/r2s12_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2s12_bench.json; tail -3 gpurun_out/r2s12_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s12_bench_ref.json 2> gpurun_out/r2s12_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2s12_bench_ref.json
