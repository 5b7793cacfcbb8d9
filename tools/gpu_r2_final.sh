#!/bin/bash
# round-2 validation: whole GPU suite, smoke, bench line (both arms), 4 clips/GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clips-per-gpu 4 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "c4 rc=$?"; cut -c1-260 gpurun_out/r2f_bench_c4.json
SECONDS=0; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$? wall=${SECONDS}s"; cut -c1-400 gpurun_out/r2f_bench_ref.json; tail -3 gpurun_out/r2f_bench_ref.err
timeout 600 python tools/bench_cfg4.py > gpurun_out/r2f_cfg4.json 2> gpurun_out/r2f_cfg4.err; echo "cfg4 rc=$?"; cut -c1-600 gpurun_out/r2f_cfg4.json
GVL_PROBE_REPS=7 timeout 300 python tools/probe_decode.py 3483 32 2>&1 | tail -1
