#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched_greedy or pipeline or eos_padding" > gpurun_out/r2s14_pytest.log 2>&1; tail -5 gpurun_out/r2s14_pytest.log
timeout 300 python tools/probe_gemv_batch.py > gpurun_out/r2s14_gemv_batch.log 2>&1; cat gpurun_out/r2s14_gemv_batch.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clips-per-gpu 4 > gpurun_out/r2s14_bench_c4.json 2> gpurun_out/r2s14_bench_c4.err; echo "c4 rc=$?"; cut -c1-260 gpurun_out/r2s14_bench_c4.json; tail -2 gpurun_out/r2s14_bench_c4.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s14_bench_c1.json 2> gpurun_out/r2s14_bench_c1.err; cut -c1-260 gpurun_out/r2s14_bench_c1.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2s14_ncu_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/r2_launches_bench.csv
