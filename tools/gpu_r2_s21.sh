#!/bin/bash
# batched decode: rope + attention of all sequences in one launch per layer; parity, then 4 clips/GPU against 1 clip/GPU on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or generate or greedy or lm_ or batched" > gpurun_out/r2s21_pytest.log 2>&1; tail -3 gpurun_out/r2s21_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clips-per-gpu 4 > gpurun_out/r2s21_bench_c4.json 2> gpurun_out/r2s21_bench_c4.err; echo "c4 rc=$?"; cut -c1-200 gpurun_out/r2s21_bench_c4.json; tail -2 gpurun_out/r2s21_bench_c4.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s21_bench_c1.json 2> gpurun_out/r2s21_bench_c1.err; echo "c1 rc=$?"; cut -c1-200 gpurun_out/r2s21_bench_c1.json
