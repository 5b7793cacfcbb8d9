#!/bin/bash
# first GPU call of a session: parity suite (mega decode on / off), smoke, bench both decode modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest_mega.log 2>&1; echo "pytest mega rc=$?"
tail -5 gpurun_out/s1_pytest_mega.log
GVL_DECODE_MEGA=0 timeout 600 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s1_pytest_chain.log 2>&1; echo "pytest chain rc=$?"
tail -3 gpurun_out/s1_pytest_chain.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s1_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench_mega.json 2> gpurun_out/s1_bench_mega.err; echo "bench mega rc=$?"
cat gpurun_out/s1_bench_mega.json
GVL_DECODE_MEGA=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_chain.json 2> gpurun_out/s1_bench_chain.err; echo "bench chain rc=$?"
cat gpurun_out/s1_bench_chain.json
