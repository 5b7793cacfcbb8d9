#!/bin/bash
# session re-entry check: parity suite, smoke, bench (chain decode = default, then single-kernel decode), decode probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s9_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s9_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s9_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s9_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s9_bench_chain.json 2> gpurun_out/s9_bench_chain.err; echo "bench chain rc=$?"
cat gpurun_out/s9_bench_chain.json
GVL_DECODE_MEGA=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_mega.json 2> gpurun_out/s9_bench_mega.err; echo "bench mega rc=$?"
cat gpurun_out/s9_bench_mega.json
GVL_DECODE_MEGA=0 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s9_probe_chain.log 2>&1; cat gpurun_out/s9_probe_chain.log
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s9_probe_mega.log 2>&1; cat gpurun_out/s9_probe_mega.log
