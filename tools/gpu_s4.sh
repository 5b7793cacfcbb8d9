#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./tools/probe_stream > gpurun_out/s4_stream.log 2>&1; echo rc=$?; cat gpurun_out/s4_stream.log
timeout 600 python -m pytest tests -m gpu -x -q -k "long_context" > gpurun_out/s4_pytest_long.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s4_pytest_long.log
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw --format=csv
