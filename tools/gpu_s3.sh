#!/bin/bash
# decode megakernel v2 bring-up: parity (mega on), then probes with the phase trace
mkdir -p gpurun_out
GVL_DECODE_MEGA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline or decode" > gpurun_out/s3_pytest_mega.log 2>&1; echo "pytest mega rc=$?"
tail -15 gpurun_out/s3_pytest_mega.log
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s3_probe_mega.log 2>&1; echo rc=$?; cat gpurun_out/s3_probe_mega.log
GVL_DECODE_MEGA=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s3_probe_mega_notrace.log 2>&1; echo rc=$?; cat gpurun_out/s3_probe_mega_notrace.log
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 64 32 > gpurun_out/s3_probe_mega_short.log 2>&1; echo rc=$?; cat gpurun_out/s3_probe_mega_short.log
