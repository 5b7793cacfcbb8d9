#!/bin/bash
mkdir -p gpurun_out
GVL_DECODE_MEGA=1 timeout 900 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s6_pytest_mega.log 2>&1; echo "pytest mega rc=$?"
tail -5 gpurun_out/s6_pytest_mega.log
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s6_probe_mega.log 2>&1; echo rc=$?; cat gpurun_out/s6_probe_mega.log
for inf in 1 2 3; do
GVL_MEGA_INFLIGHT=$inf GVL_DECODE_MEGA=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s6_probe_mega_if$inf.log 2>&1; echo rc=$?; cat gpurun_out/s6_probe_mega_if$inf.log
done
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 64 32 > gpurun_out/s6_probe_mega_short.log 2>&1; echo rc=$?; cat gpurun_out/s6_probe_mega_short.log
