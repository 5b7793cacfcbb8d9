#!/bin/bash
# ring-fed K/V in the single-kernel decode step: parity, then phase trace at inflight 1/2/3
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/s10_pytest.log
for inf in 1 2 3; do
GVL_MEGA_INFLIGHT=$inf GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s10_probe_mega_if$inf.log 2>&1; echo "inflight $inf rc=$?"; cat gpurun_out/s10_probe_mega_if$inf.log
done
