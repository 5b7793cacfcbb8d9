#!/bin/bash
mkdir -p gpurun_out
GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s7_probe_mega.log 2>&1; echo rc=$?; cat gpurun_out/s7_probe_mega.log
GVL_MEGA_INFLIGHT=3 GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s7_probe_mega_if3.log 2>&1; echo rc=$?; cat gpurun_out/s7_probe_mega_if3.log
for ab in 1 2 4 3 7; do
GVL_MEGA_ABLATE=$ab GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s7_probe_ab$ab.log 2>&1; echo "ablate $ab rc=$?"; grep -v "^    \|attention phase" gpurun_out/s7_probe_ab$ab.log
done
GVL_MEGA_INFLIGHT=3 GVL_MEGA_ABLATE=7 GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s7_probe_ab7_if3.log 2>&1; echo "ablate 7 if3 rc=$?"; grep -v "^    \|attention phase" gpurun_out/s7_probe_ab7_if3.log
