#!/bin/bash
mkdir -p gpurun_out
GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s2_probe_mega.log 2>&1; echo rc=$?; cat gpurun_out/s2_probe_mega.log
GVL_DECODE_MEGA=0 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s2_probe_chain.log 2>&1; echo rc=$?; cat gpurun_out/s2_probe_chain.log
GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 64 32 > gpurun_out/s2_probe_mega_short.log 2>&1; echo rc=$?; cat gpurun_out/s2_probe_mega_short.log
