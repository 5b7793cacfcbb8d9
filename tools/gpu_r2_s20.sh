#!/bin/bash
mkdir -p gpurun_out
GVL_MEGA_TRACE=1 GVL_PROBE_REPS=3 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/r2s20_trace.log 2>&1; tail -30 gpurun_out/r2s20_trace.log
cp gpurun_out/decode_trace_raw.npy gpurun_out/r2s20_trace_raw.npy
