#!/bin/bash
# decode consumer rewrite (weights as the A operand): parity of every decode test, then the phase trace and the plain timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or generate or greedy or lm_" > gpurun_out/r2s16_pytest.log 2>&1; tail -5 gpurun_out/r2s16_pytest.log
GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/r2s16_trace.log 2>&1; tail -40 gpurun_out/r2s16_trace.log
timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/r2s16_plain.log 2>&1; tail -6 gpurun_out/r2s16_plain.log
cp gpurun_out/decode_trace_raw.npy gpurun_out/r2s16_trace_raw.npy 2>/dev/null
