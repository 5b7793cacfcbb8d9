#!/bin/bash
# per-head attention split: decode parity tests on the working-tree build, then same-box A/B against the previous build
mkdir -p gpurun_out
L=grounded-video-llm_b200/gvl/libgvl.so
cp $L /tmp/libgvl_keep.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or generate or greedy or lm_" > gpurun_out/r2s24_pytest.log 2>&1; tail -3 gpurun_out/r2s24_pytest.log
for round in 1 2; do
  for v in base flat; do
    cp tools/_variants/libgvl_$v.so $L
    echo "== $v (round $round)"; timeout 300 python tools/probe_decode.py 3483 32 2>&1 | tail -1
  done
done > gpurun_out/r2s24_ab.log 2>&1
cp /tmp/libgvl_keep.so $L
cat gpurun_out/r2s24_ab.log
GVL_MEGA_TRACE=1 GVL_PROBE_REPS=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/r2s24_trace.log 2>&1; tail -28 gpurun_out/r2s24_trace.log
cp gpurun_out/decode_trace_raw.npy gpurun_out/r2s24_trace_raw.npy
