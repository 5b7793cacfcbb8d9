#!/bin/bash
mkdir -p gpurun_out
for pp in 0 1; do
  GVL_ATTN_TRACE=1 GVL_ATTN_PINGPONG=$pp PROBE_CHILD=1 timeout 300 python tools/probe_attn_tc.py d96_iv2_b12 > gpurun_out/r2s10_trace_pp$pp.log 2>&1
  grep -A 28 "attn trace" gpurun_out/r2s10_trace_pp$pp.log | head -32
done
