#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/probe_mufu > gpurun_out/r2s9_mufu.log 2>&1; cat gpurun_out/r2s9_mufu.log
for pp in 0 1; do
  GVL_ATTN_PINGPONG=$pp timeout 600 python tools/probe_attn_tc.py d96_iv2_b12 d64_clip_b12 d96_causal_long d128_llama_long > gpurun_out/r2s9_attn_pp$pp.log 2>&1; echo "pp=$pp"; sed 's/rows32.*time/time/' gpurun_out/r2s9_attn_pp$pp.log
done
