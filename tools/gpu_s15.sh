#!/bin/bash
# attention v3 (double-buffered 64-key score tiles): correctness + timing vs v2 per shape, then the attention parity tests
mkdir -p gpurun_out
echo "== v3"; timeout 600 python tools/probe_attn_tc.py 2>&1 | tee gpurun_out/s15_attn_v3.log
echo "== v2"; GVL_ATTN_V2=1 timeout 600 python tools/probe_attn_tc.py d96_iv2_pad d96_iv2_b12 d64_clip_b12 d96_causal_long d128_llama_long 2>&1 | tee gpurun_out/s15_attn_v2.log
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or stage or full_width or pipeline" > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s15_pytest.log
