#!/bin/bash
# L2 prefetcher warp in the single-kernel decode step: parity, then prefetch window x inflight sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s12_pytest.log
for cfg in "0 1" "40 1" "80 1" "160 1" "40 2" "80 3" "160 3"; do
set -- $cfg
GVL_MEGA_PFWIN=$1 GVL_MEGA_INFLIGHT=$2 GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s12_probe_pf$1_if$2.log 2>&1; echo "pfwin $1 inflight $2 rc=$?"; grep "mode\|qkv \|attn \|o_proj\|gate_up\|down \|wall\|landed" gpurun_out/s12_probe_pf$1_if$2.log
done
