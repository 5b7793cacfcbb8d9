#!/bin/bash
# ring-fed K/V: parity, then ablations (8 = no GEMV math, 1 = no grid barriers, 4 = no staging, 2 = no attention)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s11_pytest.log
for cfg in "0 1" "8 1" "8 3" "1 1" "5 1" "13 1" "15 1" "15 3"; do
set -- $cfg
GVL_MEGA_ABLATE=$1 GVL_MEGA_INFLIGHT=$2 GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s11_probe_ab$1_if$2.log 2>&1; echo "ablate $1 inflight $2 rc=$?"; grep "mode\|qkv \|attn \|o_proj\|gate_up\|down \|wall" gpurun_out/s11_probe_ab$1_if$2.log
done
