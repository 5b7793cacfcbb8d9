#!/bin/bash
# dynamic inflight depth sweep (ahead, current), single-kernel decode default on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lm_ or eos or pipeline" > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s13_pytest.log
for cfg in "1 1" "1 2" "1 3" "2 3"; do
set -- $cfg
GVL_MEGA_INFLIGHT=$1 GVL_MEGA_INFLIGHT_CUR=$2 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s13_probe_if$1_cur$2.log 2>&1; echo "inflight ahead $1 cur $2 rc=$?"; grep "mode\|qkv \|attn \|o_proj\|gate_up\|down \|wall" gpurun_out/s13_probe_if$1_cur$2.log
done
