#!/bin/bash
# flags-in-data decode: parity at full width, then timing LL vs barriers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "flags_in_data or lm_ or eos or sampling or pipeline" > gpurun_out/s22_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s22_pytest.log
for ll in 1 0; do
GVL_MEGA_LL=$ll GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s22_probe_ll$ll.log 2>&1; echo "LL=$ll rc=$?"; grep "mode\|qkv \|attn \|o_proj\|gate_up\|down \|wall\|timeout" gpurun_out/s22_probe_ll$ll.log
done
