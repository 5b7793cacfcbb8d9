#!/bin/bash
mkdir -p gpurun_out
PROBE_CHILD=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r2_attn_iv2 -f python tools/probe_attn_tc.py d96_iv2_b12 > gpurun_out/r2s28_ncu_attn_iv2.log 2>&1; tail -1 gpurun_out/r2s28_ncu_attn_iv2.log
PROBE_CHILD=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r2_attn_phi -f python tools/probe_attn_tc.py d96_causal_long > gpurun_out/r2s28_ncu_attn_phi.log 2>&1; tail -1 gpurun_out/r2s28_ncu_attn_phi.log
