#!/bin/bash
# ncu --set full of attention v3 and v2 on the InternVideo2 shape (B=12, H=16, S=2049, d=96)
mkdir -p gpurun_out
PROBE_CHILD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc3 -s 1 -c 1 -o gpurun_out/r1_attn_v3 -f python tools/probe_attn_tc.py d96_iv2_b12 > gpurun_out/s16_ncu_v3.log 2>&1; tail -2 gpurun_out/s16_ncu_v3.log
GVL_ATTN_V2=1 PROBE_CHILD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r1_attn_v2 -f python tools/probe_attn_tc.py d96_iv2_b12 > gpurun_out/s16_ncu_v2.log 2>&1; tail -2 gpurun_out/s16_ncu_v2.log
ls -la gpurun_out/*.ncu-rep
