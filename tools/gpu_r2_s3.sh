#!/bin/bash
# round 2, session 3: exchange-primitive micro-benchmark under HBM load; ncu --set full of the 2-CTA GEMM on the dominant shapes
mkdir -p gpurun_out
timeout 300 tools/probe_exchange 2000 > gpurun_out/r2s3_exchange.log 2>&1; tail -3 gpurun_out/r2s3_exchange.log
timeout 300 python tools/probe_gemm_shape.py > gpurun_out/r2s3_gemm_shapes.log 2>&1; cat gpurun_out/r2s3_gemm_shapes.log | tail -12
for s in iv2_fc1 iv2_fc2 iv2_qkv phi_gate_up; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 3 -c 1 -o gpurun_out/r2_gemm_$s -f python tools/probe_gemm_shape.py $s 2 > gpurun_out/r2s3_ncu_$s.log 2>&1; tail -1 gpurun_out/r2s3_ncu_$s.log
done
PROBE_CHILD=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 1 -c 1 -o gpurun_out/r2_attn_clip -f python tools/probe_attn_tc.py d64_clip_b12 > gpurun_out/r2s3_ncu_attn_clip.log 2>&1; tail -1 gpurun_out/r2s3_ncu_attn_clip.log
ls -la gpurun_out/*.ncu-rep
