#!/bin/bash
# round 2, session 8: attention v2 with single-pass softmax + XU ping-pong; GEMM epilogue prefetch
mkdir -p gpurun_out
timeout 900 python tools/probe_attn_tc.py > gpurun_out/r2s8_attn.log 2>&1; cat gpurun_out/r2s8_attn.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention or gemm or full_width or pipeline" > gpurun_out/r2s8_pytest.log 2>&1; tail -3 gpurun_out/r2s8_pytest.log
timeout 300 python tools/probe_gemm_shape.py > gpurun_out/r2s8_gemm_shapes.log 2>&1; cat gpurun_out/r2s8_gemm_shapes.log | tail -12
