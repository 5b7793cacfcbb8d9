#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or full_width or pipeline" > gpurun_out/r2s7_pytest.log 2>&1; tail -3 gpurun_out/r2s7_pytest.log
timeout 300 python tools/probe_gemm_shape.py > gpurun_out/r2s7_gemm_shapes.log 2>&1; cat gpurun_out/r2s7_gemm_shapes.log | tail -12
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 3 -c 1 -o gpurun_out/r2b_gemm_iv2_fc1 -f python tools/probe_gemm_shape.py iv2_fc1 2 > gpurun_out/r2s7_ncu_fc1.log 2>&1; tail -1 gpurun_out/r2s7_ncu_fc1.log
