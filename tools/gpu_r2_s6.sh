#!/bin/bash
# round 2, session 6: exchange probe with single-copy-atomic packets; GEMM epilogue changes (GELU exp2-poly, pipelined TMEM loads, .cta arrive)
mkdir -p gpurun_out
timeout 300 tools/probe_exchange 2000 > gpurun_out/r2s6_exchange.log 2>&1; grep -c mismatch gpurun_out/r2s6_exchange.log; grep -v mismatch gpurun_out/r2s6_exchange.log | grep "work 4000" 
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or full_width or pipeline" > gpurun_out/r2s6_pytest.log 2>&1; tail -3 gpurun_out/r2s6_pytest.log
timeout 300 python tools/probe_gemm_shape.py > gpurun_out/r2s6_gemm_shapes.log 2>&1; cat gpurun_out/r2s6_gemm_shapes.log | tail -12
