#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/probe_attn_tc.py > gpurun_out/r2s11_attn.log 2>&1; sed 's/rows32.*time/time/; s/rows32.*//' gpurun_out/r2s11_attn.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention or full_width or pipeline" > gpurun_out/r2s11_pytest.log 2>&1; tail -3 gpurun_out/r2s11_pytest.log
