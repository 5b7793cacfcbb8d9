#!/bin/bash
# round 2, session 1: parity against the reference's own CUDA forward at full depth, the GPU reference timing, a bench line
mkdir -p gpurun_out
nvidia-smi > gpurun_out/r2s1_smi.txt 2>&1
nproc >> gpurun_out/r2s1_smi.txt
timeout 1200 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q -s > gpurun_out/r2s1_ref.log 2>&1
timeout 600 python tools/gpu_reference.py --out gpurun_out/r2s1_gpu_reference.json > gpurun_out/r2s1_gpuref.log 2>&1
timeout 600 python bench.py > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err
tail -5 gpurun_out/r2s1_ref.log
