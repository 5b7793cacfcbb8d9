#!/bin/bash
# BASELINE configs[3] (Llama-3-8B, 256 decode tokens) and configs[4] (frame-count sweep, prefill only)
mkdir -p gpurun_out
timeout 900 python tools/bench_cfg4.py --decode 256 --steps 2 > gpurun_out/s20_cfg4.json 2> gpurun_out/s20_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/s20_cfg4.json; tail -3 gpurun_out/s20_cfg4.err
GVL_DECODE_MEGA=0 timeout 900 python tools/bench_cfg4.py --decode 256 --steps 2 > gpurun_out/s20_cfg4_chain.json 2> gpurun_out/s20_cfg4_chain.err; echo "cfg4 chain rc=$?"; cat gpurun_out/s20_cfg4_chain.json
timeout 1200 python tools/sweep.py --frames 16,32,64,96,128,192,256 --batch 1,4 > gpurun_out/s20_sweep.md 2> gpurun_out/s20_sweep.err; echo "sweep rc=$?"; cat gpurun_out/s20_sweep.md; tail -3 gpurun_out/s20_sweep.err
