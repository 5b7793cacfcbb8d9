#!/bin/bash
mkdir -p gpurun_out
L=grounded-video-llm_b200/gvl/libgvl.so
cp $L /tmp/libgvl_keep.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or generate or greedy or lm_" > gpurun_out/r2s25_pytest.log 2>&1; tail -3 gpurun_out/r2s25_pytest.log
for round in 1 2; do
  for v in base notrace; do
    cp tools/_variants/libgvl_$v.so $L
    echo "== $v (round $round)"; timeout 300 python tools/probe_decode.py 3483 32 2>&1 | tail -1
  done
done > gpurun_out/r2s25_ab.log 2>&1
cp /tmp/libgvl_keep.so $L
cat gpurun_out/r2s25_ab.log
echo "== 15 steps per launch (what bench.py times)"; timeout 300 python tools/probe_decode.py 3483 15 2>&1 | tail -1
GVL_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 2 -c 1 -o gpurun_out/r2b_decode_mega -f python tools/probe_decode.py 3483 8 > gpurun_out/r2s25_ncu_decode.log 2>&1; tail -1 gpurun_out/r2s25_ncu_decode.log
