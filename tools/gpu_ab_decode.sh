#!/bin/bash
# Same-box A/B of the decode step. Box-to-box (and run-to-run) differences of +-5 % are larger than most kernel changes, so variants
# are compared INSIDE one gpurun job: build each variant's libgvl.so on the CPU box into tools/_variants/libgvl_<name>.so
# (git-ignored, travels with the snapshot), then
#     gpurun -- 'bash tools/gpu_ab_decode.sh base new [more ...]'
# runs the decode parity tests on the working-tree library, swaps the variants in turn (two rounds, interleaved) under
# tools/probe_decode.py (decode launch alone, median of 7 x 32 steps), restores the library and prints the phase trace of it.
# The round-2 logs profiles/r2_decode_ab_*.log were produced this way.
mkdir -p gpurun_out
L=grounded-video-llm_b200/gvl/libgvl.so
cp $L /tmp/libgvl_keep.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or generate or greedy or lm_" > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
for round in 1 2; do
  for v in "$@"; do
    cp tools/_variants/libgvl_$v.so $L
    echo "== $v (round $round)"; timeout 300 python tools/probe_decode.py 3483 32 2>&1 | tail -1
  done
done > gpurun_out/ab_decode.log 2>&1
cp /tmp/libgvl_keep.so $L
cat gpurun_out/ab_decode.log
GVL_MEGA_TRACE=1 GVL_PROBE_REPS=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/ab_trace.log 2>&1; tail -30 gpurun_out/ab_trace.log
