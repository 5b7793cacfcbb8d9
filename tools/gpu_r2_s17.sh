#!/bin/bash
# same-box A/B of the decode step: library variants built from HEAD (base), the working tree (new), and new with 2 prefetch copies per lane
mkdir -p gpurun_out
L=grounded-video-llm_b200/gvl/libgvl.so
cp $L /tmp/libgvl_keep.so
for round in 1 2; do
  for v in base new ahead2; do
    cp tools/_variants/libgvl_$v.so $L
    echo "== $v (round $round)"; timeout 300 python tools/probe_decode.py 3483 32 2>&1 | tail -2
  done
done > gpurun_out/r2s17_ab.log 2>&1
cp /tmp/libgvl_keep.so $L
cat gpurun_out/r2s17_ab.log
