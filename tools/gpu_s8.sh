#!/bin/bash
mkdir -p gpurun_out
for cfg in "7 1" "15 1" "15 2" "15 3" "7 2"; do
set -- $cfg
GVL_MEGA_ABLATE=$1 GVL_MEGA_INFLIGHT=$2 GVL_DECODE_MEGA=1 GVL_MEGA_TRACE=1 timeout 300 python tools/probe_decode.py 3483 32 > gpurun_out/s8_probe_ab$1_if$2.log 2>&1; echo "ablate $1 inflight $2 rc=$?"; grep "mode\|wall\|landed\|total" gpurun_out/s8_probe_ab$1_if$2.log
done
# clocks / power while the decode kernel runs (512 steps ~ 1.4 s per generate, a few generates)
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 50 > gpurun_out/s8_smi_mega.csv &
SMI=$!
GVL_DECODE_MEGA=1 timeout 300 python tools/probe_decode.py 3000 512 > gpurun_out/s8_probe_long_mega.log 2>&1; cat gpurun_out/s8_probe_long_mega.log
kill $SMI
sort gpurun_out/s8_smi_mega.csv | uniq -c | sort -rn | head -12
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 50 > gpurun_out/s8_smi_chain.csv &
SMI=$!
GVL_DECODE_MEGA=0 timeout 300 python tools/probe_decode.py 3000 512 > gpurun_out/s8_probe_long_chain.log 2>&1; cat gpurun_out/s8_probe_long_chain.log
kill $SMI
sort gpurun_out/s8_smi_chain.csv | uniq -c | sort -rn | head -8
