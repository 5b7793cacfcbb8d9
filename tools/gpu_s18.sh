#!/bin/bash
# round-1 evidence run: final bench line, ncu launch list of the bench, ncu --set full of the decode kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/s18_bench.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s18_bench_ref.json 2> gpurun_out/s18_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/s18_bench_ref.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s18_ncu_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r1b_launches_bench.csv
GVL_DECODE_MEGA=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -s 2 -c 1 -o gpurun_out/r1_decode_mega -f python tools/probe_decode.py 3483 8 > gpurun_out/s18_ncu_mega.log 2>&1; echo "ncu mega rc=$?"; tail -3 gpurun_out/s18_ncu_mega.log
ls -la gpurun_out/*.ncu-rep
