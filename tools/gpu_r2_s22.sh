#!/bin/bash
# end of round 2: compute-sanitizer over the decode kernels changed late in the round (single-kernel step with the new fragment mapping /
# attention split, batched rope + attention launches), and one ncu --set full capture of the final decode_mega_kernel
mkdir -p gpurun_out
SEL="lm_prefill_logits_and_greedy_decode or eos_padding or longrope_switch or cap_chunks or batched_greedy"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_memcheck_decode_final.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_decode_final.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lm_prefill_logits_and_greedy_decode" > gpurun_out/r2_sanitizer_racecheck_decode_final.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck_decode_final.log
GVL_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 2 -c 1 -o gpurun_out/r2_decode_mega -f python tools/probe_decode.py 3483 8 > gpurun_out/r2s22_ncu_decode.log 2>&1; tail -2 gpurun_out/r2s22_ncu_decode.log
ls -la gpurun_out/r2_decode_mega.ncu-rep
