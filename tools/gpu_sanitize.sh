#!/bin/bash
# compute-sanitizer over the small-shape parity tests: memcheck for every kernel family, racecheck (shared-memory
# hazards) for the kernels with hand-rolled producer / consumer protocols. Logs -> gpurun_out/r2_sanitizer_*.log (copied to profiles/).
mkdir -p gpurun_out
SEL_SMALL="lm_prefill_logits_and_greedy_decode or eos_padding or longrope_switch or cap_chunks or norm_kernels or rope_and_cache or pool_concat"
SEL_TC="test_gemm_epilogues_vs_oracle_rounding and (2049 or 700 or 577) or test_attention_vs_oracle"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_SMALL" > gpurun_out/r2_sanitizer_memcheck_small.log 2>&1; echo "memcheck small rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_small.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_TC" > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1; echo "memcheck tc rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_tc.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lm_prefill_logits_and_greedy_decode or longrope_switch" > gpurun_out/r2_sanitizer_racecheck_decode.log 2>&1; echo "racecheck decode rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck_decode.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_attention_vs_oracle and 577 or test_gemm_epilogues_vs_oracle_rounding and 2049-6144" > gpurun_out/r2_sanitizer_racecheck_tc.log 2>&1; echo "racecheck tc rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck_tc.log
