#!/bin/bash
# final validation of the round: full parity suite, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s21_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s21_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s21_smoke.log
timeout 900 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s21_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s21_bench.json"))
print("value %.3f e2e %.3f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"]["frac"], d["extra"]["stage_ms"], "prefill frac %.3f"%d["extra"]["prefill_frac_of_tensor_peak"], "decode frac %.3f"%d["extra"]["decode"]["frac"], d["clocks"])
PY
