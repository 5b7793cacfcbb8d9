#!/bin/bash
# preprocessing kernels: parity; bench with the raw-frame e2e leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess.py -m gpu -x -q > gpurun_out/s19_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s19_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s19_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s19_bench.json"))
print("value %.3f e2e %.3f raw %s"%(d["value"], d["e2e"]["value"], d["extra"]["e2e_from_raw_frames"]))
PY
