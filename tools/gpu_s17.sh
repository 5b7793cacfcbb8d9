#!/bin/bash
# full parity suite + smoke + bench with attention v3 (default) and v2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s17_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/s17_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s17_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s17_bench_v3.json 2> gpurun_out/s17_bench_v3.err; echo "bench v3 rc=$?"
python - <<'PY'
import json
for n in ("v3",):
    d=json.load(open("gpurun_out/s17_bench_%s.json"%n))
    print(n, "value %.3f e2e %.3f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"]["frac"], d["extra"]["stage_ms_profiled"], d["extra"]["attention"], d["extra"]["decode_gemv"], d["clocks"])
PY
GVL_ATTN_V2=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s17_bench_v2.json 2> gpurun_out/s17_bench_v2.err; echo "bench v2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/s17_bench_v2.json"))
print("v2", "value %.3f e2e %.3f ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"]["frac"], d["extra"]["stage_ms_profiled"], d["extra"]["attention"], d["clocks"])
PY
