#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r2_n2_pytest.log 2>&1; tail -12 gpurun_out/r2_n2_pytest.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; echo "n2 rc=$?"; cut -c1-300 gpurun_out/r2_n2_bench.json; tail -3 gpurun_out/r2_n2_bench.err | cut -c1-300
