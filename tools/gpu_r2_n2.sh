#!/bin/bash
# 2 GPUs: the multi-rank path (all-to-all exchange of visual tokens, device-side token gather, clip-0 hash, strong-scaling latency)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; echo "n2 rc=$?"; cut -c1-300 gpurun_out/r2_n2_bench.json; tail -3 gpurun_out/r2_n2_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --clips-per-gpu 4 > gpurun_out/r2_n2_bench_c4.json 2> gpurun_out/r2_n2_bench_c4.err; echo "n2 c4 rc=$?"; cut -c1-300 gpurun_out/r2_n2_bench_c4.json; tail -3 gpurun_out/r2_n2_bench_c4.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n1_bench_same_box.json 2> gpurun_out/r2_n1_bench_same_box.err; cut -c1-200 gpurun_out/r2_n1_bench_same_box.json
