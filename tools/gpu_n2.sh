#!/bin/bash
# 2-GPU validation: NCCL all-gather path of the bench (weak scaling, 1 and 4 clips per GPU)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
cat gpurun_out/n2_bench.json; tail -5 gpurun_out/n2_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --clips-per-gpu 4 > gpurun_out/n2_bench_c4.json 2> gpurun_out/n2_bench_c4.err; echo "bench n2 c4 rc=$?"
cat gpurun_out/n2_bench_c4.json; tail -5 gpurun_out/n2_bench_c4.err
