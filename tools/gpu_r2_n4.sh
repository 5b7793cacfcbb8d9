#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err; echo "n$N rc=$?"; cut -c1-260 gpurun_out/r2_n${N}_bench.json; tail -3 gpurun_out/r2_n${N}_bench.err | cut -c1-300
