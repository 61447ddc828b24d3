#!/bin/bash
# 8-GPU weak-scaling point of the default C3 bench (2 times per GPU), final code of round 2
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err
echo "bench c3 N=$N exit $?"; cut -c1-400 gpurun_out/bench_c3_${N}gpu.json; tail -n 3 gpurun_out/bench_c3_${N}gpu.err
