#!/bin/bash
# 8-GPU weak-scaling point of the C3 bench only
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --workload c3 --nt 1 --steps 2 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench c3 N=$N exit $?"; cut -c1-400 gpurun_out/bench_c3_n$N.json; tail -n 3 gpurun_out/bench_c3_n$N.err
