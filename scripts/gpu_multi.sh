#!/bin/bash
# multi-GPU bench lines: scripts/gpu_multi.sh <N> <workload...>  (one rank per GPU, NCCL)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$1; shift
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for w in "$@"; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 2 --warmup 3 \
      --no-cpu-baseline > gpurun_out/bench_${w}_${N}gpu.json 2> gpurun_out/bench_${w}_${N}gpu.err
  echo "bench $w N=$N rc=$? $(head -c 420 gpurun_out/bench_${w}_${N}gpu.json)"
done
