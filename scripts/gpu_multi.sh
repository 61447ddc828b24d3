#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "integration or golden_cases" > gpurun_out/pytest_gpu_sub.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_sub.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --workload c3 --nt 1 --steps 2 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err
echo "bench c3 N=$N exit $?"; cat gpurun_out/bench_c3_n$N.json | cut -c1-700; tail -n 5 gpurun_out/bench_c3_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus $N --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err
echo "bench c2 N=$N exit $?"; cat gpurun_out/bench_c2_n$N.json | cut -c1-500; tail -n 5 gpurun_out/bench_c2_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
   bench.py --impl reference --gpus $N --workload c2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "bench ref N=$N exit $?"; cat gpurun_out/bench_ref_n$N.json | cut -c1-300
