#!/bin/bash
# two GPUs of one box: the non-current-device test, C3 / C5 at N=2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "non_current" > gpurun_out/pytest_two.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_two.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; echo "c3 x2 rc=$?"; head -c 300 gpurun_out/bench_c3_2gpu.json
