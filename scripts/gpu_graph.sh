#!/bin/bash
# CUDA-graph capture of the whole step (rime_model.GraphedStep) + vectorised interpolation builder:
# full GPU suite, C1 / C3 lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --graph > gpurun_out/bench_c1_graph.json 2> gpurun_out/bench_c1_graph.err; echo "c1 graph rc=$?"; head -c 300 gpurun_out/bench_c1_graph.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err; echo "c3 rc=$?"; head -c 300 gpurun_out/bench_c3_quick.json
