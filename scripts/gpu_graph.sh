#!/bin/bash
# CUDA-graph capture of the whole step (rime_model.GraphedStep): parity test, C1 / C2 lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graphed or golden_cases" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_graph.log | cut -c1-300
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --graph > gpurun_out/bench_c1_graph.json 2> gpurun_out/bench_c1_graph.err; echo "c1 graph rc=$?"; head -c 400 gpurun_out/bench_c1_graph.json; tail -5 gpurun_out/bench_c1_graph.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --graph > gpurun_out/bench_c2_graph.json 2> gpurun_out/bench_c2_graph.err; echo "c2 graph rc=$?"; head -c 400 gpurun_out/bench_c2_graph.json; tail -5 gpurun_out/bench_c2_graph.err
