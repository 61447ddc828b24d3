#!/bin/bash
# C4 after the fused Jones sandwich: 4-pol GPU parity tests, then the bench line at N=1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "4pol or c4 or pol or multimodel" > gpurun_out/pytest_pol.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pol.log
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?"; head -c 500 gpurun_out/bench_c4.json
