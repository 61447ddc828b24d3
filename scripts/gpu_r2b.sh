#!/bin/bash
# tensor-core kernels after a change: probe (parity + timing), GPU tests, short C3 bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/tc_probe.py all > gpurun_out/tc_probe_all.log 2>&1; echo "probe rc=$?"; tail -8 gpurun_out/tc_probe_all.log | cut -c1-400
timeout 600 python scripts/tc_probe.py bwd > gpurun_out/tc_probe_bwd.log 2>&1; echo "probe bwd rc=$?"; tail -8 gpurun_out/tc_probe_bwd.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_c3_quick.json
