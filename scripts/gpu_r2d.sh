#!/bin/bash
# fused cotangent pack kernel: tensor-core backward probes (parity + timing), full GPU suite, C3 / C4 lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/tc_probe.py bwd > gpurun_out/tc_probe_bwd.log 2>&1; echo "probe bwd rc=$?"; tail -7 gpurun_out/tc_probe_bwd.log | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err; echo "c3 rc=$?"; head -c 330 gpurun_out/bench_c3_quick.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; head -c 330 gpurun_out/bench_c4.json
