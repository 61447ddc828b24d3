#!/bin/bash
# after the eq2top kernel / cache / brute-force commits: full GPU suite, smoke, C1 / C2 lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"; head -c 400 gpurun_out/bench_c1.json
timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; head -c 400 gpurun_out/bench_c2.json
