#!/bin/bash
# final build: forward-only C3 line and the C4 line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python bench.py --pass fwd --no-cpu-baseline > gpurun_out/bench_c3_fwd.json 2> gpurun_out/bench_c3_fwd.err; echo "c3 fwd rc=$?"; head -c 250 gpurun_out/bench_c3_fwd.json
timeout 300 python bench.py --workload c4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; head -c 250 gpurun_out/bench_c4.json
