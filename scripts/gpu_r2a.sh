#!/bin/bash
# round-2 evidence, part A: GPU tests, C4 / C5 bench lines at N=1, ncu launch list of the default
# bench command, ncu --set full of the two tensor-core kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/nvsmi.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { # name, args...
  n=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  echo "bench $n rc=$? $(head -c 400 gpurun_out/bench_$n.json)"
}
run c4 --workload c4 --steps 2 --warmup 3 --no-cpu-baseline
run c5 --workload c5 --steps 2 --warmup 3 --no-cpu-baseline
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python scripts/tc_probe.py prof > gpurun_out/tc_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fringe_fwd -c 1 \
    -o gpurun_out/r02_tc_fwd -f python scripts/tc_probe.py prof > gpurun_out/ncu_tc_fwd.log 2>&1
echo "ncu fwd rc=$?"; tail -2 gpurun_out/ncu_tc_fwd.log
timeout 300 python scripts/tc_probe.py profbwd > gpurun_out/tc_profbwd_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fringe_bwd -c 1 \
    -o gpurun_out/r02_tc_bwd -f python scripts/tc_probe.py profbwd > gpurun_out/ncu_tc_bwd.log 2>&1
echo "ncu bwd rc=$?"; tail -2 gpurun_out/ncu_tc_bwd.log
ls -la gpurun_out
