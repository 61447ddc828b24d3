#!/bin/bash
# Round-end evidence: full GPU test suite, smoke, default bench, ncu launch list of the bench
# command and ncu --set full captures of the dominant kernels on the bench workload.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-300 gpurun_out/bench_default.json
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cut -c1-200 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference exit $?"; cut -c1-200 gpurun_out/bench_reference.json
CMD="python bench.py --workload c3 --nt 1 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fringe_sum_fwd|build_interp_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_bench.log
