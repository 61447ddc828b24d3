#!/bin/bash
# parity of the antenna kernels, default bench, ncu launch list with DRAM bytes of the bench command
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna or golden or small_c3" > gpurun_out/pytest_ant.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -n 3 gpurun_out/pytest_ant.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-300 gpurun_out/bench_default.json
CMD="python bench.py --workload c3 --nt 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
