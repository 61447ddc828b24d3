#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
python scripts/prof_case.py 8192 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fringe_sum -s 3 -c 3 \
    -o gpurun_out/prof_fringe python scripts/prof_case.py 8192 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 5 gpurun_out/ncu_full.log
python scripts/prof_case.py 8192 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python scripts/prof_case.py 8192 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"; tail -n 3 gpurun_out/ncu_list.log
