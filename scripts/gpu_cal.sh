#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_calibration.py tests/test_imaging.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_cal.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_cal.log
timeout 200 python scripts/cal_time.py > gpurun_out/cal_time.json 2> gpurun_out/cal_time.err
echo "cal_time exit $?"; cat gpurun_out/cal_time.json; tail -n 3 gpurun_out/cal_time.err
