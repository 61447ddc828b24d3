#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/vismapper_time.py 256 > gpurun_out/vismapper_time.json 2> gpurun_out/vismapper_time.err
echo "vismapper exit $?"; cat gpurun_out/vismapper_time.json; tail -n 3 gpurun_out/vismapper_time.err
