#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python scripts/prof_case.py 8192 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bwd_sky -s 1 -c 1 \
    -o gpurun_out/prof_sky python scripts/prof_case.py 8192 > gpurun_out/ncu_sky.log 2>&1
echo "ncu sky exit $?"; tail -n 3 gpurun_out/ncu_sky.log
python scripts/prof_case.py 8192 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sum_fwd -s 1 -c 1 \
    -o gpurun_out/prof_fwd python scripts/prof_case.py 8192 > gpurun_out/ncu_fwd.log 2>&1
echo "ncu fwd exit $?"; tail -n 3 gpurun_out/ncu_fwd.log
