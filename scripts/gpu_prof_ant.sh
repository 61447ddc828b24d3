#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python scripts/prof_ant.py 64 > gpurun_out/prof_ant_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ant_fringe -s 2 -c 2 \
    -o gpurun_out/prof_ant python scripts/prof_ant.py 64 > gpurun_out/ncu_ant.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_ant.log; tail -n 2 gpurun_out/prof_ant_plain.log
