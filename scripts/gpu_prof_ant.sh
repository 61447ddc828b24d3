#!/bin/bash
# ncu --set full of one antenna-kernel launch: scripts/gpu_prof_ant.sh <fwd|bwd> <NF>
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
which=${1:-bwd}; nf=${2:-64}
timeout 300 python scripts/prof_ant.py $nf > gpurun_out/prof_ant_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ant_fringe_$which -s 1 -c 1 \
    -o gpurun_out/prof_ant_$which python scripts/prof_ant.py $nf > gpurun_out/ncu_ant_$which.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_ant_$which.log; tail -n 2 gpurun_out/prof_ant_plain.log
