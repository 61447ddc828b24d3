#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
tag=${1:-x}
timeout 300 python scripts/tc_probe.py prof > gpurun_out/tc_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fringe_fwd -c 1 \
    -o gpurun_out/tc_fwd_$tag -f python scripts/tc_probe.py prof > gpurun_out/ncu_tc_fwd.log 2>&1
echo "ncu fwd rc=$?"; tail -2 gpurun_out/ncu_tc_fwd.log
