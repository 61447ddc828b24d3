#!/bin/bash
# tensor-core kernel: probe (parity + timing), then launch profile of the same command
set -o pipefail
timeout 400 python scripts/tc_probe.py all > gpurun_out/tc_probe.log 2>&1; echo "probe rc=$?"
tail -12 gpurun_out/tc_probe.log
timeout 200 python scripts/tc_probe.py prof > gpurun_out/tc_prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fringe -c 1 \
    -o gpurun_out/prof_tc -f python scripts/tc_probe.py prof > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc.log
