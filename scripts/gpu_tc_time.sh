#!/bin/bash
# quick parity + timing of the tensor-core kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python scripts/tc_probe.py all > gpurun_out/tc_probe_all.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/tc_probe_all.log | cut -c1-330
timeout 600 python scripts/tc_probe.py bwd > gpurun_out/tc_probe_bwd.log 2>&1; echo "probe bwd rc=$?"; tail -2 gpurun_out/tc_probe_bwd.log | cut -c1-330
