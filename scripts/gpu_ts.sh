#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== TS"; timeout 300 python scripts/tc_probe.py all 2>&1 | tail -6 | cut -c1-330
echo "== SS"; B200RIME_TC_TS=0 timeout 300 python scripts/tc_probe.py time 2>&1 | tail -1 | cut -c1-330
