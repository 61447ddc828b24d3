#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "graphed" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"; grep -n "Error\|passed\|failed" gpurun_out/pytest_graph.log | head -20 | cut -c1-300
