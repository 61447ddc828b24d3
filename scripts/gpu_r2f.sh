#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden_cases and (alm_sky or ylm)" > gpurun_out/pytest_alm_sky.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_alm_sky.log | cut -c1-300
