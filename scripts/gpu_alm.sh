#!/bin/bash
# spherical-harmonic GEMM: GPU tests (test_alm.py + the rime_ylm golden cases), probe (parity + timing)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_alm.py tests/test_gpu_parity.py tests/test_eq2top.py -m gpu -q -k "alm or cgemm or ylm or eq2top" > gpurun_out/pytest_alm.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_alm.log
timeout 300 python scripts/alm_probe.py > gpurun_out/alm_probe.jsonl 2> gpurun_out/alm_probe.err; echo "probe rc=$?"
tail -1 gpurun_out/alm_probe.jsonl; tail -5 gpurun_out/alm_probe.err
