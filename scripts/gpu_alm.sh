#!/bin/bash
# spherical-harmonic GEMM: probe (parity + timing)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python scripts/alm_probe.py > gpurun_out/alm_probe.jsonl 2> gpurun_out/alm_probe.err; echo "probe rc=$?"
cut -c1-260 gpurun_out/alm_probe.jsonl | tail -40; tail -5 gpurun_out/alm_probe.err
