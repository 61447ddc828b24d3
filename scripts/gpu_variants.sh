#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/variants.jsonl
for so in bayeslim_b200/csrc/variants/lib_*.so; do
  B200RIME_LIB=$PWD/$so timeout 300 python scripts/variant_time.py 8192 $(basename $so .so) >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
cat gpurun_out/variants.jsonl; tail -n 3 gpurun_out/variants.err
