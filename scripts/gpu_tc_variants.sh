#!/bin/bash
# timing (and parity) of the tensor-core forward kernel for every library variant
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/tc_variants.jsonl
for so in bayeslim_b200/csrc/variants/lib_tcv*.so; do
  echo "== $so" >> gpurun_out/tc_variants.jsonl
  B200RIME_LIB=$PWD/$so timeout 200 python scripts/tc_probe.py time >> gpurun_out/tc_variants.jsonl 2>> gpurun_out/tc_variants.err
  echo "$so exit $?"
done
cat gpurun_out/tc_variants.jsonl
