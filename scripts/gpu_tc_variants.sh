#!/bin/bash
# timing (and parity) of the tensor-core kernels for every library variant against the shipped build
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/tc_variants.jsonl
for so in bayeslim_b200/csrc/libb200rime.so bayeslim_b200/csrc/variants/lib_tcv*.so bayeslim_b200/csrc/libb200rime.so; do
  echo "== $so" >> gpurun_out/tc_variants.jsonl
  B200RIME_LIB=$PWD/$so timeout 200 python scripts/tc_probe.py time >> gpurun_out/tc_variants.jsonl 2>> gpurun_out/tc_variants.err
  B200RIME_LIB=$PWD/$so timeout 300 python scripts/tc_probe.py bwd 2>> gpurun_out/tc_variants.err | tail -2 >> gpurun_out/tc_variants.jsonl
done
tail -2 gpurun_out/tc_variants.err
