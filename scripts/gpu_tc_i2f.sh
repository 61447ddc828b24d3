#!/bin/bash
# tensor-core kernels with the integer phase-fraction conversion (no I2F on the XU pipe) against
# the I2F build: parity (all / bwd probes) and timing
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/tc_i2f.jsonl
for so in bayeslim_b200/csrc/variants/lib_tcv_i2f.so bayeslim_b200/csrc/libb200rime.so; do
  echo "== $so" >> gpurun_out/tc_i2f.jsonl
  B200RIME_LIB=$PWD/$so timeout 200 python scripts/tc_probe.py time >> gpurun_out/tc_i2f.jsonl 2>> gpurun_out/tc_i2f.err
  B200RIME_LIB=$PWD/$so timeout 300 python scripts/tc_probe.py bwd 2>> gpurun_out/tc_i2f.err | tail -2 >> gpurun_out/tc_i2f.jsonl
  echo "$so exit $?"
done
timeout 300 python scripts/tc_probe.py all >> gpurun_out/tc_i2f.jsonl 2>> gpurun_out/tc_i2f.err
cut -c1-700 gpurun_out/tc_i2f.jsonl; tail -3 gpurun_out/tc_i2f.err
