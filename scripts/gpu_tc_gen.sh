#!/bin/bash
# forward tensor-core kernel: one read of the source vectors for both operand rows (default build)
# against the previous build (variants/lib_tcv_gen1.so): timing + parity
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/tc_gen.jsonl
for so in bayeslim_b200/csrc/variants/lib_tcv_gen1.so bayeslim_b200/csrc/libb200rime.so bayeslim_b200/csrc/variants/lib_tcv_gen1.so bayeslim_b200/csrc/libb200rime.so; do
  echo "== $so" >> gpurun_out/tc_gen.jsonl
  B200RIME_LIB=$PWD/$so timeout 200 python scripts/tc_probe.py time >> gpurun_out/tc_gen.jsonl 2>> gpurun_out/tc_gen.err
done
timeout 300 python scripts/tc_probe.py all >> gpurun_out/tc_gen.jsonl 2>> gpurun_out/tc_gen.err
tail -3 gpurun_out/tc_gen.err
