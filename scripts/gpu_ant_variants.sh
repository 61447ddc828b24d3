#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna_factorised" > gpurun_out/pytest_ant0.log 2>&1
rc=$?; echo "pytest antenna exit $rc"; tail -n 5 gpurun_out/pytest_ant0.log
if [ $rc -ne 0 ]; then exit 1; fi
: > gpurun_out/ant_variants.jsonl
for so in bayeslim_b200/csrc/variants/lib_*.so; do
  B200RIME_LIB=$PWD/$so timeout 120 python scripts/ant_time.py $(basename $so .so) 256 >> gpurun_out/ant_variants.jsonl 2>> gpurun_out/ant_variants.err
  echo "$so exit $?"
done
cat gpurun_out/ant_variants.jsonl
