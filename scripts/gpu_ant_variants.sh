#!/bin/bash
# parity + timing of every library variant under bayeslim_b200/csrc/variants/
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/ant_variants.jsonl
for so in bayeslim_b200/csrc/variants/lib_*.so; do
  B200RIME_LIB=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna_factorised" > gpurun_out/pytest_$(basename $so .so).log 2>&1
  rc=$?; echo "$so pytest exit $rc"; tail -n 2 gpurun_out/pytest_$(basename $so .so).log
  if [ $rc -ne 0 ]; then continue; fi
  B200RIME_LIB=$PWD/$so timeout 120 python scripts/ant_time.py $(basename $so .so) 256 >> gpurun_out/ant_variants.jsonl 2>> gpurun_out/ant_variants.err
  echo "$so time exit $?"
done
cat gpurun_out/ant_variants.jsonl
