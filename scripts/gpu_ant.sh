#!/bin/bash
# GPU session for the antenna-factorised kernels: parity first (short leash), then A/B timing on
# C3 (1 time).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna_factorised" > gpurun_out/pytest_ant0.log 2>&1
rc=$?; echo "pytest antenna exit $rc"; tail -n 15 gpurun_out/pytest_ant0.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna or golden or small_c3" > gpurun_out/pytest_ant.log 2>&1
echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_ant.log
for flag in 1 0; do
  if [ "$flag" = "0" ] && [ "$1" = "skip0" ]; then continue; fi
  B200RIME_ANT=$flag timeout 300 python bench.py --workload c3 --nt 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_ant$flag.json 2> gpurun_out/bench_c3_ant$flag.err
  echo "bench ant=$flag exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c3_ant$flag.json"))
    print(d["value"], d["ms_per_step"], json.dumps(d["kernel_ms_per_step"]))
except Exception as e:
    print("no json", e)
PY
  tail -n 3 gpurun_out/bench_c3_ant$flag.err
done
