#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "golden or small_c3 or healpix or batched or c3_full or c4_reduced" > gpurun_out/pytest_interp.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_interp.log
timeout 300 python bench.py --workload c3 --nt 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_interp.json 2> gpurun_out/bench_interp.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_interp.json"))
print(d["value"], d["kernel_ms_per_step"]); print(d["roofline_hbm"])
PY
