#!/bin/bash
# One GPU-box session: peaks, parity tests, smoke, benches.  Everything lands in gpurun_out/.
# usage: gpurun --timeout 2400 -- bash scripts/gpu_round.sh [quick]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 120 python - > gpurun_out/microbench.json 2> gpurun_out/microbench.err <<'PY'
import json
from bayeslim_b200 import _lib
out = dict(device=_lib.device_info(0))
for kind, it in (("fp32", 4096), ("fp32x2", 4096), ("rf3_fp32", 4096), ("rf3_fp32x2", 4096), ("mix_rot_mac", 4096), ("mix_mac", 4096), ("mix_rot", 4096), ("fp64", 1024), ("mufu", 2048)):
    g, ms = _lib.microbench(kind, it)
    out[kind] = dict(gops=g, ms=ms)
print(json.dumps(out))
PY
echo "microbench exit $?"; cat gpurun_out/microbench.json
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
if [ "$1" != "quick" ]; then
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cat gpurun_out/bench_c2.json; tail -n 5 gpurun_out/bench_c2.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "bench c1 exit $?"; cut -c1-300 gpurun_out/bench_c1.json; tail -n 3 gpurun_out/bench_c1.err
timeout 1500 python bench.py --workload c3 --nt 1 --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 exit $?"; cat gpurun_out/bench_c3.json; tail -n 5 gpurun_out/bench_c3.err
fi
