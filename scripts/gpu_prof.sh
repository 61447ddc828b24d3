#!/bin/bash
# ncu capture of the fringe kernels on the fixed profiling case + microbench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 python - > gpurun_out/microbench.json 2> gpurun_out/microbench.err <<'PY'
import json
from bayeslim_b200 import _lib
out = dict(device=_lib.device_info(0))
for kind, it in (("fp32", 4096), ("fp32x2", 4096), ("rf3_fp32", 4096), ("rf3_fp32x2", 4096), ("mix_rot_mac", 4096), ("mix_mac", 4096), ("mix_rot", 4096), ("fp64", 1024), ("mufu", 2048)):
    g, ms = _lib.microbench(kind, it)
    out[kind] = dict(gops=g, ms=ms)
print(json.dumps(out))
PY
cat gpurun_out/microbench.json
python scripts/prof_case.py 8192 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fringe_sum -s 3 -c 3 \
    -o gpurun_out/prof_fringe python scripts/prof_case.py 8192 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full.log
