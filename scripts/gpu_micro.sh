#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python - > gpurun_out/microbench.json 2> gpurun_out/microbench.err <<'PY'
import json
from bayeslim_b200 import _lib
out = dict(device=_lib.device_info(0))
for kind, it in (("fp32", 4096), ("fp32x2", 4096), ("rf3_fp32", 4096), ("rf3_fp32x2", 4096), ("mix_rot_mac", 4096), ("mix_mac", 4096), ("mix_rot", 4096), ("rot_1ch", 2048), ("rot_4ch", 4096), ("rot_8ch", 4096), ("rotmac_4ch", 4096), ("rotmac_8ch", 4096), ("fp64", 1024), ("mufu", 2048)):
    g, ms = _lib.microbench(kind, it)
    out[kind] = dict(gops=g, ms=ms)
print(json.dumps(out))
PY
cat gpurun_out/microbench.json; cat gpurun_out/microbench.err | tail -3
