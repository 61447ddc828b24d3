#!/bin/bash
# parity of the antenna kernels + C3 timing with and without the antenna-position gradient
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "antenna or golden or small_c3" > gpurun_out/pytest_ant.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -n 3 gpurun_out/pytest_ant.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python scripts/ant_time.py main 256 > gpurun_out/ant_time.json 2> gpurun_out/ant_time.err; cat gpurun_out/ant_time.json
timeout 200 python - > gpurun_out/ant_time_noantpos.json 2> gpurun_out/ant_time_noantpos.err <<'PY'
import json, torch, workloads
from bayeslim_b200 import ops
from bench import KernelTimer
rime = workloads.pixel_interp(128, 256, 1, 'cuda', torch.float32, antpos_param=False)
def step():
    for p in rime.parameters():
        p.grad = None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
step(); torch.cuda.synchronize()
with KernelTimer(ops) as kt:
    step(); step()
    k = kt.summary()
print(json.dumps({n: round(d["ms"] / 2, 2) for n, d in k.items()}))
PY
cat gpurun_out/ant_time_noantpos.json; tail -n 3 gpurun_out/ant_time_noantpos.err
