"""Time the fringe kernels of one library variant on the fixed profiling case.
usage: B200RIME_LIB=<so> python scripts/variant_time.py [n_bl] [tag]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from bayeslim_b200 import ops, _lib
from bench import KernelTimer

n_bl = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
tag = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(_lib.LIB_PATH)
rime = workloads.pixel_interp(128, 1024, 1, 'cuda', torch.float32, n_bl=n_bl, antpos_param=True)

def step():
    for p in rime.parameters():
        p.grad = None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()

step(); step()
torch.cuda.synchronize()
with KernelTimer(ops) as kt:
    step(); step()
    k = kt.summary()
evals = workloads.count_evals(rime)
peak = 2 * 128 * 148 * 1.965e9 / 1e12
out = {"tag": tag, "kc": _lib.KC["f32"]}
for name, fl in (("fringe_sum_fwd", 10), ("fringe_sum_bwd_sky", 10), ("fringe_sum_bwd_bl", 12)):
    ms = k[name]["ms"] / 2
    out[name] = {"ms": round(ms, 2), "frac_theory": round(evals * fl / (ms * 1e-3) / 1e12 / peak, 4)}
print(json.dumps(out))
