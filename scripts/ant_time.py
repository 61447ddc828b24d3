"""Time the antenna-factorised kernels of one library variant on C3 (1 time, NF channels).
usage: B200RIME_LIB=<so> python scripts/ant_time.py [tag] [NF]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from bayeslim_b200 import ops, _lib
from bench import KernelTimer

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(_lib.LIB_PATH)
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rime = workloads.pixel_interp(128, nf, 1, 'cuda', torch.float32, antpos_param=True)

def step():
    for p in rime.parameters():
        p.grad = None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()

step()
torch.cuda.synchronize()
with KernelTimer(ops) as kt:
    step(); step()
    k = kt.summary()
evals = workloads.count_evals(rime)
peak = 2 * 128 * 148 * 1.965e9 / 1e12
out = {"tag": tag}
for name, fl in (("antfringe_fwd", 10), ("antfringe_bwd", 22)):
    ms = k[name]["ms"] / 2
    out[name] = {"ms": round(ms, 2), "alg_frac_theory": round(evals * fl / (ms * 1e-3) / 1e12 / peak, 4)}
print(json.dumps(out))
