"""Time calibration.apply_cal (forward, adjoint to vis, adjoint to gains) at C3 size:
61075 baselines x 2 times x 1024 channels, 1-pol complex64; HBM roofline numbers."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayeslim_b200 as ba
import workloads
from bayeslim_b200 import ops

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                   "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {}
dev = 'cuda'
ants, _ = workloads.hera350()
bls = workloads.all_cross_bls(ants)
nt, nf = 2, 1024
g = torch.Generator(device='cpu').manual_seed(0)
vis = torch.randn(1, 1, len(bls), nt, nf, 2, generator=g).to(dev)
vis = torch.view_as_complex(vis).requires_grad_(True)
gains = torch.view_as_complex(torch.randn(1, 1, len(ants), nt, nf, 2, generator=g).to(dev)).requires_grad_(True)
G = torch.view_as_complex(torch.randn(1, 1, len(bls), nt, nf, 2, generator=g).to(dev))


class Timer:
    def __init__(self):
        self.rec, self.orig = [], ops._call

    def __enter__(self):
        def timed(name, sfx, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.orig(name, sfx, *args)
            e1.record()
            self.rec.append((name, e0, e1))
        ops._call = timed
        return self

    def __exit__(self, *a):
        ops._call = self.orig


def step():
    vis.grad = gains.grad = None
    out, _ = ba.calibration.apply_cal(vis, bls, gains, ants)
    (G.real * out.real + G.imag * out.imag).sum().backward()
    return out


for _ in range(3):
    step()
torch.cuda.synchronize()
with Timer() as tm:
    for _ in range(5):
        step()
torch.cuda.synchronize()
n = len(bls) * nt * nf
ms = {}
for name, e0, e1 in tm.rec:
    ms.setdefault(name, []).append(e0.elapsed_time(e1))
order = ["apply_cal (forward)", "apply_cal (adjoint to vis)", "apply_cal_bwd_gains"]
times = {"apply_cal (forward)": ms["apply_cal"][0::2], "apply_cal (adjoint to vis)": ms["apply_cal"][1::2],
         "apply_cal_bwd_gains": ms["apply_cal_bwd_gains"]}
alg = {"apply_cal (forward)": 16 * n, "apply_cal (adjoint to vis)": 16 * n, "apply_cal_bwd_gains": 16 * n}
out = {"workload": "apply_cal 1pol complex64, %d baselines x %d times x %d freqs" % (len(bls), nt, nf),
       "hbm_peak_gbs": peak.get("hbm_gbs")}
for k in order:
    t = sorted(times[k])[len(times[k]) // 2]
    out[k] = dict(ms=t, algorithmic_bytes=alg[k], gbs=alg[k] / (t * 1e-3) / 1e9)
# the same with plain torch ops (what the reference does), for scale
def ref_step():
    g1 = torch.as_tensor([ants.index(b[0]) for b in bls], device=dev)
    g2 = torch.as_tensor([ants.index(b[1]) for b in bls], device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = gains.detach().index_select(2, g1) * gains.detach().index_select(2, g2).conj() * vis.detach()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
ref_step()
out["torch eager forward (index_select, conj, 2 multiplies)"] = dict(ms=min(ref_step() for _ in range(3)))
print(json.dumps(out))
