"""Time imaging.VisMapper.make_map at C3 size (HERA-350 x nside-128 pixels x NF channels x 1
time): the adjoint of the RIME on the same kernels.  usage: python scripts/vismapper_time.py [NF]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayeslim_b200 as ba
import workloads
from bayeslim_b200 import ops
from bench import KernelTimer

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = 'cuda'
rime = workloads.pixel_interp(128, nf, 1, dev, torch.float32, sky_param=False, beam_param=False)
with torch.no_grad():
    vd = rime()
ra, dec = workloads.healpix_sky_angles(128)
out = {"workload": "VisMapper.make_map: HERA-350 (61075 bl) x %d pixels x %d freqs x 1 time" % (len(ra), nf)}
for method in ("w", "A2w"):
    vm = ba.imaging.VisMapper(vd, ra, dec, beam=rime.beam, dtype=torch.float32)
    key = list(rime.telescope.conv_cache.keys())[0]
    vm.telescope.conv_cache[vm.telescope.hash(float(vm.times[0]), ra)] = rime.telescope.conv_cache[key]
    vm.set_normalization(method)
    vm.make_map()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with KernelTimer(ops) as kt:
        e0.record()
        maps, P = vm.make_map()
        e1.record()
        k = kt.summary()
    ns = sum(vm._geometry()['geom'].ns)
    evals = ns * len(vm.bls) * nf
    ms = e0.elapsed_time(e1)
    out[method] = dict(ms=ms, evals=evals, evals_per_s=evals / (ms * 1e-3),
                       kernels={n: round(d["ms"], 2) for n, d in k.items()},
                       finite=bool(torch.isfinite(maps).all()))
print(json.dumps(out))
