"""C3-shaped case for ncu on the antenna-factorised kernels: HERA-350 all pairs x nside-128 sky x
NF channels x 1 time, forward + backward (sky, beam, antenna positions), run twice."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rime = workloads.pixel_interp(128, nf, 1, 'cuda', torch.float32, antpos_param=True)
for it in range(2):
    for p in rime.parameters():
        p.grad = None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
    torch.cuda.synchronize()
print("ok", tuple(V.shape))
