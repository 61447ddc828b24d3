"""Small, fixed C3-shaped case for ncu: 8192 baselines x nside-128 sky x 1024 freqs x 1 time,
forward + backward (sky, beam, antenna positions), run twice (first pass warms caches/tables)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads

n_bl = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rime = workloads.pixel_interp(128, 1024, 1, 'cuda', torch.float32, n_bl=n_bl, antpos_param=True)
for it in range(2):
    for p in rime.parameters():
        p.grad = None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
    torch.cuda.synchronize()
print("ok", tuple(V.shape))
