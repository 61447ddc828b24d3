"""GPU probe of the spherical-harmonic complex GEMM (ops.alm_forward): parity against complex128
torch.matmul over ragged shapes, both operand kinds, both output kinds, split-k; timing at the
size of an nside-64 beam map x 1024 channels x lmax-60 modes."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from bayeslim_b200 import ops  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(0)


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


def case(M, C, P, pc, yc, ro, dtype=torch.float32, scale=1.0):
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    Y = torch.randn(C, P, dtype=cd if yc else dtype, device=dev) * scale
    p = torch.randn(M, C, dtype=cd if pc else dtype, device=dev)
    if not yc and not pc and not ro:
        return None
    p.requires_grad_()
    plan = ops.AlmPlan(Y)
    out = ops.alm_forward(p, plan, real_out=ro)
    G = torch.randn_like(out)
    out.backward(G)
    torch.cuda.synchronize()
    p2 = p.detach().to(torch.complex128 if pc else torch.float64).requires_grad_()
    Y2 = Y.to(torch.complex128 if yc else torch.float64)
    ct = torch.complex128 if (pc or yc) else torch.float64
    ref = p2.to(ct) @ Y2.to(ct)
    if ro:
        ref = ref.real
    if ref.is_complex() != G.is_complex():
        G2 = G.to(ref.dtype)
    else:
        G2 = G.to(ref.dtype)
    ref.backward(G2)
    return dict(M=M, C=C, P=P, p_complex=pc, y_complex=yc, real_out=ro, dtype=str(dtype),
                fwd=relmax(out.detach(), ref.detach()), adj=relmax(p.grad, p2.grad))


PROF = len(sys.argv) > 1 and sys.argv[1] == 'prof'
for dtype in (() if PROF else (torch.float32, torch.float64)):
    for (M, C, P) in [(5, 37, 300), (130, 20, 129), (64, 528, 3000), (257, 100, 1000)]:
        for pc, yc, ro in [(True, True, True), (True, True, False), (False, True, True), (True, False, False)]:
            r = case(M, C, P, pc, yc, ro, dtype)
            if r:
                print(json.dumps(r), flush=True)
if not PROF:
    print(json.dumps(case(40, 90, 5000, True, True, True, scale=1e-6)), flush=True)

# timing: 1024 channel rows x 1891 modes (lmax 60, m >= 0) x 49152 pixels (nside 64)
M, C, P = 1024, 1891, 49152
Y = torch.randn(C, P, dtype=torch.complex64, device=dev)
p = torch.randn(M, C, dtype=torch.complex64, device=dev, requires_grad=True)
plan = ops.AlmPlan(Y)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(3):
    p.grad = None
    ev[0].record()
    out = ops.alm_forward(p, plan, real_out=True)
    ev[1].record()
    G = torch.ones_like(out)
    ev[2].record()
    out.backward(G)
    ev[3].record()
    torch.cuda.synchronize()
fwd_ms, adj_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
p64 = p.detach()[:8].to(torch.complex128)
ref = (p64 @ Y.to(torch.complex128)).real
err = relmax(out.detach()[:8].double(), ref)
t0 = time.time()
for _ in range(3):
    ev[0].record()
    o2 = (p.detach() @ Y)
    ev[1].record()
    torch.cuda.synchronize()
lib_ms = ev[0].elapsed_time(ev[1])
flop = 8.0 * M * C * P
print(json.dumps(dict(kind="timing", M=M, C=C, P=P, fwd_ms=fwd_ms, adj_ms=adj_ms,
                      fwd_algorithmic_tflops=flop / fwd_ms / 1e9,
                      fwd_executed_tflops=3 * flop / fwd_ms / 1e9,
                      adj_executed_tflops=0.5 * 3 * flop / adj_ms / 1e9,
                      cublas_cgemm_ms=lib_ms, relmax_rows0_8=err)), flush=True)
