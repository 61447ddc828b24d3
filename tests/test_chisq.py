"""
Fused likelihood epilogue (reference optim.py:959-1030, apply_icov
optim.py:1836): RIME.forward_chisq / optim.LogProb against the unfused evaluation
(forward() followed by the reference's residual / icov arithmetic in float64) on the golden
RIME fixtures -- chisq value and gradients to sky, beam and antenna positions.
"""
import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from tests import model_cases as mc
from tests.oracle_cases import load

CASES = [("rime_point_airy", mc.build_point_airy), ("rime_pixel_interp", mc.build_pixel_interp),
         ("rime_4pol", mc.build_4pol), ("rime_databls", mc.build_databls)]


def _target(rime, seed, device, cdtype):
    """Synthetic data = model visibilities + noise, with inverse-variance weights."""
    with torch.no_grad():
        vd = rime()
    g = torch.Generator().manual_seed(seed)
    noise = torch.complex(torch.randn(vd.data.shape, generator=g, dtype=torch.float64),
                          torch.randn(vd.data.shape, generator=g, dtype=torch.float64))
    scale = float(vd.data.abs().max()) * 0.3
    vd.data = (vd.data.to(torch.complex128).cpu() + scale * noise).to(device=device, dtype=cdtype)
    vd.icov = (torch.rand(vd.data.shape, generator=g, dtype=torch.float64) + 0.5).to(device) / scale ** 2
    vd.cov_axis = None
    return vd


def check_fused_chisq(name, build, device, dtype, tol):
    g = load(name)
    built = build(g, device, dtype)
    rime, leaves = built[0], built[1]
    cdtype = torch.complex128 if dtype == torch.float64 else torch.complex64
    target = _target(rime, 11, device, cdtype)
    leaves = {k: v for k, v in leaves.items() if getattr(v, 'requires_grad', False)}

    def grads():
        out = {k: v.grad.detach().double().cpu().clone() for k, v in leaves.items() if v.grad is not None}
        for v in leaves.values():
            v.grad = None
        return out

    # unfused: the reference's arithmetic on forward()'s visibilities
    unf = ba.optim.LogProb(rime, target, fuse=False)
    chi_u, res = unf.forward_chisq()
    assert res is not None
    chi_u.backward()
    g_u = grads()
    # fused
    fus = ba.optim.LogProb(rime, target, fuse=True)
    chi_f, res_f = fus.forward_chisq()
    chi_f.backward()
    g_f = grads()
    assert abs(float(chi_f.detach()) - float(chi_u.detach())) <= tol * abs(float(chi_u.detach()))
    assert set(g_f) == set(g_u) and len(g_f) > 0
    for k in g_u:
        err = float((g_f[k] - g_u[k]).abs().max() / g_u[k].abs().max())
        assert err < tol * 20, (name, k, err)
    if name != "rime_databls":
        assert res_f is None and fus.cotangent is not None          # the fused route ran
        G = fus.cotangent.to(torch.complex128).cpu()
        ref = 2 * target.icov.cpu() * res.detach().to(torch.complex128).cpu()
        assert float((G - ref).abs().max() / ref.abs().max()) < tol * 20
    # log-likelihood wrapper: same value through both routes
    with torch.no_grad():
        assert abs(float(fus.forward_like()) - float(unf.forward_like())) <= tol * abs(float(chi_u))
    return float(chi_u.detach())


@pytest.mark.parametrize("name,build", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_chisq_host_logic_with_emulated_kernels(name, build, dtype):
    from tests.cpu_double import emulated_kernels
    with emulated_kernels() as calls:
        check_fused_chisq(name, build, 'cpu', dtype, 1e-11 if dtype == torch.float64 else 2e-5)
    if name != "rime_databls":
        assert "reduce_units_chisq" in calls


def test_apply_icov_forms():
    rng = np.random.default_rng(1)
    d = torch.as_tensor(rng.normal(size=(6,)) + 1j * rng.normal(size=(6,)))
    w = torch.as_tensor(rng.uniform(1, 2, size=(6,)))
    full = torch.diag(w).to(torch.complex128)
    a = ba.optim.apply_icov(d, w, None).sum()
    b = ba.optim.apply_icov(d, full, 'full')
    assert abs(complex(a) - complex(b)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name,build", CASES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_chisq_cuda(name, build, dtype):
    check_fused_chisq(name, build, 'cuda', dtype, 1e-11 if dtype == torch.float64 else 2e-5)
