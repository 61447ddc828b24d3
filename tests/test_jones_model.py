"""
calibration.JonesModel.forward (reference calibration.py:599-664):
oracle and package against golden vectors from the unmodified reference
(tests/golden/jones_model.npz, made by tests/golden/make_golden.py jones_model).
"""
import os

import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from oracle import rime_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("1pol_com_refant", '1pol', 'com'), ("1pol_com_tbatch", '1pol', 'com'),
         ("2pol_ampphs", '2pol', 'amp_phs'), ("1pol_dly", '1pol', 'dly'), ("4pol_com", '4pol', 'com')]


def load():
    return dict(np.load(os.path.join(HERE, "golden", "jones_model.npz")))


def relmax(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    return float((a.to(b.dtype) - b).abs().max() / b.abs().max())


def _meta(g):
    ants = [int(a) for a in g["ants"]]
    bls = [tuple(int(x) for x in b) for b in g["bls"]]
    return ants, bls, torch.as_tensor(g["freqs"]), g["times"]


@pytest.mark.parametrize("tag,polmode,ptype", CASES)
def test_oracle_jones_matches_reference(tag, polmode, ptype):
    g = load()
    ants, bls, freqs, times = _meta(g)
    refant = int(g[tag + "_refant"])
    ridx = ants.index(refant) if refant >= 0 else None
    p_in = torch.as_tensor(g[tag + "_params_in"])
    with torch.no_grad():
        p_fixed, _ = orc.jones_gains(p_in, ptype, freqs, ridx)
    # the reference rephases its parameter in place before the forward pass: gradients are with
    # respect to the rephased tensor
    assert relmax(p_fixed, g[tag + "_params"]) < 1e-14
    p = p_fixed.clone().requires_grad_(True)
    _, gains = orc.jones_gains(p, ptype, freqs, None)
    tsel = torch.as_tensor(g[tag + "_tsel"])
    vis = torch.as_tensor(g[tag + "_vis"]).requires_grad_(True)
    g1 = torch.as_tensor([ants.index(b[0]) for b in bls])
    g2 = torch.as_tensor([ants.index(b[1]) for b in bls])
    vout, _ = orc.apply_cal(vis, gains[..., tsel, :], g1, g2, cal_2pol=polmode == '2pol')
    assert relmax(vout, g[tag + "_out"]) < 1e-13
    G = torch.as_tensor(g[tag + "_G"])
    (G.real * vout.real + G.imag * vout.imag).sum().backward()
    assert relmax(vis.grad, g[tag + "_dvis"]) < 1e-12
    assert relmax(p.grad, g[tag + "_dparams"]) < 1e-12


def check_package(device, cdtype, tol):
    g = load()
    ants, bls, freqs, times = _meta(g)
    rdtype = torch.float64 if cdtype == torch.complex128 else torch.float32
    for tag, polmode, ptype in CASES:
        refant = int(g[tag + "_refant"])
        p_in = torch.as_tensor(g[tag + "_params_in"])
        p_in = p_in.to(device=device, dtype=cdtype if p_in.is_complex() else rdtype)
        R = ba.calibration.JonesResponse(param_type=ptype, freqs=freqs.to(device), device=device,
                                         times=torch.as_tensor(times))
        J = ba.calibration.JonesModel(p_in.clone(), ants, refant=refant if refant >= 0 else None,
                                      R=R, polmode=polmode)
        assert relmax(J.params, g[tag + "_params"]) < tol, tag
        tsel = g[tag + "_tsel"]
        vd = ba.dataset.VisData()
        vis = torch.as_tensor(g[tag + "_vis"]).to(device=device, dtype=cdtype).requires_grad_(True)
        vd.setup_data(bls, times[tsel], freqs, pol='ee' if polmode == '1pol' else None, data=vis)
        vout = J(vd)
        assert isinstance(vout, ba.dataset.VisData) and vout is not vd and vd.data is vis
        assert relmax(vout.data, g[tag + "_out"]) < tol, tag
        G = torch.as_tensor(g[tag + "_G"]).to(device=device, dtype=cdtype)
        (G.real * vout.data.real + G.imag * vout.data.imag).sum().backward()
        assert relmax(vis.grad, g[tag + "_dvis"]) < tol, tag
        assert relmax(J.params.grad, g[tag + "_dparams"]) < tol * 5, tag
        # undo inverts the forward product
        with torch.no_grad():
            back = J(vout, undo=True)
            ref = vis.detach().clone()
            if polmode == '2pol':
                ref[0, 1] = 0
                ref[1, 0] = 0
            assert relmax(back.data, ref) < tol * 50, tag
        # cached antenna rows and time rows
        assert len(J.cache_aidx) == 1 and len(J.cache_tidx) == 1


@pytest.mark.parametrize("cdtype", [torch.complex128, torch.complex64])
def test_jones_model_host_logic_with_emulated_kernels(cdtype):
    from tests.cpu_double import emulated_kernels
    with emulated_kernels() as calls:
        check_package('cpu', cdtype, 1e-12 if cdtype == torch.complex128 else 3e-6)
    assert "apply_cal" in calls and "apply_cal_bwd_gains" in calls


def test_rephase_modes():
    rng = np.random.default_rng(0)
    p = torch.as_tensor(rng.normal(size=(1, 1, 4, 2, 3)) + 1j * rng.normal(size=(1, 1, 4, 2, 3)))
    q, _ = ba.calibration.rephase_to_refant(p, 'com', 1)
    assert float(q[:, :, 1].imag.abs().max()) < 1e-15
    assert relmax(q.abs(), p.abs()) < 1e-15
    z, _ = ba.calibration.rephase_to_refant(p, 'com', 1, mode='zero')
    assert float(z[:, :, 1].imag.abs().max()) == 0 and torch.equal(z[:, :, 0], p[:, :, 0])
    # 2-real view of complex parameters
    pr = torch.view_as_real(p).clone()
    qr, _ = ba.calibration.rephase_to_refant(pr, 'com', 1)
    assert relmax(torch.view_as_complex(qr), q) < 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("cdtype", [torch.complex128, torch.complex64])
def test_jones_model_cuda_matches_reference(cdtype):
    check_package('cuda', cdtype, 1e-12 if cdtype == torch.complex128 else 3e-6)
