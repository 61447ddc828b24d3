"""
calibration.apply_cal: oracle and package against golden vectors
from the unmodified reference (tests/golden/apply_cal.npz, made by tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from oracle import rime_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
DOUBLE = os.environ.get("B200RIME_TEST_DOUBLE") == "1"
TAGS = [("1pol", False), ("2pol", True), ("4pol", False), ("1pol_bcast", False), ("4pol_bcast", False)]


def load():
    return dict(np.load(os.path.join(HERE, "golden", "apply_cal.npz")))


def relmax(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    return float((a.to(b.dtype) - b).abs().max() / b.abs().max())


def _idx(g):
    ants = [int(a) for a in g["ants"]]
    bls = [tuple(int(x) for x in b) for b in g["bls"]]
    g1 = torch.as_tensor([ants.index(b[0]) for b in bls])
    g2 = torch.as_tensor([ants.index(b[1]) for b in bls])
    return ants, bls, g1, g2


@pytest.mark.parametrize("tag,cal_2pol", TAGS)
def test_oracle_apply_cal_matches_reference(tag, cal_2pol):
    g = load()
    _, _, g1, g2 = _idx(g)
    vis = torch.as_tensor(g[tag + "_vis"]).requires_grad_(True)
    gains = torch.as_tensor(g[tag + "_gains"]).requires_grad_(True)
    cov = torch.as_tensor(g[tag + "_cov"]) if tag + "_cov" in g else None
    vout, cov_out = orc.apply_cal(vis, gains, g1, g2, cal_2pol=cal_2pol, cov=cov)
    assert relmax(vout, g[tag + "_out"]) < 1e-14
    G = torch.as_tensor(g[tag + "_G"])
    (G.real * vout.real + G.imag * vout.imag).sum().backward()
    assert relmax(vis.grad, g[tag + "_dvis"]) < 1e-13
    assert relmax(gains.grad, g[tag + "_dgains"]) < 1e-13
    if cov is not None:
        assert relmax(cov_out, g[tag + "_cov_out"]) < 1e-14
    if tag + "_undo" in g:
        vu, _ = orc.apply_cal(vis.detach(), gains.detach(), g1, g2, cal_2pol=cal_2pol, undo=True)
        assert relmax(vu, g[tag + "_undo"]) < 1e-13


def check_package(device, cdtype, tol):
    g = load()
    ants, bls, _, _ = _idx(g)
    for tag, cal_2pol in TAGS:
        vis = torch.as_tensor(g[tag + "_vis"]).to(device=device, dtype=cdtype).requires_grad_(True)
        gains = torch.as_tensor(g[tag + "_gains"]).to(device=device, dtype=cdtype).requires_grad_(True)
        cov = None
        if tag + "_cov" in g:
            cov = torch.as_tensor(g[tag + "_cov"]).to(device=device, dtype=vis.real.dtype)
        vout, cov_out = ba.calibration.apply_cal(vis, bls, gains, ants, cal_2pol=cal_2pol, cov=cov)
        assert relmax(vout, g[tag + "_out"]) < tol, tag
        G = torch.as_tensor(g[tag + "_G"]).to(device=device, dtype=cdtype)
        (G.real * vout.real + G.imag * vout.imag).sum().backward()
        assert relmax(vis.grad, g[tag + "_dvis"]) < tol, tag
        assert gains.grad.shape == gains.shape
        assert relmax(gains.grad, g[tag + "_dgains"]) < tol, tag
        if cov is not None:
            assert relmax(cov_out, g[tag + "_cov_out"]) < tol, tag
        with torch.no_grad():
            vu, _ = ba.calibration.apply_cal(vis, bls, gains, ants, cal_2pol=cal_2pol, undo=True)
            if tag + "_undo" in g:
                assert relmax(vu, g[tag + "_undo"]) < tol * 5, tag
            # undo inverts apply (4pol included: 2x2 inverse per antenna, where the reference's
            # torch.pinv call does not run)
            back, _ = ba.calibration.apply_cal(vout.detach(), bls, gains, ants, cal_2pol=cal_2pol,
                                               undo=True)
            ref = vis.detach().clone()
            if cal_2pol:
                ref[0, 1] = 0
                ref[1, 0] = 0
            assert relmax(back, ref) < tol * 50, tag


@pytest.mark.parametrize("cdtype", [torch.complex128, torch.complex64])
def test_apply_cal_host_logic_with_emulated_kernels(cdtype):
    from tests.cpu_double import emulated_kernels
    with emulated_kernels() as calls:
        check_package('cpu', cdtype, 1e-12 if cdtype == torch.complex128 else 2e-6)
    assert "apply_cal" in calls and "apply_cal_bwd_gains" in calls


def test_apply_cal_needs_cuda():
    g = load()
    ants, bls, _, _ = _idx(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        ba.calibration.apply_cal(torch.as_tensor(g["1pol_vis"]), bls, torch.as_tensor(g["1pol_gains"]),
                                 ants)


@pytest.mark.gpu
@pytest.mark.parametrize("cdtype", [torch.complex128, torch.complex64])
def test_apply_cal_cuda_matches_reference(cdtype):
    if DOUBLE:
        pytest.skip("covered by test_apply_cal_host_logic_with_emulated_kernels")
    check_package('cuda', cdtype, 1e-12 if cdtype == torch.complex128 else 2e-6)
