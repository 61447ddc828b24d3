"""
imaging.VisMapper: oracle and package against golden vectors from
the unmodified reference (tests/golden/vismapper.npz, made by tests/golden/make_golden.py).

CPU part: the oracle restatement against the golden vectors, and the package's host logic with
the torch test double of the kernels.  GPU part (-m gpu): the CUDA adjoint / forward kernels.
"""
import os

import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from bayeslim_b200 import ops
from oracle import rime_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
LOC = (21.42827, -30.72148, 1051.7)
DOUBLE = os.environ.get("B200RIME_TEST_DOUBLE") == "1"


def load():
    return dict(np.load(os.path.join(HERE, "golden", "vismapper.npz")))


def relmax(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    return float((a.to(b.dtype) - b).abs().max() / b.abs().max())


def _oracle_inputs(g):
    freqs = torch.as_tensor(g["freqs"])
    ants = [int(a) for a in g["ants"]]
    idx = {a: k for k, a in enumerate(ants)}
    vecs = torch.as_tensor(g["antvecs"])
    bls = [tuple(int(x) for x in b) for b in g["bls"]]
    blvecs = torch.stack([vecs[idx[b[1]]] - vecs[idx[b[0]]] for b in bls])
    zenaz = [(za[0], za[1]) for za in g["zen_az"]]
    p = torch.as_tensor(g["beam_params"])
    beam_fn = lambda z, a: orc.airy_response(p, z, a, freqs, powerbeam=True)[0, 0, 0]
    return freqs, blvecs, zenaz, beam_fn


@pytest.mark.parametrize("method", ["w", "Aw", "A2w"])
@pytest.mark.parametrize("weighted", [True, False])
def test_oracle_make_map_matches_reference(method, weighted):
    g = load()
    freqs, blvecs, zenaz, beam_fn = _oracle_inputs(g)
    v = torch.as_tensor(g["vis"])[0, 0]
    w = torch.as_tensor(g["icov"])[0, 0] if weighted else None
    maps, P, D = orc.vismapper_make_map(v, w, blvecs, zenaz, freqs, len(g["ra"]), beam_fn,
                                        fov=float(g["fov"]), method=method)
    tag = "%s_%s" % (method, "icov" if weighted else "ones")
    assert relmax(maps, g["maps_" + tag]) < 1e-12
    assert relmax(P, g["P_" + tag]) < 1e-12
    assert relmax(D, g["D_" + tag]) < 1e-12


def test_oracle_nobeam_and_Am_match_reference():
    g = load()
    freqs, blvecs, zenaz, beam_fn = _oracle_inputs(g)
    v, w = torch.as_tensor(g["vis"])[0, 0], torch.as_tensor(g["icov"])[0, 0]
    maps, P, _ = orc.vismapper_make_map(v, w, blvecs, zenaz, freqs, len(g["ra"]), None,
                                        fov=float(g["fov_nobeam"]), method='A2w')
    assert relmax(maps, g["maps_nobeam"]) < 1e-12 and relmax(P, g["P_nobeam"]) < 1e-12
    Am = orc.vismapper_compute_Am(torch.as_tensor(g["test_maps"]), blvecs, zenaz, freqs, beam_fn,
                                  fov=float(g["fov"]))
    assert relmax(Am, g["Am"]) < 1e-12


# ----------------------------------------------------------------------------- package
def build_mapper(g, device, dtype, weighted=True, beam=True):
    freqs = torch.as_tensor(g["freqs"], dtype=torch.float64, device=device)
    ants = [int(a) for a in g["ants"]]
    antpos = ba.utils.AntposDict(ants, torch.as_tensor(g["antvecs"], dtype=torch.float64,
                                                        device=device))
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    vd = ba.dataset.VisData()
    vd.setup_meta(telescope=tel, antpos=antpos)
    bls = [tuple(int(x) for x in b) for b in g["bls"]]
    icov = torch.as_tensor(g["icov"], device=device).to(dtype) if weighted else None
    vd.setup_data(bls, g["times"], freqs, pol='ee',
                  data=torch.as_tensor(g["vis"], device=device).to(cdt), icov=icov)
    pb = None
    if beam:
        pb = ba.beam_model.PixelBeam(torch.as_tensor(g["beam_params"], device=device).to(dtype),
                                     freqs, R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                     powerbeam=True, fov=float(g["fov"]), parameter=False)
    vm = ba.imaging.VisMapper(vd, g["ra"], g["dec"], beam=pb, fov=float(g["fov_nobeam"]),
                              dtype=dtype)
    for t, za in zip(vm.times, g["zen_az"]):
        tel.conv_cache[tel.hash(float(t), g["ra"])] = torch.as_tensor(za, dtype=torch.float64,
                                                                      device=device)
    return vm


def check_mapper(device, dtype, tol):
    g = load()
    for weighted in (True, False):
        vm = build_mapper(g, device, dtype, weighted)
        for method in ("w", "Aw", "A2w"):
            vm.set_normalization(method)
            maps, P = vm.make_map(return_P=True, contract='diag')
            tag = "%s_%s" % (method, "icov" if weighted else "ones")
            assert tuple(maps.shape) == g["maps_" + tag].shape
            assert relmax(maps, g["maps_" + tag]) < tol, tag
            assert relmax(P, g["P_" + tag]) < tol, tag
            # D = 1 / clip(sum w Re(A^2)) is ill-conditioned where the sum nearly cancels (the
            # reference's A2w normalisation is not positive definite): float32 is judged on 1 / D
            if dtype == torch.float32:
                assert relmax(1 / vm.D, 1 / torch.as_tensor(g["D_" + tag])) < tol, tag
            else:
                assert relmax(vm.D, g["D_" + tag]) < tol, tag
    vm = build_mapper(g, device, dtype, True, beam=False)
    vm.set_normalization('A2w')
    maps, P = vm.make_map()
    assert relmax(maps, g["maps_nobeam"]) < tol and relmax(P, g["P_nobeam"]) < tol
    vm = build_mapper(g, device, dtype, True)
    Am = vm.compute_Am(torch.as_tensor(g["test_maps"], device=device).to(dtype))
    assert tuple(Am.shape) == g["Am"].shape and relmax(Am, g["Am"]) < tol
    vm.set_normalization('A2w')
    vm.make_map(return_P=False)
    Pm = vm.compute_Pm(torch.as_tensor(g["test_maps"], device=device).to(dtype), D=vm.D)
    assert relmax(Pm, g["Pm"]) < tol * 3


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_mapper_host_logic_with_emulated_kernels(dtype, monkeypatch):
    from tests.cpu_double import emulated_kernels
    with emulated_kernels() as calls:
        check_mapper('cpu', dtype, 1e-11 if dtype == torch.float64 else 2e-5)
    assert "fringe_sum_bwd_sky" in calls and "fringe_sum_fwd" in calls and "unpack" in calls
    if dtype == torch.float32:
        # the antenna-factorised adjoint, one triangle of the cotangent matrix
        monkeypatch.setattr(ops, "ANT_FWD_MIN_FILL", 0.0)
        monkeypatch.setattr(ops, "ANT_BWD_MIN_FILL", 0.0)
        with emulated_kernels() as calls:
            check_mapper('cpu', dtype, 2e-5)
        assert "antfringe_bwd" in calls and "fringe_sum_bwd_sky" not in calls


def test_mapper_selections_match_oracle_on_the_subset():
    """set_freq_inds / set_time_inds / set_bl_inds (imaging.py:100-226): mapping a subset equals
    the oracle run on the subset."""
    from tests.cpu_double import emulated_kernels
    g = load()
    freqs, blvecs, zenaz, beam_fn = _oracle_inputs(g)
    fsel, tsel, bsel = [1, 2, 4], [0, 2], list(range(0, len(g["bls"]), 2))
    with emulated_kernels():
        vm = build_mapper(g, 'cpu', torch.float64)
        vm.set_freq_inds(fsel)
        vm.set_time_inds(times=g["times"][tsel])
        vm.set_bl_inds(bls=[tuple(int(x) for x in g["bls"][i]) for i in bsel])
        assert (vm.Nfreqs, vm.Ntimes, vm.Nbls) == (3, 2, len(bsel))
        vm.set_normalization('Aw')
        maps, P = vm.make_map()
    p = torch.as_tensor(g["beam_params"])
    fs = freqs[fsel]
    beam_sub = lambda z, a: orc.airy_response(p, z, a, fs, powerbeam=True)[0, 0, 0]
    v = torch.as_tensor(g["vis"])[0, 0][bsel][:, tsel][:, :, fsel]
    w = torch.as_tensor(g["icov"])[0, 0][bsel][:, tsel][:, :, fsel]
    mo, Po, Do = orc.vismapper_make_map(v, w, blvecs[bsel], [zenaz[t] for t in tsel], fs,
                                        len(g["ra"]), beam_sub, fov=float(g["fov"]), method='Aw')
    assert relmax(maps, mo) < 1e-11 and relmax(P, Po) < 1e-11 and relmax(vm.D, Do) < 1e-11


def test_mapper_needs_cuda():
    g = load()
    vm = build_mapper(g, 'cpu', torch.float64)
    with pytest.raises(RuntimeError, match="CUDA"):
        vm.make_map()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ant", [False, True])
def test_mapper_cuda_matches_reference(dtype, ant, monkeypatch):
    if DOUBLE:
        pytest.skip("covered by test_mapper_host_logic_with_emulated_kernels")
    if ant:
        monkeypatch.setattr(ops, "ANT_FWD_MIN_FILL", 0.0)
        monkeypatch.setattr(ops, "ANT_BWD_MIN_FILL", 0.0)
    check_mapper('cuda', dtype, 1e-10 if dtype == torch.float64 else 1e-5)
