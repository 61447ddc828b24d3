"""
Parity of the CUDA path (through the C ABI of libb200rime.so) with the CPU oracle and with
the golden vectors produced by the unmodified reference.

Tolerances (BASELINE.json north star): relative max-norm 1e-5 for the float32/complex64
kernels -- judged against the FLOAT64 oracle, because the reference's own complex64 path is
only good to 1e-4 on 300 m baselines -- and 1e-10 for the
float64/complex128 kernels.

Set B200RIME_TEST_DOUBLE=1 to dry-run this file's logic on a CPU with the torch test double
of the kernels (tests/cpu_double.py); that mode checks the tests, not the product.
"""
import json
import os

import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from bayeslim_b200 import ops, _lib
from oracle import rime_oracle as orc
from tests import model_cases as mc
from tests import oracle_cases as oc
import workloads

pytestmark = pytest.mark.gpu

DOUBLE = os.environ.get("B200RIME_TEST_DOUBLE") == "1"
DEV = 'cpu' if DOUBLE else 'cuda'
TOL = {torch.float32: 1e-5, torch.float64: 1e-10}
ERRLOG = {}


@pytest.fixture(autouse=True)
def _kernels():
    if DOUBLE:
        from tests.cpu_double import emulated_kernels
        with emulated_kernels():
            yield
    else:
        assert torch.cuda.is_available(), "GPU tests need a CUDA device"
        yield


@pytest.fixture(scope="module", autouse=True)
def _dump_errors():
    yield
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump(ERRLOG, f, indent=1, sort_keys=True)
    except OSError:
        pass


def relmax(a, b, tag=None):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    e = float((a.to(b.dtype) - b).abs().max() / b.abs().max())
    if tag:
        ERRLOG[tag] = e
    return e


def _rand_problem(nbl, nfreq, ns_list, dtype, seed=0, blmax=300.0, fmax=200e6):
    g = torch.Generator().manual_seed(seed)
    zen = [torch.rand(n, generator=g, dtype=torch.float64) * 89.0 for n in ns_list]
    az = [torch.rand(n, generator=g, dtype=torch.float64) * 360.0 for n in ns_list]
    geom = ops.Geometry([z.to(DEV) for z in zen], [a.to(DEV) for a in az], DEV)
    blv = (torch.rand(nbl, 3, generator=g, dtype=torch.float64) - 0.5) * 2 * blmax
    blv[:, 2] *= 0.01
    if nbl > 2:
        blv[1] = 0.0                                  # an autocorrelation baseline
    freqs = torch.linspace(100e6, fmax, nfreq, dtype=torch.float64)
    planes = [torch.rand(2, nfreq, n, generator=g, dtype=torch.float64) for n in ns_list]
    return geom, zen, az, blv, freqs, planes


def _oracle_fringe_sum(planes, zen, az, blv, freqs, conj=False):
    out = []
    for X, z, a in zip(planes, zen, az):
        F = orc.gen_fringe(blv, z, a, freqs, conj=conj)            # (nbl, nf, ns) complex128
        out.append(torch.einsum('bfs,pfs->pbf', F, X.to(F.dtype)))
    return torch.stack(out, dim=2)                                 # (nplane, nbl, nt, nf)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("nbl,nfreq,ns_list,uniform", [
    (5, 10, [1, 130], True),              # tiny, ragged, Nf < KC
    (130, 100, [300, 0, 64], True),       # Nbl not a multiple of 128, empty time, Nf % KC != 0
    (40, 70, [257], False),               # direct per-channel phases
])
def test_fringe_sum_forward_backward(dtype, nbl, nfreq, ns_list, uniform):
    geom, zen, az, blv, freqs, planes = _rand_problem(nbl, nfreq, ns_list, dtype, seed=nbl)
    if not uniform:
        freqs = freqs + torch.linspace(0, 1, nfreq, dtype=torch.float64) ** 2 * 3e6
    f64 = freqs.to(DEV)
    X = [p.detach().clone().to(device=DEV, dtype=dtype).requires_grad_(True) for p in planes]
    b = blv.detach().clone().to(DEV).requires_grad_(True)
    for conj in (False, True):
        A = ops.pack_planes(geom, X)
        V = ops.fringe_sum(A, b, geom, f64, nfreq, conj=conj, uniform=uniform)
        Xo = [p.detach().clone().requires_grad_(True) for p in planes]
        bo = blv.detach().clone().requires_grad_(True)
        Vo = _oracle_fringe_sum(Xo, zen, az, bo, freqs, conj=conj)
        tag = "fringe_sum/%s/nbl%d/conj%d" % (str(dtype)[6:], nbl, conj)
        assert V.shape == Vo.shape
        assert relmax(V, Vo, tag + "/V") < TOL[dtype]
        # zenith-free known answer: the autocorrelation row is the plain source sum
        if nbl > 2:
            auto = torch.stack([x.sum(-1) for x in planes], dim=1)      # (nplane, nt, nf)
            assert relmax(V[:, 1].real, auto) < TOL[dtype]
        gen = torch.Generator().manual_seed(7)
        G = torch.complex(torch.randn(Vo.shape, generator=gen, dtype=torch.float64),
                          torch.randn(Vo.shape, generator=gen, dtype=torch.float64))
        oc.real_loss(Vo, G).backward()
        Gd = G.to(device=DEV, dtype=V.dtype)
        torch.sum(Gd.real * V.real + Gd.imag * V.imag).backward()
        for t in range(len(ns_list)):
            if ns_list[t]:
                assert relmax(X[t].grad, Xo[t].grad, tag + "/dA%d" % t) < TOL[dtype] * 2
        assert relmax(b.grad, bo.grad, tag + "/dbl") < TOL[dtype] * 2
        for x in X:
            x.grad = None
        b.grad = None


def _antenna_problem(na, seed):
    g = torch.Generator().manual_seed(seed)
    antv = (torch.rand(na, 3, generator=g, dtype=torch.float64) - 0.5) * 600.0
    antv[:, 2] *= 0.01
    ii, jj = np.triu_indices(na, k=1)
    flip = (np.arange(len(ii)) % 3) == 0              # every third pair listed as (j, i)
    i = np.where(flip, jj, ii)
    j = np.where(flip, ii, jj)
    autos = np.asarray([0, na // 2, na - 1])
    return antv, np.concatenate([i, autos]), np.concatenate([j, autos])


@pytest.mark.parametrize("na,nfreq,ns_list", [
    (70, 70, [300, 0, 64]),        # two antenna blocks, empty time, Nf % 64 != 0
    (130, 20, [520]),              # three blocks (64, 64, 2), several source tiles per unit
    (100, 24, [200]),              # last block of 36 antennas: not narrow (> 32)
])
def test_antenna_factorised_fringe_sum(na, nfreq, ns_list):
    """float32 antenna-factorised kernels (conj(E_i) E_j complex MACs) against the float64
    oracle: visibilities, dL/dA and dL/d(antenna positions), both fringe signs, baselines in
    either orientation and autocorrelations."""
    dtype = torch.float32
    geom, zen, az, _, freqs, planes = _rand_problem(4, nfreq, ns_list, dtype, seed=na)
    antv, i, j = _antenna_problem(na, seed=na)
    til = ops.AntTiling(i, j, na, DEV)
    assert til.unique and til.ntile == (na + 63) // 64 * ((na + 63) // 64 + 1) // 2
    f64 = freqs.to(DEV)
    X = [p.detach().clone().to(device=DEV, dtype=dtype).requires_grad_(True) for p in planes]
    a = antv.detach().clone().to(DEV).requires_grad_(True)
    it, jt = torch.as_tensor(i), torch.as_tensor(j)
    for conj in (False, True):
        A = ops.pack_planes(geom, X)
        V = ops.fringe_sum_ant(A, a, til, geom, f64, nfreq, conj=conj)
        Xo = [p.detach().clone().requires_grad_(True) for p in planes]
        ao = antv.detach().clone().requires_grad_(True)
        Vo = _oracle_fringe_sum(Xo, zen, az, ao[jt] - ao[it], freqs, conj=conj)
        tag = "antfringe/na%d/conj%d" % (na, conj)
        assert V.shape == Vo.shape
        assert relmax(V, Vo, tag + "/V") < TOL[dtype]
        gen = torch.Generator().manual_seed(7)
        G = torch.complex(torch.randn(Vo.shape, generator=gen, dtype=torch.float64),
                          torch.randn(Vo.shape, generator=gen, dtype=torch.float64))
        oc.real_loss(Vo, G).backward()
        Gd = G.to(device=DEV, dtype=V.dtype)
        torch.sum(Gd.real * V.real + Gd.imag * V.imag).backward()
        for t in range(len(ns_list)):
            if ns_list[t]:
                assert relmax(X[t].grad, Xo[t].grad, tag + "/dA%d" % t) < TOL[dtype] * 2
        assert relmax(a.grad, ao.grad, tag + "/dant") < TOL[dtype] * 2
        for x in X:
            x.grad = None
        a.grad = None
    # one owner per output, fixed summation order: bitwise reproducible
    with torch.no_grad():
        A = ops.pack_planes(geom, X)
        V1 = ops.fringe_sum_ant(A, a, til, geom, f64, nfreq)
        V2 = ops.fringe_sum_ant(A, a, til, geom, f64, nfreq)
    assert torch.equal(V1, V2)


@pytest.mark.parametrize("antpos", [True, False])
def test_antenna_path_is_selected_and_matches_baseline_path(antpos):
    """RIME picks the antenna-factorised kernels for an all-pairs HERA-350 group and the
    baseline-owned kernels otherwise; both give the same visibilities and gradients.  Without
    an antenna-position gradient the backward walks one triangle of the cotangent matrix."""
    if DOUBLE:
        pytest.skip("61075-baseline problem is too large for the CPU test double")
    out = {}
    for flag in ("1", "0"):
        os.environ["B200RIME_ANT"] = flag
        try:
            rime = workloads.pixel_interp(16, 96, 2, DEV, torch.float32, antpos_param=antpos)
            assert (rime._ant_tiling(torch.device(DEV)) is not None) == (flag == "1")
            V = rime().data
            gen = torch.Generator().manual_seed(5)
            G = torch.randn(V.shape, generator=gen, dtype=torch.float64).to(DEV)
            torch.sum(G.to(V.real.dtype) * (V.real + 0.5 * V.imag)).backward()
            out[flag] = (V.detach(), rime.sky.params.grad, rime.beam.params.grad,
                         rime.array.antvecs.grad if antpos else rime.sky.params.grad)
        finally:
            os.environ.pop("B200RIME_ANT", None)
    for x, y, nm in zip(out["1"], out["0"], ("V", "dsky", "dbeam", "dant")):
        assert relmax(x, y, "ant_vs_bl_path/antpos%d/%s" % (antpos, nm)) < 2e-5, nm


def test_c4_reduced_polarised_paths_agree():
    """BASELINE config 4 at reduced size (HERA-350 all pairs, 4-pol Jones beams, nside-16 sky with
    Stokes I, Q, U, 96 channels, 2 times): the tiled polarised route through the antenna-
    factorised kernels against the same through the baseline-owned kernels, and a baseline /
    channel subset against the float64 oracle."""
    if DOUBLE:
        pytest.skip("61075-baseline problem is too large for the CPU test double")
    out = {}
    for flag in ("1", "0"):
        os.environ["B200RIME_ANT"] = flag
        try:
            rime = workloads.pixel_interp_pol(16, 96, 2, DEV, torch.float32)
            V = rime().data
            assert tuple(V.shape) == (2, 2, 61075, 2, 96)
            gen = torch.Generator().manual_seed(11)
            G = torch.randn(V.shape, generator=gen, dtype=torch.float64).to(DEV)
            torch.sum(G.to(V.real.dtype) * (V.real - 0.3 * V.imag)).backward()
            out[flag] = (V.detach(), rime.sky.sky.params.grad, rime.beam.params.grad)
        finally:
            os.environ.pop("B200RIME_ANT", None)
    for x, y, nm in zip(out["1"], out["0"], ("V", "dsky", "dbeam")):
        assert relmax(x, y, "c4_reduced/ant_vs_bl/" + nm) < 2e-5, nm
    # oracle on a subset: 40 baselines, all polarisations
    sel = list(range(0, 61075, 1600))
    with torch.no_grad():
        sky = rime.sky.forward().data.cpu().double()                     # (2, 2, Nf, Npix)
        bmap = rime.beam.params.detach().cpu().double()
    freqs = rime.array.freqs.cpu().double()
    R = rime.beam.R
    theta, phi = R.theta_grid.cpu().double(), R.phi_grid.cpu().double()
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(theta, phi, z, a, 'linear')
        return orc.interp_map(bmap, inds, wgts)
    bls = [rime.sim_bls[i] for i in sel]
    Vo = orc.rime_forward(sky, zenaz, beam_fn, bls, rime.sim_blvecs.cpu().double()[sel], freqs,
                          fov=180.0, powerbeam=False)
    assert relmax(out["1"][0][:, :, sel], Vo, "c4_reduced/float32/V_vs_oracle") < 1e-5


def test_forward_is_bitwise_reproducible():
    geom, zen, az, blv, freqs, planes = _rand_problem(200, 128, [500, 700], torch.float32, seed=3)
    X = [p.to(device=DEV, dtype=torch.float32) for p in planes]
    f64 = freqs.to(DEV)
    outs = []
    for _ in range(3):
        A = ops.pack_planes(geom, X)
        outs.append(ops.fringe_sum(A, blv.to(DEV), geom, f64, 128))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_pack_unpack_roundtrip_exact(dtype):
    geom, zen, az, blv, freqs, planes = _rand_problem(3, 77, [5, 200, 129], dtype, seed=11)
    X = [p.detach().clone().to(device=DEV, dtype=dtype).requires_grad_(True) for p in planes]
    A = ops.pack_planes(geom, X)
    kc = _lib.KC[ops._sfx(dtype)]
    assert A.shape == (2, (77 + kc - 1) // kc, geom.S, kc)
    # padded sources / channels are exactly zero; payload is bit exact
    rows = A.permute(0, 1, 3, 2).reshape(2, -1, geom.S)
    for t, n in enumerate(geom.ns):
        seg = rows[:, :, geom.toff[t]:geom.toff[t + 1]]
        assert torch.equal(seg[:, :77, :n], X[t].detach())
        assert float(seg[:, :77, n:].abs().max() if geom.ns_pad[t] > n else 0) == 0
        assert float(seg[:, 77:].abs().max()) == 0
    w = torch.rand_like(A)
    (A * w).sum().backward()
    wrows = w.permute(0, 1, 3, 2).reshape(2, -1, geom.S)
    for t, n in enumerate(geom.ns):
        assert torch.equal(X[t].grad, wrows[:, :77, geom.toff[t]:geom.toff[t] + n])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", list(mc.CASES))
def test_golden_cases(name, dtype):
    """Every golden fixture of the unmodified reference through the CUDA path."""
    g = oc.load(name)
    build, gkeys = mc.CASES[name]
    rime, leaves = build(g, DEV, dtype)
    vd = rime()
    V = vd.data
    assert tuple(V.shape) == g["vis"].shape
    assert V.dtype == (torch.complex64 if dtype == torch.float32 else torch.complex128)
    tag = "golden/%s/%s" % (name, str(dtype)[6:])
    assert relmax(V, g["vis"], tag + "/V") < TOL[dtype]
    G = torch.as_tensor(g["G"]).to(device=DEV, dtype=V.dtype)
    torch.sum(G.real * V.real + G.imag * V.imag).backward()
    for k, gk in gkeys.items():
        tol = TOL[dtype] * (5 if dtype == torch.float32 else 10)
        assert relmax(leaves[k].grad, g[gk], tag + "/" + gk) < tol, (k, gk)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_golden_batched_and_minibatch_invariance(dtype):
    g = oc.load("rime_batched")
    rime, _ = mc.build_pixel_interp(g, DEV, dtype, params=(), interp_mode='quadratic')
    with torch.no_grad():
        vis = rime()
        assert relmax(vis.data, g["vis"], "golden/rime_batched/%s/V" % str(dtype)[6:]) < \
            max(TOL[dtype], 1e-9)
        rime.setup_sim_times(ba.utils.split_into_groups(torch.as_tensor(g["times"]), Nelem=2))
        bls = mc.bl_list(g["bls"])
        rime.setup_sim_bls([bls[:11], bls[11:]])
        batched = rime.run_batches()
    assert batched.data.shape == vis.data.shape
    # same kernels, same per-(baseline, time) summation order -> identical to rounding of the
    # unit split; reference asserts 1e-10 in float64 (tests/test_rime.py:51)
    assert relmax(batched.data, vis.data) < (1e-10 if dtype == torch.float64 else 2e-6)


def _oracle_of_workload(rime, kind, bl_sel, f_sel, dtype=torch.float64):
    """fp64 oracle visibilities of a workloads.* model on a subset of baselines / channels."""
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]
    freqs = rime.array.freqs.detach().cpu().double()[f_sel]
    bls = [rime.sim_bls[i] for i in bl_sel]
    blvecs = rime.sim_blvecs.detach().cpu().double()[bl_sel]
    with torch.no_grad():
        sky = rime.sky.forward().data.detach().cpu().double()[:, :, f_sel]
        if kind == 'airy':
            p = rime.beam.params.detach().cpu().double()
            beam_fn = lambda z, a: orc.airy_response(p, z, a, freqs, powerbeam=True)
        else:
            bmap = rime.beam.params.detach().cpu().double().abs()[:, :, :, f_sel]
            tg, pg = rime.beam.R.theta_grid.cpu(), rime.beam.R.phi_grid.cpu()

            def beam_fn(z, a):
                inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
                return orc.interp_map(bmap, inds, wgts)
        return orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=rime.beam.fov,
                                bl_chunk=16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_c2_reduced_vs_oracle(dtype):
    """HERA-37 all cross baselines x 1500 sources x 96 freqs x 3 times, full oracle."""
    rime = workloads.point_airy(1500, 96, 3, DEV, dtype)
    vd = rime()
    V = vd.data
    Vo = _oracle_of_workload(rime, 'airy', list(range(len(rime.sim_bls))), slice(None))
    assert relmax(V, Vo, "c2_reduced/%s/V" % str(dtype)[6:]) < TOL[dtype]
    (V.abs() ** 2).sum().backward()
    assert torch.isfinite(rime.sky.params.grad).all() and rime.sky.params.grad.abs().max() > 0


def test_c2_full_size_properties():
    """BASELINE config 2 at full size (666 bl x 10^4 src x 256 freqs x 60 times), complex64:
    size-independent checks + oracle on a baseline/channel subset."""
    if DOUBLE:
        pytest.skip("full size needs the GPU")
    rime = workloads.point_airy(10000, 256, 60, DEV, torch.float32)
    with torch.no_grad():
        V = rime().data
    assert V.shape == (1, 1, 666, 60, 256)
    evals = workloads.count_evals(rime)
    assert 3e10 < evals < 1.1e11
    # oracle on 12 baselines x 16 channels x all sources x 3 of the times
    bl_sel = list(range(0, 666, 60))
    f_sel = slice(0, 256, 16)
    rime_t = workloads.point_airy(10000, 256, 60, DEV, torch.float32)
    Vo = _oracle_of_workload(rime, 'airy', bl_sel, f_sel)
    sub = V[:, :, bl_sel][..., f_sel]
    assert relmax(sub[:, :, :, ::20], Vo[:, :, :, ::20], "c2_full/f32/V_subset") < 1e-5
    # linearity in the sky: V(2.5 * I) = 2.5 * V(I)
    with torch.no_grad():
        rime.sky.params.data[0, 0, 0] *= 2.5
        V2 = rime().data
    assert relmax(V2, 2.5 * V) < 5e-6     # two float32 runs, each ~1e-6 from the truth
    # Hermitian symmetry: swapping the antennas of a baseline conjugates the visibility
    rime_c = workloads.point_airy(10000, 256, 60, DEV, torch.float32)
    rime_c.setup_sim_bls([(b[1], b[0]) for b in rime_c.sim_bls[:50]])
    rime_c.setup_sim_times(rime_c.all_sim_times[:4])
    with torch.no_grad():
        Vc = rime_c().data
    assert relmax(Vc, (V2 / 2.5)[:, :, :50, :4].conj()) < 5e-6
    del rime_t


def test_c3_full_size_subset_and_gradients():
    """BASELINE config 3 shape at full source/frequency size on a baseline subset that the
    fp64 oracle can follow: nside-128 PixelSky, rect-interpolated PixelBeam, 1024 freqs."""
    if DOUBLE:
        pytest.skip("full size needs the GPU")
    rime = workloads.pixel_interp(128, 1024, 1, DEV, torch.float32, n_bl=256, antpos_param=True)
    vd = rime()
    V = vd.data
    assert V.shape[2:] == (256, 1, 1024)
    bl_sel = list(range(0, 256, 32))
    f_sel = slice(0, 1024, 64)
    Vo = _oracle_of_workload(rime, 'interp', bl_sel, f_sel)
    assert relmax(V[:, :, bl_sel][..., f_sel], Vo, "c3_subset/f32/V") < 1e-5
    (V.abs() ** 2).sum().backward()
    for p in (rime.sky.params, rime.beam.params, rime.array.antvecs):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0
    # sources below the horizon at this time receive exactly zero gradient
    rec = list(rime._geom_cache.values())[0]
    mask = torch.ones(rime.sky.params.shape[-1], dtype=torch.bool, device=V.device)
    mask[rec.cuts[0]] = False
    assert float(rime.sky.params.grad[..., mask].abs().max()) == 0.0


def _oracle_grad_subset(rime, bl_sel, f_idx, G_sub, kind='interp'):
    """fp64 oracle V and autograd gradients (sky params, beam params, antenna positions) of a
    workloads.pixel_interp model restricted to baselines bl_sel and channels f_idx, for the
    cotangent G_sub (1, 1, len(bl_sel), Nt, len(f_idx)) -- exact gradients of a loss whose
    cotangent vanishes elsewhere."""
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]
    f_idx = torch.as_tensor(f_idx)
    freqs = rime.array.freqs.detach().cpu().double()[f_idx]
    bls = [rime.sim_bls[i] for i in bl_sel]
    antvecs = rime.array.antvecs.detach().cpu().double().requires_grad_(True)
    sp = rime.sky.params.detach().cpu().double()[:, :, f_idx].requires_grad_(True)
    bp = rime.beam.params.detach().cpu().double()[:, :, :, f_idx].requires_grad_(True)
    blvecs = orc.get_blvecs(antvecs, rime.array.ants, bls)
    bmap = orc.pixel_response_forward(bp, powerbeam=True)
    tg, pg = rime.beam.R.theta_grid.cpu(), rime.beam.R.phi_grid.cpu()

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
        return orc.interp_map(bmap, inds, wgts)

    Vo = orc.rime_forward(sp * float(rime.sky.px_area), zenaz, beam_fn, bls, blvecs, freqs,
                          fov=rime.beam.fov, bl_chunk=4)
    oc.real_loss(Vo, G_sub).backward()
    return Vo.detach(), sp.grad, bp.grad, antvecs.grad


@pytest.mark.parametrize("route", ["tc", "fp32"])
def test_c3_full_size_all_pairs_factorised_vs_oracle(route):
    """BASELINE config 3 exactly as benchmarked -- HERA-350, all 61,075 cross baselines,
    nside-128 PixelSky (98 k sources above the horizon), rect-interpolated PixelBeam, 1024
    channels, 1 time, float32 -- through the antenna-factorised kernels (route 'tc': tensor-core
    forward; 'fp32': FP32-pipe forward), against the fp64 oracle:
      * V on 16 baselines (shortest, longest, outriggers, random) x 64 channels x all sources;
      * gradients to sky, beam map and antenna positions for a cotangent that is non-zero only on
        8 baselines x 32 channels, for which the oracle's autograd gives the exact answer."""
    if DOUBLE:
        pytest.skip("full size needs the GPU")
    os.environ["B200RIME_TC"] = "1" if route == "tc" else "0"
    try:
        rime = workloads.pixel_interp(128, 1024, 1, DEV, torch.float32, antpos_param=True)
        dev = torch.device(DEV, torch.cuda.current_device())
        assert rime._ant_tiling(dev) is not None, "C3 must run on the antenna-factorised kernels"
        assert (rime._tc_tiling(dev) is not None) == (route == "tc")
        nbl = len(rime.sim_bls)
        assert nbl == 61075
        V = rime().data
        assert V.shape == (1, 1, nbl, 1, 1024)
        blen = rime.sim_blvecs.detach().norm(dim=1).cpu().numpy()
        order = np.argsort(blen)
        rng = np.random.default_rng(5)
        bl_sel = sorted(set(order[:3].tolist() + order[-3:].tolist()
                            + rng.choice(nbl, 10, replace=False).tolist()))
        f_sel = slice(0, 1024, 16)
        Vo = _oracle_of_workload(rime, 'interp', bl_sel, f_sel)
        tag = "c3_full_allpairs/%s" % route
        # max-norm over the visibility tensor: the subset holds the shortest baselines,
        # whose visibilities are the largest of the tensor
        scale = float(V.abs().max())
        errV = float((V[:, :, bl_sel][..., f_sel].cpu().to(torch.complex128) - Vo).abs().max()) / scale
        ERRLOG[tag + "/V"] = errV
        assert errV < 1e-5

        # sparse cotangent -> exact oracle gradients
        gb = sorted(set(order[:2].tolist() + order[-2:].tolist()
                        + rng.choice(nbl, 4, replace=False).tolist()))
        gf = list(range(7, 1024, 32))
        gen = torch.Generator().manual_seed(11)
        G_sub = torch.complex(torch.randn(1, 1, len(gb), 1, len(gf), generator=gen, dtype=torch.float64),
                              torch.randn(1, 1, len(gb), 1, len(gf), generator=gen, dtype=torch.float64))
        G = torch.zeros(V.shape, dtype=torch.complex64, device=V.device)
        gbt, gft = torch.as_tensor(gb, device=V.device), torch.as_tensor(gf, device=V.device)
        G[0, 0, gbt[:, None], 0, gft[None, :]] = G_sub[0, 0, :, 0].to(V.device, torch.complex64)
        torch.sum(G.real * V.real + G.imag * V.imag).backward()
        Vo2, dsp, dbp, dant = _oracle_grad_subset(rime, gb, gf, G_sub)
        assert relmax(V[:, :, gb][..., gf], Vo2, tag + "/V_gradsubset") < 1e-5 * scale / float(Vo2.abs().max())
        gs = rime.sky.params.grad
        assert relmax(gs[:, :, gf], dsp, tag + "/dsky") < 5e-5
        assert relmax(rime.beam.params.grad[:, :, :, gf], dbp, tag + "/dbeam") < 5e-5
        assert relmax(rime.array.antvecs.grad, dant, tag + "/dantvecs") < 5e-5
        # channels without cotangent receive exactly zero
        mask = torch.ones(1024, dtype=torch.bool, device=V.device)
        mask[gft] = False
        assert float(gs[:, :, mask].abs().max()) == 0.0
    finally:
        os.environ.pop("B200RIME_TC", None)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("which", ["interp", "airy"])
def test_golden_pointing_offset(which, dtype):
    """Beam pointing offset (reference beam_model.py:244-256, 1631-1678) against golden vectors
    of the unmodified reference: fused interpolation route and generic (Airy) route."""
    g = oc.load("rime_pointing")
    build = mc.build_pointing_interp if which == "interp" else mc.build_pointing_airy
    rime, leaves = build(g, DEV, dtype)
    V = rime().data
    G = torch.as_tensor(g["G"]).to(device=DEV, dtype=V.dtype)
    torch.sum(G.real * V.real + G.imag * V.imag).backward()
    tol = TOL[dtype]
    tag = "golden/rime_pointing_%s/%s" % (which, str(dtype)[6:])
    assert relmax(V, g["vis_" + which], tag + "/V") < tol
    assert relmax(leaves["sky"].grad, g["grad_sky_" + which], tag + "/dsky") < 5 * tol
    if which == "interp":
        assert relmax(leaves["beam"].grad, g["grad_beam_interp"], tag + "/dbeam") < 5 * tol
        assert relmax(leaves["antvecs"].grad, g["grad_antvecs_interp"], tag + "/dant") < 5 * tol


def test_c4_nside64_all_pols_and_gradients_vs_oracle():
    """BASELINE config 4 family at nside 64: HERA-350, all 61,075 cross baselines, 4-pol real
    Jones beams interpolated from rect-grid maps, PixelSky with Stokes I, Q, U (24.5 k sources
    above the horizon), 256 channels, float32, through the tensor-core kernels (one launch per
    coherency plane), against the fp64 oracle: all four polarisation products of V on a baseline
    / channel subset, and the gradients to the sky parameters, the Jones maps and the antenna
    positions for a sparse cotangent whose oracle autograd is exact."""
    if DOUBLE:
        pytest.skip("full size needs the GPU")
    rime = workloads.pixel_interp_pol(64, 256, 1, DEV, torch.float32, antpos_param=True)
    dev = torch.device(DEV, torch.cuda.current_device())
    assert rime._tc_tiling(dev) is not None
    nbl = len(rime.sim_bls)
    V = rime().data
    assert tuple(V.shape) == (2, 2, nbl, 1, 256)
    blen = rime.sim_blvecs.detach().norm(dim=1).cpu().numpy()
    order = np.argsort(blen)
    rng = np.random.default_rng(7)
    gb = sorted(set(order[:2].tolist() + order[-1:].tolist() + rng.choice(nbl, 3, replace=False).tolist()))
    gf = list(range(5, 256, 32))
    gen = torch.Generator().manual_seed(13)
    shp = (2, 2, len(gb), 1, len(gf))
    G_sub = torch.complex(torch.randn(shp, generator=gen, dtype=torch.float64),
                          torch.randn(shp, generator=gen, dtype=torch.float64))
    G = torch.zeros(V.shape, dtype=torch.complex64, device=V.device)
    gbt, gft = torch.as_tensor(gb, device=V.device), torch.as_tensor(gf, device=V.device)
    G[:, :, gbt[:, None], 0, gft[None, :]] = G_sub[:, :, :, 0].to(V.device, torch.complex64)
    torch.sum(G.real * V.real + G.imag * V.imag).backward()
    # oracle with autograd on the subset
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]
    fi = torch.as_tensor(gf)
    freqs = rime.array.freqs.detach().cpu().double()[fi]
    antvecs = rime.array.antvecs.detach().cpu().double().requires_grad_(True)
    pix = rime.sky.sky
    sp = pix.params.detach().cpu().double()[:, :, fi].requires_grad_(True)
    bp = rime.beam.params.detach().cpu().double()[:, :, :, fi].requires_grad_(True)
    bls = [rime.sim_bls[i] for i in gb]
    blvecs = orc.get_blvecs(antvecs, rime.array.ants, bls)
    sky = orc.stokes_to_coherency(sp * float(pix.px_area))
    bmap = orc.pixel_response_forward(bp, powerbeam=False, realbeam=True)
    tg, pg = rime.beam.R.theta_grid.cpu().double(), rime.beam.R.phi_grid.cpu().double()

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
        return orc.interp_map(bmap, inds, wgts)

    Vo = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=180.0, powerbeam=False,
                          bl_chunk=2)
    oc.real_loss(Vo, G_sub).backward()
    scale = float(V.abs().max())
    sub = V[:, :, gbt[:, None], 0, gft[None, :]].detach().cpu().to(torch.complex128)
    errV = float((sub - Vo[:, :, :, 0].detach()).abs().max()) / scale
    ERRLOG["c4_nside64/tc/V"] = errV
    assert errV < 1e-5
    assert relmax(pix.params.grad[:, :, gf], sp.grad, "c4_nside64/tc/dsky") < 5e-5
    assert relmax(rime.beam.params.grad[:, :, :, gf], bp.grad, "c4_nside64/tc/dbeam") < 5e-5
    assert relmax(rime.array.antvecs.grad, antvecs.grad, "c4_nside64/tc/dantvecs") < 5e-5


def test_gradients_small_c3_vs_oracle_autograd():
    """Gradients to sky, beam map and antenna positions of a C3-shaped model small enough for
    the oracle's autograd (nside 8, HERA-37, 48 freqs, 2 times), float32 and float64."""
    for dtype in (torch.float32, torch.float64):
        rime = workloads.pixel_interp(8, 48, 2, DEV, dtype, n_bl=60, antpos_param=True,
                                      layout='hera37', dgrid=5.0)
        V = rime().data
        gen = torch.Generator().manual_seed(1)
        G = torch.complex(torch.randn(V.shape, generator=gen, dtype=torch.float64),
                          torch.randn(V.shape, generator=gen, dtype=torch.float64))
        Gd = G.to(device=DEV, dtype=V.dtype)
        torch.sum(Gd.real * V.real + Gd.imag * V.imag).backward()
        # oracle
        zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]
        freqs = rime.array.freqs.detach().cpu().double()
        antvecs = rime.array.antvecs.detach().cpu().double().requires_grad_(True)
        sp = rime.sky.params.detach().cpu().double().requires_grad_(True)
        bp = rime.beam.params.detach().cpu().double().requires_grad_(True)
        blvecs = orc.get_blvecs(antvecs, rime.array.ants, rime.sim_bls)
        bmap = orc.pixel_response_forward(bp, powerbeam=True)
        tg, pg = rime.beam.R.theta_grid.cpu(), rime.beam.R.phi_grid.cpu()

        def beam_fn(z, a):
            inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
            return orc.interp_map(bmap, inds, wgts)

        Vo = orc.rime_forward(sp * float(rime.sky.px_area), zenaz, beam_fn, rime.sim_bls, blvecs,
                              freqs, fov=180.0)
        oc.real_loss(Vo, G).backward()
        tag = "small_c3/%s" % str(dtype)[6:]
        tol = TOL[dtype]
        assert relmax(V, Vo, tag + "/V") < tol
        assert relmax(rime.sky.params.grad, sp.grad, tag + "/dsky") < 5 * tol
        assert relmax(rime.beam.params.grad, bp.grad, tag + "/dbeam") < 5 * tol
        assert relmax(rime.array.antvecs.grad, antvecs.grad, tag + "/dantvecs") < 5 * tol


def test_airy_full_gradient_option_matches_finite_difference():
    """full_grad=True uses d(2J1/x)/dx = 2J0/x - 4J1/x^2 (the reference's autograd drops the
    J1' term); check against a central finite difference of the CUDA forward in float64."""
    g = oc.load("rime_point_airy")
    rime, leaves = mc.build_point_airy(g, DEV, torch.float64)
    rime.beam.R.full_grad = True
    G = torch.as_tensor(g["G"]).to(DEV)
    V = rime().data
    torch.sum(G.real * V.real + G.imag * V.imag).backward()
    analytic = float(rime.beam.params.grad.reshape(-1)[0])
    eps = 1e-5
    vals = []
    for s in (+1, -1):
        with torch.no_grad():
            rime.beam.params.data += s * eps
            Vs = rime().data
            vals.append(float(torch.sum(G.real * Vs.real + G.imag * Vs.imag)))
            rime.beam.params.data -= s * eps
    fd = (vals[0] - vals[1]) / (2 * eps)
    ERRLOG["airy_full_grad/rel_err_vs_fd"] = abs(analytic - fd) / abs(fd)
    # the forward uses the reference's (torch) J1, which is itself only good to ~5e-7, so its
    # numerical derivative and the analytic Bessel derivative agree to that level, not better
    assert abs(analytic - fd) / abs(fd) < 2e-5


def test_integration_md_stub_runs():
    """The ctypes stub printed in INTEGRATION.md (what a BayesLIM maintainer would paste into
    rime_model.py) must run as written against the built library."""
    if DOUBLE:
        pytest.skip("calls the real library")
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S)
             if "def b200_prod_and_sum" in b][0]
    block = block.replace('"libb200rime.so"', repr(_lib.LIB_PATH))
    ns = {}
    exec(block, ns)
    g = torch.Generator().manual_seed(4)
    Nf, Ns, Nbl = 70, 300, 9
    X = torch.rand(Nf, Ns, generator=g, dtype=torch.float32)   # session default may be float64
    zen = torch.rand(Ns, generator=g, dtype=torch.float64) * 89
    az = torch.rand(Ns, generator=g, dtype=torch.float64) * 360
    blv = (torch.rand(Nbl, 3, generator=g, dtype=torch.float64) - 0.5) * 300
    freqs = torch.linspace(100e6, 200e6, Nf, dtype=torch.float64)
    V = ns["b200_prod_and_sum"](X.cuda(), zen.cuda(), az.cuda(), blv.cuda(), freqs.cuda())
    F = orc.gen_fringe(blv, zen, az, freqs)
    Vo = torch.einsum('bfs,fs->bf', F, X.to(F.dtype))
    assert relmax(V, Vo, "integration_stub/V") < 1e-5


# ------------------------------------------------------------------ remaining SURVEY 8(a) rows
def _small_array(device, freqs):
    ants, vecs = ba.utils._make_hex(2, D=14.6)
    return ants, ba.telescope_model.ArrayModel(
        ba.utils.AntposDict(ants, torch.as_tensor(vecs, dtype=torch.float64, device=device)),
        freqs=freqs, device=device)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_c1_full_size_vs_oracle(dtype):
    """BASELINE config 1 at full size: HERA-37, 63 unique baselines, 1000 point sources
    (power law), Airy beam, 64 freqs, 10 times."""
    rime = workloads.point_airy(1000, 64, 10, DEV, dtype, bls='uniq')
    with torch.no_grad():
        V = rime().data
    assert V.shape == (1, 1, 63, 10, 64)
    Vo = _oracle_of_workload(rime, 'airy', list(range(63)), slice(None))
    assert relmax(V, Vo, "c1_full/%s/V" % str(dtype)[6:]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gauss_response_generic_path(dtype):
    """A response without a fused builder (GaussResponse, beam_model.py:848-899) goes through
    torch + pack_planes + the CUDA fringe sum; gradients reach the Gaussian widths."""
    g = torch.Generator().manual_seed(21)
    freqs = torch.linspace(120e6, 180e6, 20, dtype=torch.float64, device=DEV)
    ants, array = _small_array(DEV, freqs)
    bls = workloads.all_cross_bls(ants)
    Ns = 90
    ra = torch.rand(Ns, generator=g, dtype=torch.float64) * 360
    dec = torch.rand(Ns, generator=g, dtype=torch.float64) * 80 - 70
    sp = torch.rand(1, 1, 20, Ns, generator=g, dtype=torch.float64).to(dtype)
    sky = ba.sky_model.PointSky(sp.to(DEV), torch.stack([ra, dec]).to(DEV),
                                R=ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='channel',
                                                                device=DEV), parameter=True)
    bp = torch.tensor([0.3, 0.4], dtype=dtype).reshape(1, 1, 1, 1, 2).repeat(1, 1, 1, 20, 1)
    beam = ba.beam_model.PixelBeam(bp.to(DEV), freqs, R=ba.beam_model.GaussResponse(powerbeam=True),
                                   pol='e', powerbeam=True, fov=180, parameter=True)
    rime = ba.RIME(sky, ba.telescope_model.TelescopeModel(workloads.LOCATION, device=DEV), beam,
                   array, bls, np.linspace(2458148.15, 2458148.2, 2), freqs, device=DEV)
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
    # oracle
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(rime)]
    spo = sp.double().clone().requires_grad_(True)
    bpo = bp.double().clone().requires_grad_(True)
    blvecs = orc.get_blvecs(array.antvecs.cpu().double(), ants, bls)
    Vo = orc.rime_forward(spo, zenaz, lambda z, a: orc.gauss_response(bpo, z, a),
                          bls, blvecs, freqs.cpu(), fov=180.0)
    (Vo.real ** 2 + Vo.imag ** 2).sum().backward()
    tag = "gauss_generic/%s" % str(dtype)[6:]
    assert relmax(V, Vo, tag + "/V") < TOL[dtype]
    assert relmax(sky.params.grad, spo.grad, tag + "/dsky") < 5 * TOL[dtype]
    assert relmax(beam.params.grad, bpo.grad, tag + "/dbeam") < 5 * TOL[dtype]


def test_healpix_pixel_beam_and_composite_sky():
    """PixelResponse with pixtype='healpix' (RING bilinear weights; parity unpinned, checked
    against the oracle's independent restatement) and a two-component CompositeModel whose
    visibilities RIME sums (the reference's own sum raises, rime_model.py:377)."""
    dtype = torch.float64
    g = torch.Generator().manual_seed(22)
    freqs = torch.linspace(120e6, 180e6, 6, dtype=torch.float64, device=DEV)
    ants, array = _small_array(DEV, freqs)
    bls = workloads.all_cross_bls(ants)[:9]
    nside = 8
    theta, phi = ba.healpix.pix2ang(nside)
    bmap = torch.as_tensor(np.exp(-0.5 * (theta / 0.5) ** 2))[None, None, None, None, :].repeat(
        1, 1, 1, 6, 1).to(dtype)
    R = ba.beam_model.PixelResponse(freqs, 'healpix', nside=nside, freq_mode='channel',
                                    powerbeam=True, device=DEV)
    beam = ba.beam_model.PixelBeam(bmap.to(DEV), freqs, R=R, pol='e', powerbeam=True, fov=180,
                                   parameter=True)
    comps, refs = {}, []
    for name, Ns in (("a", 40), ("b", 25)):
        ra = torch.rand(Ns, generator=g, dtype=torch.float64) * 360
        dec = torch.rand(Ns, generator=g, dtype=torch.float64) * 80 - 70
        sp = torch.rand(1, 1, 6, Ns, generator=g, dtype=torch.float64)
        comps[name] = ba.sky_model.PointSky(
            sp.to(DEV), torch.stack([ra, dec]).to(DEV), name=name,
            R=ba.sky_model.PointSkyResponse(freqs, freq_mode='channel', device=DEV), parameter=True)
        refs.append((name, Ns, ra, dec, sp))
    sky = ba.sky_model.CompositeModel(comps)
    tel = ba.telescope_model.TelescopeModel(workloads.LOCATION, device=DEV)
    times = np.linspace(2458148.15, 2458148.2, 2)
    rime = ba.RIME(sky, tel, beam, array, bls, times, freqs, device=DEV)
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
    # oracle: sum of the two components, healpix weights from the oracle's own restatement
    blvecs = orc.get_blvecs(array.antvecs.cpu().double(), ants, bls)
    bo = bmap.double().clone().requires_grad_(True)
    Vo = 0
    for name, Ns, ra, dec, sp in refs:
        zenaz = [tuple(x.cpu().double() for x in tel.conv_cache[(name, Ns, t)]) for t in times]

        def beam_fn(z, a):
            inds, wgts = orc.healpix_interp_weights(nside, (z * orc.D2R).numpy(), (a * orc.D2R).numpy())
            return orc.interp_map(bo.abs(), inds, wgts)
        Vo = Vo + orc.rime_forward(sp.double(), zenaz, beam_fn, bls, blvecs, freqs.cpu(), fov=180.0)
    (Vo.real ** 2 + Vo.imag ** 2).sum().backward()
    assert relmax(V, Vo, "healpix_composite/float64/V") < 1e-10
    assert relmax(beam.params.grad, bo.grad, "healpix_composite/float64/dbeam") < 1e-9


def test_float32_frequency_grid_takes_exact_path():
    """A frequency grid rounded to float32 (+-8 Hz jitter) is not uniform enough for the rotation
    recurrence on long baselines: RIME must detect it and evaluate every channel directly, with
    the same accuracy."""
    rime = workloads.point_airy(400, 96, 2, DEV, torch.float32, layout='hera350', bls='all')
    rime.setup_sim_bls(rime.sim_bls[::400])
    f32grid = torch.linspace(100e6, 200e6, 96, dtype=torch.float32).double().to(DEV)
    rime.array.set_freqs(f32grid)
    rime.beam.freqs = f32grid
    rime.sky.R.freqs = f32grid.float()
    rime.clear_geometry_cache()
    rime._bl_meta = {}
    with torch.no_grad():
        V = rime().data
    assert list(rime._bl_meta.values()) == [False]
    Vo = _oracle_of_workload(rime, 'airy', list(range(len(rime.sim_bls))), slice(None))
    assert relmax(V, Vo, "f32grid_exact_path/V") < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_edge_cases_empty_single_and_degenerate(dtype):
    """Empty and degenerate shapes the reference handles implicitly: no source above the horizon
    (every time empty), a single frequency channel, a single baseline, an autocorrelation only."""
    g = torch.Generator().manual_seed(31)
    # (a) one channel, one (zero-length) baseline, ragged tiny source counts
    zen = [torch.tensor([10.0, 50.0, 80.0], dtype=torch.float64), torch.zeros(0, dtype=torch.float64)]
    az = [torch.tensor([0.0, 120.0, 300.0], dtype=torch.float64), torch.zeros(0, dtype=torch.float64)]
    geom = ops.Geometry([z.to(DEV) for z in zen], [a.to(DEV) for a in az], DEV)
    freqs = torch.tensor([150e6], dtype=torch.float64)
    X = [torch.rand(1, 1, 3, generator=g, dtype=torch.float64), torch.zeros(1, 1, 0, dtype=torch.float64)]
    for blv in (torch.zeros(1, 3, dtype=torch.float64), torch.tensor([[14.6, 0.0, 0.0]], dtype=torch.float64)):
        A = ops.pack_planes(geom, [x.to(device=DEV, dtype=dtype) for x in X])
        V = ops.fringe_sum(A, blv.to(DEV), geom, freqs.to(DEV), 1)
        Vo = _oracle_fringe_sum([X[0]], zen[:1], az[:1], blv, freqs)
        assert V.shape == (1, 1, 2, 1)
        assert relmax(V[:, :, :1], Vo) < TOL[dtype]
        assert float(V[:, :, 1].abs().max()) == 0.0          # the empty time contributes exactly 0
    # (b) every time empty: S == 0
    geom0 = ops.Geometry([torch.zeros(0, dtype=torch.float64, device=DEV)] * 2,
                         [torch.zeros(0, dtype=torch.float64, device=DEV)] * 2, DEV)
    A0 = ops.pack_planes(geom0, [torch.zeros(1, 4, 0, dtype=dtype, device=DEV, requires_grad=True)] * 2)
    b0 = torch.rand(5, 3, generator=g, dtype=torch.float64).to(DEV).requires_grad_(True)
    f4 = torch.linspace(1e8, 2e8, 4, dtype=torch.float64, device=DEV)
    V0 = ops.fringe_sum(A0, b0, geom0, f4, 4)
    assert V0.shape == (1, 5, 2, 4) and float(V0.abs().max()) == 0.0
    V0.real.sum().backward()
    assert float(b0.grad.abs().max()) == 0.0
    # (c) RIME with a field of view so narrow that nothing is ever inside it
    rime = workloads.point_airy(50, 8, 2, DEV, dtype)
    rime.beam.fov = 1e-6
    rime.clear_geometry_cache()
    with torch.no_grad():
        Vn = rime().data
    assert Vn.shape == (1, 1, 666, 2, 8) and float(Vn.abs().max()) == 0.0


@pytest.mark.parametrize("name", ["rime_point_airy", "rime_pixel_interp"])
def test_graphed_step_replays_the_eager_step(name):
    """rime_model.GraphedStep: forward + loss + backward captured once into a CUDA graph; replays
    reproduce the eager step bit for bit, also after an in-place parameter update."""
    if DOUBLE:
        pytest.skip("CUDA graphs need the GPU")
    g = oc.load(name)
    build, gkeys = mc.CASES[name]
    rime, leaves = build(g, DEV, torch.float32)
    params = [leaves[k] for k in gkeys]
    G = torch.as_tensor(g["G"]).to(device=DEV, dtype=torch.complex64)
    loss_fn = lambda vd: torch.sum(G.real * vd.data.real + G.imag * vd.data.imag)

    def eager():
        for p in params:
            p.grad = None
        loss = loss_fn(rime())
        loss.backward()
        return loss.detach().clone(), [p.grad.clone() for p in params]

    l0, g0 = eager()
    step = ba.rime_model.GraphedStep(rime, loss_fn, params)
    # antenna-position gradients pass through torch's index_select backward (atomic adds: not
    # bitwise reproducible run to run, graph or not); everything else is
    def same(a, b, key):
        return relmax(a, b) < 1e-6 if key == "antvecs" else torch.equal(a, b)

    keys = list(gkeys)
    l1 = step()
    torch.cuda.synchronize()
    assert torch.equal(l1, l0)
    for p, ref, key in zip(params, g0, keys):
        assert same(p.grad, ref, key), key
    for k, gk in gkeys.items():
        assert relmax(leaves[k].grad, g[gk]) < 5e-5
    with torch.no_grad():
        params[0].mul_(1.25)
    l2 = step().clone()
    grads2 = [p.grad.clone() for p in params]
    l3, g3 = eager()
    assert torch.equal(l2, l3)
    for a, b, key in zip(grads2, g3, keys):
        assert same(a, b, key), key
    assert float((l2 - l0).abs()) > 0


@pytest.mark.skipif(DOUBLE or torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_model_on_a_non_current_device():
    """kernels launch on the device that owns the tensors and on torch's current
    stream of THAT device; per-device function attributes (dynamic shared memory) are set on every
    device a process drives.  The process's current device stays cuda:0 throughout."""
    assert torch.cuda.current_device() == 0
    dev = 'cuda:1'
    for name in ("rime_point_airy", "rime_pixel_interp"):
        g = oc.load(name)
        build, gkeys = mc.CASES[name]
        rime, leaves = build(g, dev, torch.float32)
        V = rime().data
        assert V.device == torch.device(dev)
        assert relmax(V, g["vis"]) < TOL[torch.float32]
        G = torch.as_tensor(g["G"]).to(device=dev, dtype=V.dtype)
        torch.sum(G.real * V.real + G.imag * V.imag).backward()
        for k, gk in gkeys.items():
            assert relmax(leaves[k].grad, g[gk]) < 5 * TOL[torch.float32], (k, gk)
    # kernels that need more than 48 KB of dynamic shared memory, first use on cuda:1
    rime = workloads.pixel_interp(16, 96, 1, dev, torch.float32, antpos_param=True)
    assert rime._tc_tiling(torch.device(dev)) is not None
    V = rime().data
    (V.real ** 2 + V.imag ** 2).sum().backward()
    ref = workloads.pixel_interp(16, 96, 1, 'cuda:0', torch.float32, antpos_param=True)
    V0 = ref().data
    assert relmax(V, V0.to(dev)) < 1e-6
    Y = torch.randn(40, 300, dtype=torch.complex64, device=dev)
    p = torch.randn(6, 40, dtype=torch.complex64, device=dev)
    out = ops.alm_forward(p, ops.AlmPlan(Y))
    assert relmax(out, p.cpu().to(torch.complex128) @ Y.cpu().to(torch.complex128)) < 1e-5
    assert torch.cuda.current_device() == 0


@pytest.mark.skipif(DOUBLE, reason="compares the CUDA kernel with the dense restatement")
@pytest.mark.parametrize("lower_only", [False, True])
def test_cotangent_pack_kernel_equals_dense_construction(lower_only):
    """b200rime_tc_pack_cotangent_f32 against the dense torch construction it replaced (kept as
    the CPU test double): bit-exact float16 operands, for a baseline list with both orientations,
    auto-correlations, missing pairs, ragged antenna / channel counts and a time sub-range."""
    from tests import cpu_double
    rng = np.random.default_rng(5)
    na, nt_all, nf = 150, 3, 70
    pairs = [(i, j) for i in range(na) for j in range(i, na) if rng.random() < 0.6]
    pairs = [(j, i) if rng.random() < 0.3 else (i, j) for i, j in pairs]
    i_idx, j_idx = [p[0] for p in pairs], [p[1] for p in pairs]
    tc = ops.TcTiling(i_idx, j_idx, na, 'cuda')
    assert tc.usable and len(tc.auto) > 0 and len(tc.igt) > 0
    gen = torch.Generator().manual_seed(8)
    G = torch.complex(torch.randn(len(pairs), nt_all, nf, generator=gen, dtype=torch.float32),
                      torch.randn(len(pairs), nt_all, nf, generator=gen, dtype=torch.float32)).cuda()
    nfp = ops.nchunks(nf, torch.float32) * _lib.KC["f32"]
    Gs = G[:, 1:3]                                           # a time sub-range: strided view
    Hq, hscale = tc.cotangent_operand(Gs, nfp, lower_only=lower_only)
    ref = torch.empty(Hq.shape, dtype=torch.float16)
    Gc = Gs.cpu().contiguous()
    cpu_double.tc_pack_cotangent("f32", Gc, Gc.stride(0), tc.pair_bl.cpu(), tc.ldp, 2, nf, na, tc.nm_pad,
                                 int(lower_only), hscale.cpu(), ref)
    assert torch.equal(Hq.cpu().view(torch.int16), ref.view(torch.int16))
    assert float(Hq.abs().max()) >= 2.0 ** 13          # the scale uses the float16 range


def test_c5_minibatch_grid_on_tensor_core_kernels_vs_oracle():
    """BASELINE config 5 family (reduced): HERA-350, all cross baselines in two block-aligned
    baseline groups x single-time groups = the reference's minibatch grid (rime_model.py:253-289).
    Every minibatch stays on the tensor-core kernels; visibilities of each minibatch match the
    fp64 oracle on a baseline / channel subset; gradients accumulated over the grid equal the
    gradients of the same model run as ONE batch (minibatch invariance, reference
    tests/test_rime.py:49-51) and match oracle autograd for a sparse cotangent."""
    if DOUBLE:
        pytest.skip("needs the GPU")
    nf, nt = 96, 2
    rime = workloads.pixel_interp(16, nf, nt, DEV, torch.float32, bl_groups=True, time_groups=True)
    one = workloads.pixel_interp(16, nf, nt, DEV, torch.float32)
    dev = torch.device(DEV, torch.cuda.current_device())
    assert rime.Nbatch == 2 * nt and one.Nbatch == 1
    nbl_all = len(one.sim_bls)
    row = {bl: k for k, bl in enumerate(one.sim_bls)}
    gen = torch.Generator().manual_seed(3)
    V1 = one().data
    G = torch.zeros(V1.shape, dtype=torch.complex64, device=V1.device)
    gb = sorted(np.random.default_rng(1).choice(nbl_all, 6, replace=False).tolist())
    gf = list(range(3, nf, 16))
    gbt, gft = torch.as_tensor(gb, device=V1.device), torch.as_tensor(gf, device=V1.device)
    shp = (1, 1, len(gb), nt, len(gf))
    G_sub = torch.complex(torch.randn(shp, generator=gen, dtype=torch.float64),
                          torch.randn(shp, generator=gen, dtype=torch.float64))
    G[0, 0, gbt[:, None, None], torch.arange(nt, device=V1.device)[None, :, None], gft[None, None, :]] = \
        G_sub[0, 0].to(V1.device, torch.complex64)
    torch.sum(G.real * V1.real + G.imag * V1.imag).backward()
    Vgrid = torch.zeros_like(V1)
    seen = 0
    for b in range(rime.Nbatch):
        rime.batch_idx = b
        assert rime._tc_tiling(dev) is not None, "minibatch %d left the tensor-core kernels" % b
        vd = rime()
        rows = torch.as_tensor([row[bl] for bl in rime.sim_bls], device=V1.device)
        tsel = [int(np.argmin(np.abs(np.asarray(one.sim_times) - t))) for t in rime.sim_times]
        assert len(tsel) == 1
        Vgrid[:, :, rows, tsel[0]] = vd.data[:, :, :, 0].detach()
        Gb = G[:, :, rows][:, :, :, tsel]
        torch.sum(Gb.real * vd.data.real + Gb.imag * vd.data.imag).backward()
        seen += len(rows)
    assert seen == nbl_all * nt
    assert relmax(Vgrid, V1, "c5_grid/V_vs_one_batch") < 2e-6
    assert relmax(rime.sky.params.grad, one.sky.params.grad, "c5_grid/dsky_vs_one_batch") < 5e-6
    assert relmax(rime.beam.params.grad, one.beam.params.grad, "c5_grid/dbeam_vs_one_batch") < 5e-6
    # fp64 oracle on the cotangent's support
    zenaz = [(za[0].cpu().double(), za[1].cpu().double()) for za in workloads.zenaz_of(one)]
    fi = torch.as_tensor(gf)
    freqs = one.array.freqs.detach().cpu().double()[fi]
    antvecs = one.array.antvecs.detach().cpu().double()
    sp = one.sky.params.detach().cpu().double()[:, :, fi].requires_grad_(True)
    bp = one.beam.params.detach().cpu().double()[:, :, :, fi].requires_grad_(True)
    bls = [one.sim_bls[i] for i in gb]
    blvecs = orc.get_blvecs(antvecs, one.array.ants, bls)
    bmap = orc.pixel_response_forward(bp, powerbeam=True)
    tg, pg = one.beam.R.theta_grid.cpu().double(), one.beam.R.phi_grid.cpu().double()

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
        return orc.interp_map(bmap, inds, wgts)

    Vo = orc.rime_forward(sp * float(one.sky.px_area), zenaz, beam_fn, bls, blvecs, freqs, fov=180.0)
    oc.real_loss(Vo, G_sub).backward()
    sub = Vgrid[0, 0][gbt][:, :, gft].cpu().to(torch.complex128)
    err = float((sub - Vo[0, 0].detach()).abs().max()) / float(V1.abs().max())
    ERRLOG["c5_grid/V_vs_oracle"] = err
    assert err < 1e-5
    assert relmax(rime.sky.params.grad[:, :, gf], sp.grad, "c5_grid/dsky_vs_oracle") < 5e-5
    assert relmax(rime.beam.params.grad[:, :, :, gf], bp.grad, "c5_grid/dbeam_vs_oracle") < 5e-5
