"""
Drop-in check with the REFERENCE's own component objects (build container only: needs
/root/reference; skipped elsewhere).  sky / telescope / beam / array are built with the
unmodified reference classes and handed to BOTH ``bayeslim.rime_model.RIME`` and
``bayeslim_b200.rime_model.RIME``; outputs and gradients must agree.  No GPU here, so the CUDA
kernels are replaced by the torch test double (tests/cpu_double.py): this pins the
duck-typed host logic against real reference objects, the kernels themselves are pinned on
the B200 by tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
import torch

from tests.golden import _refshim

pytestmark = pytest.mark.skipif(not _refshim.available(), reason="reference tree not present")

torch.set_default_dtype(torch.float64)
LOC = (21.42827, -30.72148, 1051.7)


def _inject(rime, name, ra, dec):
    from oracle import rime_oracle as orc
    for t in rime.sim_times:
        zen, az = orc.eq2top_synth(t, ra, dec, lat=LOC[1])
        rime.telescope.conv_cache[(name, len(ra), t)] = torch.stack(
            [torch.as_tensor(zen), torch.as_tensor(az)])


def _models(ref, kind):
    rng = np.random.default_rng(7)
    freqs = torch.linspace(120e6, 180e6, 7)
    times = np.linspace(2458148.15, 2458148.25, 3)
    ants, vecs = ref.utils._make_hex(2, D=14.6)
    array = ref.telescope_model.ArrayModel(dict(zip(ants, vecs)), freqs=freqs)
    array.set_param('antvecs')
    bls = [(ants[i], ants[j]) for i in range(len(ants)) for j in range(i + 1, len(ants))][:12]
    Ns = 45
    ra, dec = rng.uniform(0, 360, Ns), rng.uniform(-75, 15, Ns)
    angs = torch.as_tensor(np.stack([ra, dec]))
    params = torch.as_tensor(np.abs(rng.normal(size=(1, 1, len(freqs), Ns))))
    sky = ref.sky_model.PointSky(params, angs, parameter=True,
                                 R=ref.sky_model.PointSkyResponse(freqs, freq_mode='channel'))
    if kind == 'airy':
        beam = ref.beam_model.PixelBeam(torch.tensor([14.0, 12.0]).reshape(1, 1, 1, 1, 2), freqs,
                                        R=ref.beam_model.AiryResponse(powerbeam=True), pol='e',
                                        powerbeam=True, fov=160, parameter=True)
    elif kind == 'interp':
        theta = torch.arange(0, 90.1, 5.0)
        phi = torch.arange(0, 360, 10.0)
        b_phi, b_theta = torch.meshgrid(phi, theta, indexing='xy')
        m = ref.beam_model.airy_disk(b_theta.ravel() * ref.D2R, b_phi.ravel() * ref.D2R, 12.0, freqs)
        R = ref.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta_grid=theta,
                                         phi_grid=phi, powerbeam=True)
        beam = ref.beam_model.PixelBeam(torch.as_tensor(m[None, None, None]), freqs, R=R, pol='e',
                                        powerbeam=True, fov=180, parameter=True)
    else:   # gauss: no fused builder -> generic route through the reference's own response
        beam = ref.beam_model.PixelBeam(torch.tensor([0.3, 0.4]).reshape(1, 1, 1, 1, 2).repeat(
            1, 1, 1, len(freqs), 1), freqs, R=ref.beam_model.GaussResponse(powerbeam=True), pol='e',
            powerbeam=True, fov=180, parameter=True)
    tel = ref.telescope_model.TelescopeModel(LOC)
    return sky, tel, beam, array, bls, times, freqs, ra, dec


@pytest.mark.parametrize("kind", ["airy", "interp", "gauss"])
def test_reference_objects_through_b200_rime(kind):
    ref = _refshim.load()
    import bayeslim_b200 as ba
    from tests.cpu_double import emulated_kernels
    sky, tel, beam, array, bls, times, freqs, ra, dec = _models(ref, kind)
    gen = torch.Generator().manual_seed(3)

    rime_ref = ref.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    _inject(rime_ref, sky.name, ra, dec)
    vd_ref = rime_ref()
    G = torch.complex(torch.randn(vd_ref.data.shape, generator=gen),
                      torch.randn(vd_ref.data.shape, generator=gen))
    torch.sum(G.real * vd_ref.data.real + G.imag * vd_ref.data.imag).backward()
    gref = [p.grad.clone() for p in (sky.params, beam.params, array.antvecs)]
    for p in (sky.params, beam.params, array.antvecs):
        p.grad = None

    with emulated_kernels() as calls:
        rime = ba.RIME(sky, tel, beam, array, bls, times, freqs)     # SAME reference objects
        vd = rime()
        torch.sum(G.real * vd.data.real + G.imag * vd.data.imag).backward()
    assert vd.data.shape == vd_ref.data.shape
    assert float((vd.data - vd_ref.data).abs().max() / vd_ref.data.abs().max()) < 1e-11
    assert vd.bls == vd_ref.bls and np.allclose(vd.times, vd_ref.times) and vd.pol == vd_ref.pol
    for p, g0 in zip((sky.params, beam.params, array.antvecs), gref):
        assert float((p.grad - g0).abs().max() / g0.abs().max()) < 1e-9
    expected = {"airy": "build_airy", "interp": "build_interp_t", "gauss": "pack"}[kind]
    assert expected in calls and "fringe_sum_fwd" in calls
