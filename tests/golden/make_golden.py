"""
Generate golden input/output vectors from the UNMODIFIED reference.

Run by hand in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Each fixture is a small .npz holding the inputs needed to rebuild the model with
``bayeslim_b200``'s own classes (or to feed ``oracle/rime_oracle.py``) and the
reference's outputs: visibilities and, where stated, autograd gradients obtained
with a fixed random cotangent G (loss = Re sum(conj(G) * V), so dL/dV = G in the
PyTorch convention).  The reference cannot travel to the GPU box; these files can.

Geometry (zen, az per time) comes from oracle.eq2top_synth and is injected through
``telescope.conv_cache`` (key format of rime_model.py:345), because astropy is
not installed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _refshim  # noqa: E402
from oracle import rime_oracle as orc  # noqa: E402

torch.set_default_dtype(torch.float64)
ba = _refshim.load()

LOC = (21.42827, -30.72148, 1051.7)   # tests/test_telescope.py:13


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f kB)" % (path, os.path.getsize(path) / 1e3))


def inject_geometry(rime, sky_name, ra, dec, times):
    zen_az = []
    for t in times:
        zen, az = orc.eq2top_synth(t, ra, dec, lat=LOC[1])
        za = torch.stack([torch.as_tensor(zen), torch.as_tensor(az)])
        rime.telescope.conv_cache[(sky_name, len(ra), t)] = za
        zen_az.append(za.numpy())
    return np.asarray(zen_az)


def cotangent(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(shape, generator=g), torch.randn(shape, generator=g))


def backward_with(V, G):
    loss = torch.sum(G.real * V.real + G.imag * V.imag)
    loss.backward()
    return loss.detach()


# ------------------------------------------------------------------ unit pieces
def gold_fringe():
    rng = np.random.default_rng(1)
    ants, vecs = ba.utils._make_hex(3, D=15)
    antpos = dict(zip(ants, vecs))
    freqs = torch.linspace(120e6, 130e6, 10)
    array = ba.telescope_model.ArrayModel(antpos, freqs=freqs)
    bls = [(0, 1), (0, 0), (0, 18), (3, 11), (7, 2)]
    blvecs = array.get_blvecs(bls)
    zen = torch.as_tensor(np.concatenate([[0.0, 90.0, 120.0], rng.uniform(0, 100, 37)]))
    az = torch.as_tensor(np.concatenate([[0.0, 270.0, 45.0], rng.uniform(0, 360, 37)]))
    fr = array.gen_fringe(blvecs, zen, az)
    array.clear_cache()
    frc = array.gen_fringe(blvecs, zen, az, conj=True)
    save("fringe", antvecs=vecs, ants=ants, bls=bls, blvecs=blvecs, freqs=freqs, zen=zen, az=az,
         fringe=fr, fringe_conj=frc)


def gold_airy():
    rng = np.random.default_rng(2)
    freqs = torch.linspace(100e6, 200e6, 7)
    zen = torch.as_tensor(np.concatenate([[0.0, 1e-9, 90.0, 95.0], rng.uniform(0, 100, 60)]))
    az = torch.as_tensor(np.concatenate([[0.0, 10.0, 180.0, 33.0], rng.uniform(0, 360, 60)]))
    out = {}
    for nm, p, pb in [("sym_power", torch.ones(1, 1, 1, 1, 1) * 14.0, True),
                      ("asym_power", torch.tensor([14.0, 12.5]).reshape(1, 1, 1, 1, 2), True),
                      ("asym_volt", torch.tensor([14.0, 12.5]).reshape(1, 1, 1, 1, 2), False),
                      ("twopol_power", torch.tensor([[14.0, 12.5], [13.0, 15.0]]).reshape(2, 1, 1, 1, 2), True)]:
        R = ba.beam_model.AiryResponse(powerbeam=pb)
        out["params_" + nm] = p
        out["beam_" + nm] = R(p, zen, az, freqs)
    R = ba.beam_model.AiryResponse(powerbeam=True, freq_ratio=1.1)
    out["beam_ratio"] = R(out["params_sym_power"], zen, az, freqs)
    save("airy", freqs=freqs, zen=zen, az=az, **out)


def gold_rect_interp():
    rng = np.random.default_rng(3)
    theta_grid = torch.arange(0, 90.1, 2.0)
    phi_grid = torch.arange(0, 360, 4.0)
    zen = torch.as_tensor(np.concatenate([[0.0, 2.0, 89.99, 90.7, 45.0, 13.3, 88.0],
                                          rng.uniform(0, 92, 80)]))
    az = torch.as_tensor(np.concatenate([[0.0, 359.9, 357.0, 1.0, 4.0, 180.0, 356.0],
                                         rng.uniform(0, 360, 80)]))
    m = torch.as_tensor(rng.normal(size=(2, 3, len(theta_grid) * len(phi_grid))))
    out = {}
    for mode in ['nearest', 'linear', 'quadratic', 'cubic', 'linear,quadratic']:
        P = ba.utils.PixInterp('rect', interp_mode=mode, theta_grid=theta_grid, phi_grid=phi_grid)
        inds, wgts = P.get_interp(zen, az)
        key = mode.replace(',', '_')
        out["inds_" + key] = inds
        out["wgts_" + key] = wgts
        out["interp_" + key] = P.interp(m, zen, az)
    save("rect_interp", theta_grid=theta_grid, phi_grid=phi_grid, zen=zen, az=az, m=m, **out)


# ------------------------------------------------------------------ full RIME cases
def hera_array(N, freqs, D=14.6, set_param=False):
    ants, vecs = ba.utils._make_hex(N, D=D)
    antpos = dict(zip(ants, vecs))
    array = ba.telescope_model.ArrayModel(antpos, freqs=freqs)
    if set_param:
        array.set_param('antvecs')
    return ants, vecs, array


def gold_rime_point_airy():
    """C1-like: hex-7, all cross baselines, power-law point sources, Airy power beam;
    gradients to sky params, Airy diameter (reference's truncated gradient) and antvecs."""
    rng = np.random.default_rng(10)
    freqs = torch.linspace(100e6, 200e6, 12)
    times = np.linspace(2458148.15, 2458148.25, 3)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(ants[i], ants[j]) for i in range(len(ants)) for j in range(i + 1, len(ants))]
    Ns = 60
    ra = rng.uniform(0, 360, Ns)
    dec = np.degrees(np.arcsin(rng.uniform(-1, np.sin(np.radians(29)), Ns)))
    angs = torch.as_tensor(np.stack([ra, dec]))
    params = torch.zeros(1, 1, 2, Ns)
    params[0, 0, 0] = torch.as_tensor(np.exp(rng.normal(size=Ns)))
    params[0, 0, 1] = torch.as_tensor(rng.normal(-0.8, 0.2, Ns))
    R = ba.sky_model.PointSkyResponse(freqs, freq_mode='powerlaw', f0=150e6)
    sky = ba.sky_model.PointSky(params.clone(), angs, R=R, parameter=True)
    bp = torch.ones(1, 1, 1, 1, 1) * 14.0
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=ba.beam_model.AiryResponse(powerbeam=True),
                                   pol='e', powerbeam=True, fov=180, parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 100)
    backward_with(vd.data, G)
    save("rime_point_airy", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times,
         ra=ra, dec=dec, zen_az=zen_az, sky_params=params, f0=150e6, beam_params=bp,
         vis=vd.data, G=G, grad_sky=sky.params.grad, grad_beam_truncated=beam.params.grad,
         grad_antvecs=array.antvecs.grad, fov=180.0)


def gold_rime_airy_brute():
    """AiryResponse(brute_force=True): the differentiable trapezoid J1 (special.py:498-535) through
    the RIME; the gradient to the Airy diameter is the FULL one here."""
    rng = np.random.default_rng(12)
    freqs = torch.linspace(100e6, 200e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs)
    bls = [(ants[i], ants[j]) for i in range(len(ants)) for j in range(i + 1, len(ants))]
    Ns = 40
    ra = rng.uniform(0, 360, Ns)
    dec = np.degrees(np.arcsin(rng.uniform(-1, np.sin(np.radians(29)), Ns)))
    angs = torch.as_tensor(np.stack([ra, dec]))
    params = torch.zeros(1, 1, 2, Ns)
    params[0, 0, 0] = torch.as_tensor(np.exp(rng.normal(size=Ns)))
    params[0, 0, 1] = torch.as_tensor(rng.normal(-0.8, 0.2, Ns))
    R = ba.sky_model.PointSkyResponse(freqs, freq_mode='powerlaw', f0=150e6)
    sky = ba.sky_model.PointSky(params.clone(), angs, R=R, parameter=True)
    bp = torch.ones(1, 1, 1, 1, 1) * 14.0
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=True, brute_force=True, Ntau=64),
                                   pol='e', powerbeam=True, fov=180, parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 101)
    backward_with(vd.data, G)
    save("rime_airy_brute", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times,
         ra=ra, dec=dec, zen_az=zen_az, sky_params=params, f0=150e6, beam_params=bp, Ntau=64,
         vis=vd.data, G=G, grad_sky=sky.params.grad, grad_beam=beam.params.grad, fov=180.0)


def gold_alm():
    """AlmModel.forward_alm (sph_harm.py:1289-1373): full and separable transforms, complex and
    real output, alm_mult; Ylm from the reference's gen_sph2pix(high_prec=False) (lmax 8, m >= 0);
    gradients to the coefficients."""
    rng = np.random.default_rng(31)
    l, m = ba.sph_harm.gen_lm(8, real_field=True)
    theta = np.degrees(np.arccos(rng.uniform(0, 1, 150)))
    phi = rng.uniform(0, 360, 150)
    out = dict(l=l, m=m, theta=theta, phi=phi)
    g = torch.Generator().manual_seed(5)
    params = torch.complex(torch.randn(2, 3, len(l), generator=g), torch.randn(2, 3, len(l), generator=g))
    out["params"] = params
    for tag, real in (("complex", False), ("real", True)):
        A = ba.sph_harm.AlmModel(l, m, real_output=real, default_kw=dict(high_prec=False))
        A.setup_Ylm(theta, phi, generate=True)
        out["Ylm"], out["alm_mult"] = A.Ylm, A.alm_mult
        p = params.clone().requires_grad_()
        y = A(p)
        G = cotangent(y.shape, 7) if not real else cotangent(y.shape, 7).real
        (backward_with(y, G) if not real else torch.sum(G * y).backward())
        out["out_" + tag], out["G_" + tag], out["grad_" + tag] = y, G, p.grad
    # separable grid
    tg, pg = np.linspace(2, 88, 7), np.linspace(0, 330, 12)
    A = ba.sph_harm.AlmModel(l, m, real_output=True, default_kw=dict(high_prec=False))
    A.setup_Ylm(tg, pg, generate=True, separable=True)
    p = params.clone().requires_grad_()
    y = A(p)
    G = cotangent(y.shape, 9).real
    torch.sum(G * y).backward()
    out.update(theta_grid=tg, phi_grid=pg, Theta=A.Ylm[0], Phi=A.Ylm[1], out_sep=y, G_sep=G,
               grad_sep=p.grad)
    save("alm_forward", **out)


def gold_rime_ylm():
    """Spherical-harmonic beam through the RIME (beam_model.py:1019-1267): YlmResponse in
    'interpolate' mode on a rect grid (a_lm -> beam_cache once per forward, bilinear
    interpolation at the sources), complex coefficients in 2-real form (comp_params), a_lm as
    perturbation about an Airy beam0, HEALPix nside-4 PixelSky; gradients to sky, a_lm, antvecs."""
    rng = np.random.default_rng(41)
    freqs = torch.linspace(100e6, 200e6, 7)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(0, 1), (0, 2), (0, 3), (1, 5), (2, 6), (0, 6), (3, 4), (1, 1)]
    ra, dec, px_area, sparams = healpix_sky(4, freqs, rng)
    angs = torch.as_tensor(np.stack([ra, dec]))
    sky = ba.sky_model.PixelSky(sparams.clone(), angs, px_area,
                                R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)
    theta, phi, b_theta, b_phi, airy = rect_airy_beam(freqs, 5.0, 10.0)
    l, m = ba.sph_harm.gen_lm(6, real_field=True)
    beam0 = torch.as_tensor(airy[None, None, None, :, :]).clone()
    R = ba.beam_model.YlmResponse(l, m, freqs, pixtype='rect', mode='interpolate',
                                  interp_mode='linear', theta=b_theta, phi=b_phi,
                                  theta_grid=theta, phi_grid=phi, powerbeam=True,
                                  comp_params=True, beam0=beam0, freq_mode='channel',
                                  Ylm_kwargs=dict(high_prec=False))
    R.setup_Ylm(b_theta, b_phi, generate=True)
    Ylm, alm_mult = R.Ylm, R.alm_mult
    g = torch.Generator().manual_seed(6)
    bp = 0.02 * torch.randn(1, 1, 1, len(freqs), len(l), 2, generator=g) \
        / (1.0 + torch.as_tensor(l, dtype=torch.float64))[:, None]
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=R, pol='e', powerbeam=True, fov=180,
                                   parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 103)
    backward_with(vd.data, G)
    bc = ba.beam_model.YlmResponse.forward(R, bp, b_theta, b_phi).detach()
    save("rime_ylm", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra, dec=dec,
         zen_az=zen_az, sky_params=sparams, px_area=px_area, beam_params=bp, beam0=beam0, l=l, m=m,
         theta_grid=theta, phi_grid=phi, Ylm_sample=Ylm[:, ::23], alm_mult=alm_mult,
         beam_cache=bc, vis=vd.data, G=G, grad_sky=sky.params.grad, grad_beam=beam.params.grad,
         grad_antvecs=array.antvecs.grad, fov=180.0)


def gold_rime_alm_sky():
    """Spherical-harmonic sky (sky_model.py:510-732, spatial_mode='alm'): PixelSky whose pixel
    intensities are a_lm coefficients forward-modelled by an AlmModel (sph_harm.py:1289-1373),
    complex coefficients in 2-real form, sky0 offset; Airy beam; gradients to the a_lm and antvecs."""
    rng = np.random.default_rng(51)
    freqs = torch.linspace(100e6, 200e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(0, 1), (0, 2), (0, 3), (1, 5), (2, 6), (0, 6), (3, 4)]
    ra, dec, px_area, sparams = healpix_sky(4, freqs, rng)
    angs = torch.as_tensor(np.stack([ra, dec]))
    l, m = ba.sph_harm.gen_lm(5, real_field=True)
    A = ba.sph_harm.AlmModel(l, m, real_output=False, default_kw=dict(high_prec=False))
    A.setup_Ylm(90.0 - dec, ra, generate=True)
    g = torch.Generator().manual_seed(9)
    p = 0.3 * torch.randn(1, 1, len(freqs), len(l), 2, generator=g) \
        / (1.0 + torch.as_tensor(l, dtype=torch.float64))[:, None]
    R = ba.sky_model.PixelSkyResponse(freqs, comp_params=True, spatial_mode='alm', spat_LM=A,
                                      freq_mode='channel', sky0=sparams.clone())
    sky = ba.sky_model.PixelSky(p.clone(), angs, px_area, R=R, parameter=True)
    beam = ba.beam_model.PixelBeam(torch.ones(1, 1, 1, 1, 1) * 14.0, freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                   powerbeam=True, fov=180, parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 105)
    backward_with(vd.data, G)
    save("rime_alm_sky", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra, dec=dec,
         zen_az=zen_az, sky_params=p, sky0=sparams, px_area=px_area, l=l, m=m,
         Ylm_sample=A.Ylm[:, ::17], alm_mult=A.alm_mult, vis=vd.data, G=G,
         grad_sky=sky.params.grad, grad_antvecs=array.antvecs.grad, fov=180.0)


def healpix_sky(nside, freqs, rng, dec_max=59.27852):
    theta, phi = orc.healpix_pix2ang(nside)
    dec = np.pi / 2 - theta
    cut = dec < dec_max * np.pi / 180
    ra_deg, dec_deg = np.degrees(phi[cut]), np.degrees(dec[cut])
    px_area = orc.healpix_pixarea(nside)
    params = np.abs(rng.normal(size=(1, 1, len(freqs), cut.sum()))) * \
        (np.asarray(freqs)[:, None] / 150e6) ** -2.5
    return ra_deg, dec_deg, px_area, torch.as_tensor(params)


def rect_airy_beam(freqs, dth, dph, D=14.0, npol=1, nvec=1, nmodel=1):
    theta = torch.arange(0, 90.1, dth)
    phi = torch.arange(0, 360, dph)
    b_phi, b_theta = torch.meshgrid(phi, theta, indexing='xy')
    b_phi, b_theta = b_phi.ravel(), b_theta.ravel()
    airy = ba.beam_model.airy_disk(b_theta * ba.D2R, b_phi * ba.D2R, D, freqs, square=True)
    return theta, phi, b_theta, b_phi, airy


def gold_rime_pixel_interp():
    """C3-like in miniature: HEALPix nside-4 PixelSky, rect-bilinear PixelBeam (parameter),
    hex-7 unique+some baselines; grads to sky params, beam map, antvecs."""
    rng = np.random.default_rng(11)
    freqs = torch.linspace(100e6, 200e6, 9)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(0, 1), (0, 2), (0, 3), (1, 5), (2, 6), (0, 6), (3, 4), (1, 1)]
    ra, dec, px_area, sparams = healpix_sky(4, freqs, rng)
    angs = torch.as_tensor(np.stack([ra, dec]))
    sky = ba.sky_model.PixelSky(sparams.clone(), angs, px_area,
                                R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)
    theta, phi, b_theta, b_phi, airy = rect_airy_beam(freqs, 5.0, 10.0)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta=b_theta, phi=b_phi,
                                    theta_grid=theta, phi_grid=phi, freq_mode='channel',
                                    powerbeam=True, realbeam=True, log=False)
    bp = torch.as_tensor(airy[None, None, None, :, :]).clone()
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=R, pol='e', powerbeam=True, fov=180,
                                   parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 101)
    backward_with(vd.data, G)
    save("rime_pixel_interp", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times,
         ra=ra, dec=dec, zen_az=zen_az, sky_params=sparams, px_area=px_area, beam_params=bp,
         theta_grid=theta, phi_grid=phi, vis=vd.data, G=G, grad_sky=sky.params.grad,
         grad_beam=beam.params.grad, grad_antvecs=array.antvecs.grad, fov=180.0)


def gold_rime_pointing():
    """Pointing offset (beam_model.py:244-256, 1631-1678): the pixel-interp case with the beam
    tilted by (theta_x, theta_y) = (0.03, 0.05) rad, fov 150 deg so that the interpolation
    stays inside the beam grid, and the Airy beam with the same offset; values and gradients."""
    rng = np.random.default_rng(21)
    freqs = torch.linspace(100e6, 200e6, 7)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(0, 1), (0, 2), (0, 3), (1, 5), (2, 6), (0, 6), (3, 4)]
    ra, dec, px_area, sparams = healpix_sky(4, freqs, rng)
    angs = torch.as_tensor(np.stack([ra, dec]))
    offset = (0.03, 0.05)
    out = {}
    # interpolated pixel beam
    sky = ba.sky_model.PixelSky(sparams.clone(), angs, px_area,
                                R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)
    theta, phi, b_theta, b_phi, airy = rect_airy_beam(freqs, 5.0, 10.0)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta=b_theta, phi=b_phi,
                                    theta_grid=theta, phi_grid=phi, freq_mode='channel',
                                    powerbeam=True, realbeam=True, log=False)
    bp = torch.as_tensor(airy[None, None, None, :, :]).clone()
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=R, pol='e', powerbeam=True, fov=150,
                                   parameter=True, offset=offset)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 121)
    backward_with(vd.data, G)
    out.update(vis_interp=vd.data, grad_sky_interp=sky.params.grad, grad_beam_interp=beam.params.grad,
               grad_antvecs_interp=array.antvecs.grad.clone())
    # Airy beam with the same offset
    array.antvecs.grad = None
    sky2 = ba.sky_model.PixelSky(sparams.clone(), angs, px_area,
                                 R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)
    beam2 = ba.beam_model.PixelBeam(torch.ones(1, 1, 1, 1, 1) * 14.0, freqs,
                                    R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                    powerbeam=True, fov=150, parameter=False, offset=offset)
    rime2 = ba.rime_model.RIME(sky2, tel, beam2, array, bls, times, freqs)
    inject_geometry(rime2, sky2.name, ra, dec, rime2.sim_times)
    vd2 = rime2()
    backward_with(vd2.data, G)
    out.update(vis_airy=vd2.data, grad_sky_airy=sky2.params.grad)
    save("rime_pointing", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra,
         dec=dec, zen_az=zen_az, sky_params=sparams, px_area=px_area, beam_params=bp,
         theta_grid=theta, phi_grid=phi, G=G, fov=150.0, offset=np.asarray(offset), **out)


def gold_rime_batched():
    """test_RIME analogue (tests/test_rime.py:29-51): minibatched == single shot, plus values.
    Also exercises a narrower FOV (fov=120) and quadratic interpolation."""
    rng = np.random.default_rng(12)
    freqs = torch.linspace(120e6, 130e6, 10)
    times = np.linspace(2458148.10, 2458148.30, 5)
    ants, vecs, array = hera_array(3, freqs, D=15.0)
    bls = array.get_bls(uniq_bls=True, keep_autos=False)
    ra, dec, px_area, sparams = healpix_sky(4, freqs, rng)
    sparams = torch.as_tensor(rng.normal(size=tuple(sparams.shape)))
    angs = torch.as_tensor(np.stack([ra, dec]))
    sky = ba.sky_model.PixelSky(sparams.clone(), angs, px_area,
                                R=ba.sky_model.PixelSkyResponse(freqs), parameter=False)
    theta, phi, b_theta, b_phi, airy = rect_airy_beam(freqs, 2.0, 5.0, D=10.0)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='quadratic', theta=b_theta, phi=b_phi,
                                    theta_grid=theta, phi_grid=phi, freq_mode='channel',
                                    powerbeam=True, realbeam=True, log=False)
    bp = torch.as_tensor(airy[None, None, None, :, :]).clone()
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=R, pol='e', powerbeam=True, fov=120,
                                   parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    with torch.no_grad():
        vd = rime()
        rime.setup_sim_times(ba.utils.split_into_groups(torch.as_tensor(times), Nelem=2))
        vb = rime.run_batches()
    assert (vd.data - vb.data).abs().max() < 1e-10
    save("rime_batched", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times,
         ra=ra, dec=dec, zen_az=zen_az, sky_params=sparams, px_area=px_area, beam_params=bp,
         theta_grid=theta, phi_grid=phi, vis=vd.data, fov=120.0)


def gold_rime_2pol():
    """2-pol power-beam mode (beam_model.py:351-357): Npol=2, two Airy diameters per feed."""
    rng = np.random.default_rng(13)
    freqs = torch.linspace(140e6, 160e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs)
    bls = [(0, 1), (0, 4), (2, 5), (1, 6)]
    Ns = 50
    ra = rng.uniform(0, 360, Ns)
    dec = rng.uniform(-70, 10, Ns)
    angs = torch.as_tensor(np.stack([ra, dec]))
    sparams = torch.as_tensor(np.abs(rng.normal(size=(1, 1, len(freqs), Ns))))
    sky = ba.sky_model.PointSky(sparams.clone(), angs,
                                R=ba.sky_model.PointSkyResponse(freqs, freq_mode='channel'),
                                parameter=True)
    bp = torch.tensor([[14.0, 12.5], [13.0, 15.0]]).reshape(2, 1, 1, 1, 2)
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=ba.beam_model.AiryResponse(powerbeam=True),
                                   powerbeam=True, fov=180, parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 102)
    backward_with(vd.data, G)
    save("rime_2pol", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra, dec=dec,
         zen_az=zen_az, sky_params=sparams, beam_params=bp, vis=vd.data, G=G,
         grad_sky=sky.params.grad, fov=180.0)


def gold_rime_4pol():
    """4-pol mode (beam_model.py:359-363): real 2x2 Jones PixelResponse x real coherency (I,Q,U)."""
    rng = np.random.default_rng(14)
    freqs = torch.linspace(140e6, 160e6, 5)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs, set_param=True)
    bls = [(0, 1), (0, 4), (2, 5), (1, 6), (3, 3)]
    ra, dec, px_area, _ = healpix_sky(4, freqs, rng)
    Npix = len(ra)
    sparams = np.zeros((3, 1, len(freqs), Npix))
    sparams[0] = np.abs(rng.normal(size=(1, len(freqs), Npix)))
    sparams[1] = 0.1 * rng.normal(size=(1, len(freqs), Npix))
    sparams[2] = 0.1 * rng.normal(size=(1, len(freqs), Npix))
    sparams = torch.as_tensor(sparams)
    angs = torch.as_tensor(np.stack([ra, dec]))
    skymod = ba.sky_model.PixelSky(sparams.clone(), angs, 1.0,
                                   R=ba.sky_model.PixelSkyResponse(freqs), parameter=True)

    class SkyChain(ba.utils.Module):
        """PixelSky followed by a Stokes2Coherency block (the documented usage)."""
        def __init__(self, sky):
            super().__init__(name=sky.name)
            self.sky = sky
            self.s2c = ba.sky_model.Stokes2Coherency()
            self.device = sky.device

        def forward(self, prior_cache=None):
            return self.s2c(self.sky(prior_cache=prior_cache))

    sky = SkyChain(skymod)
    theta, phi, b_theta, b_phi, airy = rect_airy_beam(freqs, 5.0, 10.0)
    a = torch.sqrt(airy)
    J = torch.zeros(2, 2, 1, len(freqs), a.shape[-1])
    J[0, 0, 0] = a
    J[1, 1, 0] = a
    J[0, 1, 0] = 0.05 * a * torch.sin(b_phi * ba.D2R)
    J[1, 0, 0] = 0.05 * a * torch.cos(b_phi * ba.D2R)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta=b_theta, phi=b_phi,
                                    theta_grid=theta, phi_grid=phi, freq_mode='channel',
                                    powerbeam=False, realbeam=True, log=False)
    beam = ba.beam_model.PixelBeam(J.clone(), freqs, R=R, powerbeam=False, fov=180, parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 103)
    backward_with(vd.data, G)
    save("rime_4pol", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra, dec=dec,
         zen_az=zen_az, sky_params=sparams, beam_params=J, theta_grid=theta, phi_grid=phi,
         vis=vd.data, G=G, grad_sky=skymod.params.grad, grad_beam=beam.params.grad,
         grad_antvecs=array.antvecs.grad, fov=180.0)


def gold_rime_multimodel():
    """1-pol voltage beams with two antenna beam models (Nmodel=2, ant2beam set after
    construction -- SURVEY section 9 item 4) and data_bls inflation of redundant baselines."""
    rng = np.random.default_rng(15)
    freqs = torch.linspace(140e6, 160e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs)
    bls = [(0, 1), (0, 4), (2, 5), (1, 6), (1, 3)]
    Ns = 40
    ra = rng.uniform(0, 360, Ns)
    dec = rng.uniform(-70, 10, Ns)
    angs = torch.as_tensor(np.stack([ra, dec]))
    sparams = torch.as_tensor(np.abs(rng.normal(size=(1, 1, len(freqs), Ns))))
    sky = ba.sky_model.PointSky(sparams.clone(), angs,
                                R=ba.sky_model.PointSkyResponse(freqs, freq_mode='channel'),
                                parameter=True)
    bp = torch.tensor([[14.0], [12.0]]).reshape(1, 1, 2, 1, 1)
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=ba.beam_model.AiryResponse(powerbeam=False),
                                   pol='e', powerbeam=False, fov=180, parameter=False,
                                   ant2beam={a: a % 2 for a in ants})
    beam.ant2beam = {a: a % 2 for a in ants}
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, bls, times, freqs)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 104)
    backward_with(vd.data, G)
    save("rime_multimodel", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra,
         dec=dec, zen_az=zen_az, sky_params=sparams, beam_params=bp,
         ant2beam=np.asarray([a % 2 for a in ants]), vis=vd.data, G=G,
         grad_sky=sky.params.grad, fov=180.0)


def gold_rime_databls():
    """Redundant-baseline inflation through data_bls (rime_model.py:201-224, 436-437)."""
    rng = np.random.default_rng(16)
    freqs = torch.linspace(140e6, 160e6, 4)
    times = np.linspace(2458148.15, 2458148.25, 2)
    ants, vecs, array = hera_array(2, freqs)
    sim_bls = array.get_bls(uniq_bls=True, keep_autos=False)
    data_bls = []
    for red in array.reds:
        for bl in red:
            if bl[0] != bl[1]:
                data_bls.append(bl)
    Ns = 30
    ra = rng.uniform(0, 360, Ns)
    dec = rng.uniform(-70, 10, Ns)
    angs = torch.as_tensor(np.stack([ra, dec]))
    sparams = torch.as_tensor(np.abs(rng.normal(size=(1, 1, len(freqs), Ns))))
    sky = ba.sky_model.PointSky(sparams.clone(), angs,
                                R=ba.sky_model.PointSkyResponse(freqs, freq_mode='channel'),
                                parameter=True)
    bp = torch.ones(1, 1, 1, 1, 1) * 14.0
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=ba.beam_model.AiryResponse(powerbeam=True),
                                   pol='e', powerbeam=True, fov=180, parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC)
    rime = ba.rime_model.RIME(sky, tel, beam, array, sim_bls, times, freqs, data_bls=data_bls)
    zen_az = inject_geometry(rime, sky.name, ra, dec, rime.sim_times)
    vd = rime()
    G = cotangent(vd.data.shape, 105)
    backward_with(vd.data, G)
    save("rime_databls", antvecs=vecs, ants=ants, sim_bls=sim_bls, data_bls=rime.data_bls,
         sim2data=rime._sim2data[0], freqs=freqs, times=times, ra=ra, dec=dec, zen_az=zen_az,
         sky_params=sparams, beam_params=bp, vis=vd.data, G=G, grad_sky=sky.params.grad, fov=180.0)


def gold_vismapper():
    """imaging.VisMapper: dirty maps, PSF diagonal and normalisation for the
    three normalisation methods, weighted and unweighted, and compute_Am, hex-7 x 6 freqs x 3
    times x 80 pixels, Airy power beam with a 150 deg field of view."""
    rng = np.random.default_rng(77)
    freqs = torch.linspace(120e6, 180e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 3)
    ants, vecs, array = hera_array(2, freqs)
    bls = [(ants[i], ants[j]) for i in range(len(ants)) for j in range(i + 1, len(ants))]
    Npix = 80
    ra = rng.uniform(0, 360, Npix)
    dec = np.degrees(np.arcsin(rng.uniform(-1, np.sin(np.radians(29)), Npix)))
    tel = ba.telescope_model.TelescopeModel(LOC)
    data = cotangent((1, 1, len(bls), len(times), len(freqs)), 5)
    icov = torch.as_tensor(rng.uniform(0.5, 2.0, data.shape))
    vd = ba.dataset.VisData()
    vd.setup_meta(telescope=tel, antpos=ba.utils.AntposDict(ants, torch.as_tensor(vecs)))
    vd.setup_data(bls, times, freqs, pol='ee', data=data, icov=icov, cov_axis=None)
    bp = torch.ones(1, 1, 1, 1, 1) * 14.0
    beam = ba.beam_model.PixelBeam(bp.clone(), freqs, R=ba.beam_model.AiryResponse(powerbeam=True),
                                   pol='e', powerbeam=True, fov=150, parameter=False)
    zen_az = []
    for t in times:
        zen, az = orc.eq2top_synth(t, ra, dec, lat=LOC[1])
        zen_az.append(np.stack([zen, az]))

    def mapper(*args, **kwargs):
        # the mapper works on its own copy of the VisData metadata: inject (zen, az) there
        vm = ba.imaging.VisMapper(*args, **kwargs)
        for t, za in zip(vm.times, zen_az):
            vm.telescope.conv_cache[vm.telescope.hash(t, ra)] = torch.as_tensor(za)
        return vm

    out = {}
    for weighted in (True, False):
        vdw = vd if weighted else vd.copy()
        if not weighted:
            vdw.icov = None
        vm = mapper(vdw, ra, dec, beam=beam)
        for method in ('w', 'Aw', 'A2w'):
            vm.set_normalization(method)
            maps, P = vm.make_map(return_P=True, contract='diag')
            tag = "%s_%s" % (method, 'icov' if weighted else 'ones')
            out["maps_" + tag], out["P_" + tag], out["D_" + tag] = maps, P, vm.D
    vm_nobeam = mapper(vd, ra, dec, beam=None, fov=120)
    vm_nobeam.set_normalization('A2w')
    maps_nb, P_nb = vm_nobeam.make_map(return_P=True, contract='diag')
    test_maps = torch.as_tensor(rng.normal(size=(2, len(freqs), Npix)))
    vm = mapper(vd, ra, dec, beam=beam)
    Am = vm.compute_Am(test_maps)
    vm.set_normalization('A2w')
    vm.make_map(return_P=False)
    Pm = vm.compute_Pm(test_maps, D=vm.D)
    save("vismapper", antvecs=vecs, ants=ants, bls=bls, freqs=freqs, times=times, ra=ra, dec=dec,
         zen_az=np.asarray(zen_az), vis=data, icov=icov, beam_params=bp, fov=150.0,
         fov_nobeam=120.0, maps_nobeam=maps_nb, P_nobeam=P_nb, test_maps=test_maps, Am=Am, Pm=Pm,
         **out)


def gold_apply_cal():
    """calibration._apply_cal: 1pol, 2pol (cal_2pol) and 4pol products, the
    undo direction, broadcast gains, the covariance update and autograd gradients to vis and gains
    for a fixed cotangent."""
    rng = np.random.default_rng(21)
    ants = [0, 1, 2, 3, 4, 5, 6]
    bls = [(a, b) for a in ants for b in ants if a <= b][::2]          # crosses and autos
    nbl, nt, nf = len(bls), 3, 5
    c = lambda *shape: torch.as_tensor(rng.normal(size=shape) + 1j * rng.normal(size=shape))
    out = dict(ants=ants, bls=bls)
    for tag, npol, cal_2pol, gshape_tf in (("1pol", 1, False, (nt, nf)), ("2pol", 2, True, (nt, nf)),
                                           ("4pol", 2, False, (nt, nf)), ("1pol_bcast", 1, False, (1, nf)),
                                           ("4pol_bcast", 2, False, (nt, 1))):
        vis = c(npol, npol, nbl, nt, nf).requires_grad_(True)
        gains = (c(npol, npol, len(ants), *gshape_tf) * 0.3 + torch.eye(npol)[:, :, None, None, None]
                 ).requires_grad_(True)
        cov = torch.as_tensor(rng.uniform(0.5, 2, size=(npol, npol, nbl, nt, nf))) \
            if tag in ("1pol", "2pol") else None
        vout, cov_out = ba.calibration.apply_cal(vis, bls, gains, ants, cal_2pol=cal_2pol, cov=cov)
        G = cotangent(vout.shape, 300 + npol)
        backward_with(vout, G)
        out.update({tag + "_vis": vis, tag + "_gains": gains, tag + "_out": vout, tag + "_G": G,
                    tag + "_dvis": vis.grad, tag + "_dgains": gains.grad})
        if not tag.startswith("4pol"):
            # the reference's 4pol undo calls torch.pinv, which torch 2.11 does not have
            # (calibration.py:2444), so only the diagonal modes have a reference answer
            with torch.no_grad():
                vundo, _ = ba.calibration.apply_cal(vis, bls, gains, ants, cal_2pol=cal_2pol,
                                                    undo=True)
            out[tag + "_undo"] = vundo
        if cov is not None:
            out.update({tag + "_cov": cov, tag + "_cov_out": cov_out})
    save("apply_cal", **out)


def gold_jones_model():
    """calibration.JonesModel.forward (calibration.py:599-664): model
    visibilities through the Jones term for 1pol complex gains with a reference antenna, a
    time-minibatched VisData (gain times indexed with atol), 2pol amplitude+phase gains, delay
    gains, 4pol complex gains; outputs and autograd gradients to the gain parameters and the
    model visibilities for a fixed cotangent."""
    rng = np.random.default_rng(33)
    ants = [0, 1, 2, 3, 4, 5]
    bls = [(a, b) for a in ants for b in ants if a < b]
    freqs = torch.linspace(120e6, 180e6, 6)
    times = np.linspace(2458148.15, 2458148.25, 4)
    nbl, nt, nf, na = len(bls), len(times), len(freqs), len(ants)
    tel = ba.telescope_model.TelescopeModel(LOC)
    c = lambda *shape: torch.as_tensor(rng.normal(size=shape) + 1j * rng.normal(size=shape))

    def visdata(npol, tsel):
        vd = ba.dataset.VisData()
        vd.setup_meta(telescope=tel)
        data = c(npol, npol, nbl, len(tsel), nf).requires_grad_(True)
        vd.setup_data(bls, times[tsel], freqs, pol='ee' if npol == 1 else None, data=data)
        return vd

    out = dict(ants=ants, bls=bls, freqs=freqs, times=times)
    cases = (("1pol_com_refant", 1, '1pol', 'com', 2, [0, 1, 2, 3]),
             ("1pol_com_tbatch", 1, '1pol', 'com', None, [1, 3]),
             ("2pol_ampphs", 2, '2pol', 'amp_phs', 0, [0, 1, 2, 3]),
             ("1pol_dly", 1, '1pol', 'dly', 1, [0, 1, 2, 3]),
             ("4pol_com", 2, '4pol', 'com', None, [0, 2, 3]))
    for tag, npol, polmode, ptype, refant, tsel in cases:
        if ptype == 'com':
            params = c(npol, npol, na, nt, nf) * 0.3 + torch.eye(npol)[:, :, None, None, None]
        elif ptype == 'amp_phs':
            params = torch.as_tensor(rng.normal(size=(npol, npol, na, nt, nf, 2))) * 0.2
        else:
            params = torch.as_tensor(rng.normal(size=(npol, npol, na, nt, nf))) * 3.0   # ns
        R = ba.calibration.JonesResponse(param_type=ptype, freqs=freqs, times=torch.as_tensor(times))
        J = ba.calibration.JonesModel(params.clone(), ants, refant=refant, R=R, polmode=polmode)
        vd = visdata(npol, tsel)
        vout = J(vd)
        G = cotangent(vout.data.shape, 500 + len(tag))
        backward_with(vout.data, G)
        out.update({tag + "_params_in": params, tag + "_params": J.params.detach().clone(),
                    tag + "_vis": vd.data, tag + "_tsel": np.asarray(tsel), tag + "_out": vout.data,
                    tag + "_G": G, tag + "_dparams": J.params.grad, tag + "_dvis": vd.data.grad})
        out[tag + "_refant"] = -1 if refant is None else refant
    save("jones_model", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["gold_" + name]()
        sys.exit(0)
    gold_fringe()
    gold_airy()
    gold_rect_interp()
    gold_rime_point_airy()
    gold_rime_airy_brute()
    gold_rime_pixel_interp()
    gold_alm()
    gold_rime_ylm()
    gold_rime_alm_sky()
    gold_rime_pointing()
    gold_rime_batched()
    gold_rime_2pol()
    gold_rime_4pol()
    gold_rime_multimodel()
    gold_rime_databls()
    gold_vismapper()
    gold_apply_cal()
    gold_jones_model()
