"""
Synthetic workloads of BASELINE.json, built with bayeslim_b200's own
classes exactly as a BayesLIM user would build them with the reference package.  Shared by
bench.py, __graft_entry__.smoke() and the GPU parity tests.  No dataset is read: antenna
layouts, skies and beams are generated from fixed seeds.

  C1  HERA-37, 1k point sources (power law), Airy beam, 64 freqs, 10 times
  C2  HERA-37, 10k point sources, 256 freqs, 60 times, grads to sky
  C3  HERA-350, HEALPix nside-128 PixelSky, rect-grid interpolated PixelBeam, 1024 freqs,
      grads to sky, beam (and optionally antenna positions)
  C4  HERA-350, 4-pol Jones beams, nside-256 PixelSky with Stokes I, Q, U, 1024 freqs
      (pixel_interp_pol), time-sharded
  C5  HERA-350, nside-256 PixelSky, 1024 freqs, full-night time axis in single-time minibatches x
      two block-aligned baseline groups (pixel_interp(bl_groups=True, time_groups=True))
"""
import itertools
import math

import numpy as np
import torch

import bayeslim_b200 as ba

LOCATION = (21.42827, -30.72148, 1051.7)     # reference tests/test_telescope.py:13


def hera37():
    return ba.utils._make_hex(4, D=14.6)


def hera350():
    """Hex core (N=11, 331 antennas) + 19 outriggers at radius 320+12k m, angle 2 pi k/19."""
    ants, vecs = ba.utils._make_hex(11, D=14.6)
    k = np.arange(19)
    r = 320.0 + 12.0 * k
    ang = 2 * np.pi * k / 19
    out = np.stack([r * np.cos(ang), r * np.sin(ang), np.zeros(19)], axis=1)
    vecs = np.concatenate([vecs, out], axis=0)
    return list(range(len(vecs))), vecs


def all_cross_bls(ants):
    return list(itertools.combinations(ants, 2))


def make_array(ants, vecs, freqs, device, antpos_param=False):
    antpos = ba.utils.AntposDict(ants, torch.as_tensor(vecs, dtype=torch.float64, device=device))
    array = ba.telescope_model.ArrayModel(antpos, freqs=freqs, device=device, skip_reds=True)
    if antpos_param:
        array.set_param('antvecs')
    return array


def point_airy(n_src, n_freq, n_time, device, dtype=torch.float32, bls='all', seed=0,
               sky_param=True, beam_param=False, antpos_param=False, layout='hera37'):
    """C1 / C2 family."""
    rng = np.random.default_rng(seed)
    freqs = torch.linspace(100e6, 200e6, n_freq, dtype=torch.float64, device=device)
    ants, vecs = hera37() if layout == 'hera37' else hera350()
    array = make_array(ants, vecs, freqs, device, antpos_param)
    if bls == 'all':
        sim_bls = all_cross_bls(ants)
    else:
        full = ba.telescope_model.ArrayModel(dict(zip(ants, vecs)), freqs=freqs)
        sim_bls = full.get_bls(uniq_bls=True, keep_autos=False)
    ra = rng.uniform(0, 360, n_src)
    dec = np.degrees(np.arcsin(rng.uniform(-1, math.sin(math.radians(29)), n_src)))
    params = np.zeros((1, 1, 2, n_src))
    params[0, 0, 0] = np.exp(rng.normal(size=n_src))
    params[0, 0, 1] = rng.normal(-0.8, 0.2, n_src)
    R = ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='powerlaw', f0=150e6, device=device)
    sky = ba.sky_model.PointSky(torch.as_tensor(params, dtype=dtype, device=device),
                                torch.as_tensor(np.stack([ra, dec]), device=device), R=R,
                                parameter=sky_param)
    beam = ba.beam_model.PixelBeam(torch.ones(1, 1, 1, 1, 1, dtype=dtype, device=device) * 14.0, freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                   powerbeam=True, fov=180, parameter=beam_param)
    tel = ba.telescope_model.TelescopeModel(LOCATION, device=device)
    times = np.linspace(2458148.15, 2458148.25, n_time)
    rime = ba.RIME(sky, tel, beam, array, sim_bls, times, freqs, device=device)
    return rime


def healpix_sky_angles(nside, dec_max=59.27852):
    theta, phi = ba.healpix.pix2ang(nside)
    dec = np.pi / 2 - theta
    keep = dec < math.radians(dec_max)
    return np.degrees(phi[keep]), np.degrees(dec[keep])


def rect_airy_map(freqs, dtheta=1.0, dphi=1.0, D=14.0, dtype=torch.float32, device='cpu'):
    """Airy power beam sampled on a (phi, theta) grid -- construction of reference
    tests/test_beam.py:13-32."""
    theta = torch.arange(0, 90.0 + 1e-6, dtheta, dtype=torch.float64, device=device)
    phi = torch.arange(0, 360.0 - 1e-6, dphi, dtype=torch.float64, device=device)
    b_phi, b_theta = torch.meshgrid(phi, theta, indexing='xy')
    airy = ba.beam_model.airy_disk(b_theta.ravel() * ba.D2R, b_phi.ravel() * ba.D2R, D,
                                   freqs.double(), square=True)
    return theta, phi, airy.to(dtype)


def block_aligned_groups(ants, bls, block=128):
    """Baseline groups aligned to the 128-antenna row blocks of the tensor-core items: group g
    holds the pairs whose lower antenna row falls into... block 0 (group 0) or any later block
    (group 1).  Each group then maps onto whole items (no antenna term is generated for a pair
    the group does not own) and onto disjoint blocks of the cotangent matrix in the backward."""
    row = {a: k for k, a in enumerate(ants)}
    g0 = [b for b in bls if min(row[b[0]], row[b[1]]) < block]
    g1 = [b for b in bls if min(row[b[0]], row[b[1]]) >= block]
    return [g for g in (g0, g1) if g]


def pixel_interp(nside, n_freq, n_time, device, dtype=torch.float32, n_bl=None, seed=0,
                 sky_param=True, beam_param=True, antpos_param=False, layout='hera350',
                 dgrid=1.0, bl_groups=False, time_groups=False):
    """C3 family (bl_groups / time_groups: the C5 minibatch grid of block-aligned baseline
    groups x single times)."""
    gen = torch.Generator(device='cpu').manual_seed(seed)
    freqs = torch.linspace(100e6, 200e6, n_freq, dtype=torch.float64, device=device)
    ants, vecs = hera350() if layout == 'hera350' else hera37()
    array = make_array(ants, vecs, freqs, device, antpos_param)
    sim_bls = all_cross_bls(ants)
    if n_bl is not None:
        step = max(1, len(sim_bls) // n_bl)
        sim_bls = sim_bls[::step][:n_bl]
    ra, dec = healpix_sky_angles(nside)
    npix = len(ra)
    spec = (freqs / 150e6) ** -2.5
    base = torch.randn(npix, generator=gen).abs()
    params = (spec[:, None].cpu() * base[None, :]).to(dtype)[None, None].to(device)
    sky = ba.sky_model.PixelSky(params, torch.as_tensor(np.stack([ra, dec]), device=device),
                                ba.healpix.nside2pixarea(nside),
                                R=ba.sky_model.PixelSkyResponse(freqs.to(dtype), device=device),
                                parameter=sky_param)
    theta, phi, airy = rect_airy_map(freqs, dgrid, dgrid, 14.0, dtype, device)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta_grid=theta,
                                    phi_grid=phi, freq_mode='channel', powerbeam=True,
                                    realbeam=True, log=False, device=device)
    beam = ba.beam_model.PixelBeam(airy[None, None, None].contiguous(), freqs, R=R, pol='e',
                                   powerbeam=True, fov=180, parameter=beam_param)
    tel = ba.telescope_model.TelescopeModel(LOCATION, device=device)
    times = np.linspace(2458148.15, 2458148.25, n_time)
    if bl_groups:
        sim_bls = block_aligned_groups(ants, sim_bls)
    if time_groups:
        times = [np.asarray([t]) for t in times]
    rime = ba.RIME(sky, tel, beam, array, sim_bls, times, freqs, device=device)
    return rime


class _CoherencySky(ba.utils.Module):
    """PixelSky (Stokes I and polarisation fractions) followed by Stokes2Coherency: the
    documented way to feed a polarised RIME (sky_model.py:1160-1353)."""

    def __init__(self, sky):
        super().__init__(name=sky.name)
        self.sky = sky
        self.s2c = ba.sky_model.Stokes2Coherency()
        self.device = sky.device
        self.angs = sky.angs

    def forward(self, prior_cache=None):
        return self.s2c(self.sky(prior_cache=prior_cache))


def pixel_interp_pol(nside, n_freq, n_time, device, dtype=torch.float32, n_bl=None, seed=0,
                     antpos_param=False, layout='hera350', dgrid=2.0):
    """C4 family: 4-pol real Jones beams (2 feeds x 2 sky vectors) interpolated from rect-grid
    maps, PixelSky with Stokes I, Q, U fractions -> real coherency."""
    gen = torch.Generator(device='cpu').manual_seed(seed)
    freqs = torch.linspace(100e6, 200e6, n_freq, dtype=torch.float64, device=device)
    ants, vecs = hera350() if layout == 'hera350' else hera37()
    array = make_array(ants, vecs, freqs, device, antpos_param)
    sim_bls = all_cross_bls(ants)
    if n_bl is not None:
        step = max(1, len(sim_bls) // n_bl)
        sim_bls = sim_bls[::step][:n_bl]
    ra, dec = healpix_sky_angles(nside)
    npix = len(ra)
    spec = (freqs / 150e6) ** -2.5
    base = torch.randn(npix, generator=gen).abs()
    params = torch.zeros(3, 1, n_freq, npix, dtype=dtype)
    params[0, 0] = (spec[:, None].cpu() * base[None, :]).to(dtype)
    params[1, 0] = 0.1 * torch.randn(npix, generator=gen).to(dtype)[None]      # Q / I
    params[2, 0] = 0.1 * torch.randn(npix, generator=gen).to(dtype)[None]      # U / I
    pix = ba.sky_model.PixelSky(params.to(device), torch.as_tensor(np.stack([ra, dec]), device=device),
                                ba.healpix.nside2pixarea(nside),
                                R=ba.sky_model.PixelSkyResponse(freqs.to(dtype), device=device),
                                parameter=True)
    sky = _CoherencySky(pix)
    theta, phi, airy = rect_airy_map(freqs, dgrid, dgrid, 14.0, dtype, device)
    # voltage-like Jones maps: co-polar sqrt(Airy), cross-polar leakage with an azimuthal pattern
    co = airy.clamp(min=0).sqrt()
    b_phi, _ = torch.meshgrid(phi, theta, indexing='xy')
    leak = 0.05 * torch.sin(2 * b_phi.ravel() * ba.D2R).to(dtype)[None] * co
    jones = torch.stack([torch.stack([co, leak]), torch.stack([-leak, co])])[:, :, None]
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta_grid=theta,
                                    phi_grid=phi, freq_mode='channel', powerbeam=False,
                                    realbeam=True, log=False, device=device)
    beam = ba.beam_model.PixelBeam(jones.contiguous(), freqs, R=R, powerbeam=False, fov=180,
                                   parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOCATION, device=device)
    times = np.linspace(2458148.15, 2458148.25, n_time)
    return ba.RIME(sky, tel, beam, array, sim_bls, times, freqs, device=device)


def count_evals(rime):
    """source x baseline x freq x time evaluations of the CURRENT batch (time group x baseline
    group), sources counted after the FOV cut.  Requires one forward of the batch
    to have populated the geometry cache."""
    times = tuple(float(t) for t in rime.sim_times)
    total = 0
    for key, rec in rime._geom_cache.items():
        if key[2] == times:
            total += sum(rec.geom.ns)
    return total * len(rime.sim_bls) * len(rime.array.freqs)


def zenaz_of(rime, sky_name=None):
    """Per-time (zen, az) [deg] tensors of the current time group, from the telescope cache."""
    out = []
    name = sky_name if sky_name is not None else rime.sky.name
    npix = rime.sky.angs.shape[1] if isinstance(rime.sky.angs, torch.Tensor) else len(rime.sky.angs[0])
    for t in rime.sim_times:
        out.append(rime.telescope.conv_cache[(name, npix, t)])
    return out
