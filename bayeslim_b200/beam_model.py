"""
Antenna primary-beam models, mirroring the reference's beam_model
(bayeslim/beam_model.py): ``PixelBeam`` (:19-568) with its response functions
``PixelResponse`` (:570-845), ``GaussResponse`` (:848-899), ``AiryResponse`` (:902-988) and
``UniformResponse`` (:991-1016), plus ``airy_disk`` (:1418-1482) and ``cut_sky_fov``
(:1681-1698).

The classes keep the reference's constructor arguments, attributes and torch-level methods
(``gen_beam``, ``apply_beam``) so user code and non-RIME callers keep working.  The RIME hot
path does not go through ``apply_beam`` / ``PixInterp.interp`` / ``airy_disk``: it reads the
model state (params, p0, response type, interpolation tables) and evaluates the beam inside
the fused CUDA builders (ops.build_interp / ops.build_airy), see rime_model.RIME.
"""
import math

import numpy as np
import torch

from . import utils, sph_harm
from .utils import _float, D2R

C_LIGHT = 2.99792458e8


class PixelBeam(utils.Module):
    """Beam evaluated at the (zen, az) of every sky source: params -> R(params + p0, zen, az,
    freqs) of shape (Npol, Nvec, Nmodel, Nfreqs, Nsources)."""

    def __init__(self, params, freqs, R=None, ant2beam=None, parameter=True, pol=None,
                 powerbeam=True, fov=180, name=None, p0=None, offset=None, skycut_cache=False,
                 skycut_device=None):
        super().__init__(name=name)
        self.params = params
        self.p0 = p0
        self.device = self.params.device
        if parameter:
            self.params = torch.nn.Parameter(self.params)
        self.R = R if R is not None else UniformResponse()
        self.powerbeam = powerbeam
        if hasattr(self.R, 'powerbeam'):
            assert self.powerbeam == self.R.powerbeam
        self.Npol, self.Nvec, self.Nmodel = params.shape[0], params.shape[1], params.shape[2]
        if self.powerbeam:
            assert self.Nmodel == self.Nvec == 1
        self.freqs = freqs
        self.Nfreqs = len(freqs)
        self.fov = fov
        self.pol = pol
        if ant2beam is None:
            assert params.shape[2] == 1, "only 1 model for default ant2beam"
            self.ant2beam = utils.SimpleIndex()
        else:
            # NOTE: the reference drops a user-supplied ant2beam here (beam_model.py:153-155,
            # SURVEY section 9 item 4); we keep it, which is what its docstring promises.
            self.ant2beam = ant2beam
        offset = (0, 0) if offset is None else offset
        self.set_pointing_offset(*offset)
        self.skycut_cache = skycut_cache
        self.skycut_device = skycut_device
        self.clear_cache()
        self._args = dict(powerbeam=powerbeam, fov=fov, Npol=self.Npol, Nmodel=self.Nmodel)
        self._args[self.R.__class__.__name__] = getattr(self.R, '_args', None)

    def push(self, device):
        if not isinstance(device, torch.dtype):
            self.device = device
        self.params = utils.push(self.params, device)
        self.R.push(device)
        self.freqs = self.freqs.to(device)
        if self.p0 is not None:
            self.p0 = utils.push(self.p0, device)
        for priors in (self.priors_inp_params, self.priors_out_params):
            if priors is not None:
                for pr in priors:
                    if pr is not None:
                        pr.push(device)

    # ---- FOV cut -----------------------------------------------------------------
    def sky_cut(self, zen):
        """Indices of sources with zen < fov/2 (strict), or slice(None) for fov >= 360
        (beam_model.py:221-224); cached per arr_hash(zen) when skycut_cache is on."""
        cached = self.query_cache(zen) if self.skycut_cache else None
        if cached is not None:
            return cached
        cut = torch.where(zen < self.fov / 2)[0] if self.fov < 360 else slice(None)
        if self.skycut_cache:
            self.set_skycut_cache(zen, cut, device=self.skycut_device)
        return cut

    def total_params(self):
        return self.params if self.p0 is None else self.params + self.p0

    def gen_beam(self, zen, az, prior_cache=None):
        """(beam, cut, zen[cut], az[cut]) -- torch evaluation of the response, reference
        semantics (beam_model.py:197-271)."""
        zen_hash = getattr(zen, '_arr_hash', None)
        cut = self.sky_cut(zen)
        zen, az = zen[cut], az[cut]
        if zen_hash:
            zen._arr_hash = zen_hash
        p = self.total_params()
        new_zen, new_az = zen, az
        new_zen, new_az = offset_zen_az(self, zen, az)
        beam = self.R(p, new_zen, new_az, self.freqs)
        if getattr(self, '_hook_registry', None) is not None:
            bc = getattr(self.R, 'beam_cache', None)
            if bc is not None and bc.requires_grad:
                for hook in self._hook_registry:
                    bc.register_hook(hook)
        self.eval_prior(prior_cache)
        return beam, cut, zen, az

    def model_pairs(self, bls):
        """Sorted unique (model_i, model_j) pairs of a baseline list and, per baseline, the index
        of its pair (beam_model.py:303-305, 366-367)."""
        bls = utils.blnum2ants(bls)
        if isinstance(bls, tuple):
            bls = [bls]
        pairs = [(self.ant2beam[b[0]], self.ant2beam[b[1]]) for b in bls]
        uniq = sorted(set(pairs))
        lookup = {mp: i for i, mp in enumerate(uniq)}
        return uniq, [lookup[p] for p in pairs]

    def apply_beam(self, beam, bls, sky):
        """Perceived sky (Npol, Npol, Nbls, Nfreqs, Nsources) = beam_p . sky . beam_q^H, torch
        evaluation with reference semantics (beam_model.py:273-372)."""
        modelpairs, mp_idx = self.model_pairs(bls)
        if not utils.check_devices(beam.device, self.device):
            beam = beam.to(self.device)
        if not utils.check_devices(sky.device, self.device):
            sky = sky.to(self.device)
        psky = perceived_sky(beam, sky, modelpairs, self.Npol, self.Nvec, self.powerbeam)
        if len(modelpairs) > 1:
            return torch.index_select(psky, 2, torch.as_tensor(mp_idx, device=psky.device))
        return psky.expand(psky.shape[:2] + (len(mp_idx),) + psky.shape[3:])

    def forward(self, sky_comp, telescope, time, bls, prior_cache=None, **kwargs):
        zen, az = telescope.eq2top(time, sky_comp.angs[0], sky_comp.angs[1], store=False)
        beam, cut, zen, az = self.gen_beam(zen, az, prior_cache=prior_cache)
        sky = cut_sky_fov(sky_comp.data, cut)
        return dict(sky=self.apply_beam(beam, bls, sky), angs=cut_sky_fov(sky_comp.angs, cut),
                    zenaz=torch.vstack([zen, az]))

    def eval_prior(self, prior_cache, inp_params=None, out_params=None):
        """Priors on params and on the forwarded pixel beam (beam_model.py:421-465)."""
        if prior_cache is None or self.name in prior_cache:
            return
        total = torch.as_tensor(0.0)
        if self.priors_inp_params is not None:
            if inp_params is None:
                inp_params = self.params
            for prior in self.priors_inp_params:
                if prior is not None:
                    total = total + prior(inp_params)
        if self.priors_out_params is not None:
            if out_params is None and hasattr(self.R, 'beam_cache'):
                if self.R.beam_cache is None:
                    self.R.set_beam_cache(self.total_params())
                out_params = self.R.beam_cache
            for prior in self.priors_out_params:
                if prior is not None:
                    total = total + prior(out_params)
        prior_cache[self.name] = total

    def clear_graph_tensors(self):
        if hasattr(self.R, 'clear_beam_cache'):
            self.R.clear_beam_cache()

    def set_pointing_offset(self, theta_x=0, theta_y=0):
        self.theta_x = theta_x
        self.theta_y = theta_y

    def set_skycut_cache(self, zen, cut, device=None):
        h = utils.arr_hash(zen)
        if h not in self.cache:
            if isinstance(cut, torch.Tensor) and not utils.check_devices(cut.device, device):
                cut = cut.to(device)
            self.cache[h] = cut

    def query_cache(self, zen):
        return self.cache.get(utils.arr_hash(zen), None)

    def clear_cache(self):
        self.cache = {}


def perceived_sky(beam, sky, modelpairs, Npol, Nvec, powerbeam):
    """beam (Npol, Nvec, Nmodel, Nf, Ns), sky (Nvec, Nvec, Nf, Ns) -> (Npol|2, Npol|1, Nmp, Nf, Ns).
    The trailing (Nf, Ns) axes may be any broadcastable trailing shape (RIME passes tensors in
    the tiled layout (nchunk, S, KC)).
    The four polarisation modes of beam_model.py:334-363; mixed real/complex operands of the
    Jones product are promoted to a common dtype (the reference's einsum raises otherwise)."""
    i1 = torch.as_tensor([mp[0] for mp in modelpairs], device=beam.device)
    beam1 = beam.index_select(2, i1)
    if sky.ndim == beam.ndim - 1:          # give the sky a model-pair axis
        sky = sky[:, :, None]
    if powerbeam:
        if Npol == 1:
            assert sky.shape[:2] == (1, 1)
            return beam1 * sky
        assert Nvec == 1 and sky.shape[:2] == (1, 1)
        return torch.stack([beam1[0, 0] * sky[0, 0], beam1[1, 0] * sky[0, 0]])[:, None]
    i2 = torch.as_tensor([mp[1] for mp in modelpairs], device=beam.device)
    beam2 = beam.index_select(2, i2)
    if Npol == 1 and Nvec == 1:
        assert sky.shape[:2] == (1, 1)
        return (beam1 * beam2.conj()) * sky
    assert sky.shape[:2] == (2, 2)
    dt = torch.result_type(beam1, sky)
    return torch.einsum("ab...,bc...,dc...->ad...", beam1.to(dt), sky.to(dt), beam2.conj().to(dt))


class PixelResponse(utils.PixInterp):
    """Pixelised beam map interpolated at the source directions (beam_model.py:570-845).

    params (Npol, Nvec, Nmodel, Nfreqs, Npix) --forward()--> beam_cache, computed once per RIME
    forward and interpolated at every time."""

    def __init__(self, freqs, pixtype, beam0=None, comp_params=False, interp_mode='nearest',
                 theta=None, phi=None, theta_grid=None, phi_grid=None, freq_mode='channel',
                 freq_LM=None, nside=None, device=None, log=False, powerbeam=True, realbeam=True,
                 Rchi=None, interp_cache_depth=None, taper_kwargs=None, LM=None, norm_pix=None):
        super().__init__(pixtype, interp_mode=interp_mode, nside=nside, device=device,
                         theta_grid=theta_grid, phi_grid=phi_grid,
                         interp_cache_depth=interp_cache_depth)
        assert isinstance(comp_params, bool)
        if Rchi is not None:
            raise NotImplementedError("Rchi is not implemented (nor in the reference, :842)")
        self.beam0 = beam0
        self.theta, self.phi = theta, phi
        self.powerbeam = powerbeam
        self.realbeam = True if powerbeam else realbeam
        self.freqs = freqs
        self.comp_params = comp_params
        self.device = device
        self.log = log
        self.freq_mode = freq_mode
        self.freq_ax = 3
        self.Rchi = Rchi
        self.clear_beam_cache()
        self.taper_kwargs = taper_kwargs
        self.LM = LM
        self.norm_pix = norm_pix
        self.freq_LM = freq_LM
        self._args = dict(interp_mode=interp_mode, freq_mode=freq_mode)

    def _setup(self, **kwargs):
        pass

    def push(self, device):
        super().push(device)
        self.freqs = self.freqs.to(device)
        for name in ('theta', 'phi', 'beam0'):
            setattr(self, name, utils.push(getattr(self, name), device))
        for lm in (self.LM, self.freq_LM):
            if lm is not None:
                lm.push(device)

    def forward(self, params):
        """params -> pixel beam map (beam_model.py:750-793)."""
        if not utils.check_devices(params.device, self.device):
            params = params.to(self.device)
        if self.LM is not None:
            params = self.LM(params)
        if self.comp_params and not torch.is_complex(params):
            params = utils.viewcomp(params)
        p = params if self.freq_mode == 'channel' else self.freq_LM(params)
        if self.realbeam:
            p = p.real
        if self.log:
            p = torch.exp(p)
        elif self.powerbeam:
            p = torch.abs(p)
        if self.beam0 is not None:
            p = p + self.beam0
        if self.taper_kwargs is not None:
            p = p * beam_edge_taper(self.theta, device=p.device, **self.taper_kwargs)
        if self.norm_pix is not None:
            p = p / p[..., self.norm_pix:self.norm_pix + 1].detach().abs()
        return p

    def __call__(self, params, zen, az, *args):
        if self.beam_cache is None:
            self.set_beam_cache(params)
        return self.interp(self.beam_cache, zen, az)

    def clear_beam_cache(self):
        self.beam_cache = None

    def set_beam_cache(self, params):
        self.beam_cache = self.forward(params)
        return self.beam_cache


class YlmResponse(PixelResponse, sph_harm.AlmModel):
    """Spherical-harmonic beam: params (Npol, Nvec, Nmodel, Ndeg, Ncoeff) hold a_lm coefficients
    that are forward-modelled to the pixel map (Npol, Nvec, Nmodel, Nfreqs, Npix)
    (beam_model.py:1019-1267).  mode 'interpolate' (the RIME mode): the map at the fixed
    (theta, phi) is set once per forward as beam_cache -- by the CUDA tensor-core product
    (ops.alm_forward) -- and interpolated at the source directions by the fused builder like any
    PixelResponse; mode 'generate' evaluates the harmonics at every (zen, az) it is called with.
    The polynomial compression along l (lm_poly_setup, experimental in the reference) is not
    provided."""

    def __init__(self, l, m, freqs, pixtype='healpix', beam0=None, comp_params=False,
                 mode='interpolate', device=None, interp_mode='nearest', theta=None, phi=None,
                 theta_grid=None, phi_grid=None, nside=None, powerbeam=True, realbeam=True,
                 log=False, freq_mode='channel', freq_LM=None, Ylm_kwargs=None, Rchi=None,
                 separable=False, interp_cache_depth=None, taper_kwargs=None, LM=None,
                 norm_pix=None):
        realbeam = True if powerbeam else realbeam
        PixelResponse.__init__(self, freqs, pixtype, nside=nside, beam0=beam0,
                               interp_mode=interp_mode, theta=theta, phi=phi, freq_mode=freq_mode,
                               comp_params=comp_params, freq_LM=freq_LM, Rchi=Rchi,
                               theta_grid=theta_grid, phi_grid=phi_grid, norm_pix=norm_pix,
                               interp_cache_depth=interp_cache_depth, powerbeam=powerbeam,
                               realbeam=realbeam, device=device, log=log,
                               taper_kwargs=taper_kwargs)
        sph_harm.AlmModel.__init__(self, l, m, default_kw=Ylm_kwargs, real_output=realbeam, LM=LM)
        self.mode = mode
        self.separable = separable
        self.device = device
        self.lm_poly_setup()
        self._args = dict(mode=mode, interp_mode=interp_mode, freq_mode=freq_mode)

    def lm_poly_setup(self, lm_poly_kwargs=None):
        if lm_poly_kwargs not in (None, {}):
            raise NotImplementedError("polynomial compression along l is not provided")
        self._lm_poly = False

    def forward(self, params, zen, az, *args):
        """a_lm -> pixel beam at (zen, az) [deg] (beam_model.py:1166-1233)."""
        if not utils.check_devices(params.device, self.device):
            params = params.to(self.device)
        if self.LM is not None:
            params = self.LM(params)
        if self.comp_params and not torch.is_complex(params):
            params = utils.viewcomp(params)
        p = params if self.freq_mode == 'channel' else self.freq_LM(params)
        Ylm, alm_mult = self.get_Ylm(zen, az, h=utils.arr_hash(zen), separable=self.separable)
        beam = self.forward_alm(p, Ylm=Ylm, alm_mult=alm_mult, ignoreLM=True)
        if self.log:
            beam = torch.exp(beam)
        elif self.powerbeam:
            beam = torch.abs(beam)
        if self.beam0 is not None:
            beam = beam + self.beam0
        if self.taper_kwargs is not None:
            beam = beam * beam_edge_taper(zen, device=beam.device, **self.taper_kwargs)
        if self.norm_pix is not None:
            beam = beam / beam[..., self.norm_pix:self.norm_pix + 1].detach().abs()
        return beam

    def __call__(self, params, zen, az, *args):
        if self.mode == 'generate':
            return self.forward(params, zen, az)
        if self.beam_cache is None:
            self.set_beam_cache(params)
        return self.interp(self.beam_cache, zen, az)

    def set_beam_cache(self, params):
        if self.separable:
            self.beam_cache = self.forward(params, self.theta_grid, self.phi_grid)
        else:
            self.beam_cache = self.forward(params, self.theta, self.phi)
        return self.beam_cache

    def push(self, device):
        if self.beam_cache is not None:
            self.beam_cache = utils.push(self.beam_cache, device)
        LM, self.LM = self.LM, None           # pushed once, by AlmModel.push
        PixelResponse.push(self, device)
        self.LM = LM
        sph_harm.AlmModel.push(self, device)


class GaussResponse:
    """exp(-(l^2/sig_ew^2 + m^2/sig_ns^2)/2) beam (beam_model.py:848-899); params (..., 2)."""

    def __init__(self, powerbeam=True):
        self.freq_mode = 'channel'
        self.freq_ax = 3
        self.powerbeam = powerbeam

    def _setup(self):
        pass

    def __call__(self, params, zen, az, freqs):
        zen_rad = torch.as_tensor(zen).to(params.device) * D2R
        az_rad = torch.as_tensor(az).to(params.device) * D2R
        srad = torch.sin(zen_rad)
        srad = torch.where(zen_rad > math.pi / 2, torch.ones_like(srad), srad)
        l, m = srad * torch.sin(az_rad), srad * torch.cos(az_rad)
        beam = torch.exp(-0.5 * ((l / params[..., 0:1]) ** 2 + (m / params[..., 1:2]) ** 2))
        return beam if self.powerbeam else torch.sqrt(beam)

    def push(self, device):
        pass


class AiryResponse:
    """Airy-disk beam; params (Npol, Nvec, Nmodel, 1, 1|2) = aperture diameter(s) [m]
    (beam_model.py:902-988).

    full_grad: the reference's gradient w.r.t. the diameter is truncated because
    torch.special.bessel_j1 has no derivative formula (J1 is treated as a constant).
    full_grad=False (default) reproduces that for parity; True uses the analytic derivative."""

    def __init__(self, freq_ratio=1.0, powerbeam=True, brute_force=False, Ntau=100,
                 taper_kwargs=None, full_grad=False):
        self.freq_ratio = freq_ratio
        self.freq_mode = 'other'
        self.freq_ax = None
        self.powerbeam = powerbeam
        self.brute_force = brute_force
        self.Ntau = Ntau
        self.taper_kwargs = taper_kwargs
        self.full_grad = full_grad

    def _setup(self):
        pass

    def __call__(self, params, zen, az, freqs):
        Dew = params[..., 0:1]
        Dns = params[..., 1:2] if params.shape[-1] > 1 else None
        zen = torch.as_tensor(zen).to(params.device)
        az = torch.as_tensor(az).to(params.device)
        beam = airy_disk(zen * D2R, az * D2R, Dew, freqs, Dns, self.freq_ratio,
                         square=self.powerbeam, brute_force=self.brute_force, Ntau=self.Ntau)
        if self.taper_kwargs is not None:
            beam = beam * beam_edge_taper(zen, device=beam.device, **self.taper_kwargs)
        return beam

    def push(self, device):
        pass


class UniformResponse:
    """Unit response everywhere (beam_model.py:991-1016)."""

    def __init__(self, freqs=None, device=None, taper_kwargs=None):
        self.freqs = freqs
        self.taper_kwargs = taper_kwargs
        self.device = device

    def _setup(self):
        pass

    def __call__(self, params, zen, az, freqs):
        out = torch.ones(params.shape[:3] + (len(freqs), len(zen)), dtype=_float(),
                         device=self.device)
        if self.taper_kwargs is not None:
            out = out * beam_edge_taper(zen, device=out.device, **self.taper_kwargs)
        return out

    def push(self, device):
        if not isinstance(device, torch.dtype):
            self.device = device


def j1(x, Ntau=100, brute_force=False):
    """Bessel J1: torch.special.bessel_j1, or the trapezoid Bessel integral (differentiable)
    when brute_force (special.py:498-535)."""
    if not brute_force:
        return torch.special.bessel_j1(x)
    t = torch.linspace(0, math.pi, Ntau, device=x.device, dtype=x.dtype)
    dt = t[1] - t[0]
    t = t.reshape((-1,) + (1,) * x.ndim)
    integrand = torch.cos(t - x * torch.sin(t))
    w = torch.full_like(integrand, 2.0)
    w[0] = 1.0
    w[-1] = 1.0
    return torch.sum(w * integrand, dim=0) * dt / 2.0 / math.pi


def airy_disk(zen, az, Dew, freqs, Dns=None, freq_ratio=1.0, square=True, Ntau=100,
              brute_force=False):
    """(2 J1(x)/x)^(2|1), x = pi nu D(az) sin(min(zen, pi/2)) freq_ratio / c clipped at 1e-10;
    zen, az in radians (beam_model.py:1418-1482)."""
    zen = torch.clamp(torch.as_tensor(zen), max=math.pi / 2)
    az = torch.as_tensor(az)
    diameter = Dew if Dns is None else Dns + torch.abs(torch.sin(az)) ** 2 * (Dew - Dns)
    freqs = torch.as_tensor(freqs).to(zen.device)
    x = diameter * torch.sin(zen) * math.pi * freqs.reshape(-1, 1) * freq_ratio / C_LIGHT
    x = x.clip(1e-10)
    beam = 2.0 * j1(x, Ntau=Ntau, brute_force=brute_force) / x
    return beam ** 2 if square else beam


def pointing_offset(theta, phi, theta_x=0, theta_y=0):
    """Small-angle pointing offset of (zenith, azimuth) [rad], reference conventions
    (beam_model.py:1631-1678): the direction is put on the unit sphere as
    (sin t cos p, sin t sin p, cos t), rotated about x-hat by theta_x, then about y-hat by
    theta_y (each only when > 0, `rotation` beam_model.py:1514-1545), and converted back with
    the reference's quadrant rules (new_phi in [0, 2 pi))."""
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    r = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
    if theta_x > 0:
        c, s = np.cos(theta_x), np.sin(theta_x)
        r = np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]]) @ r
    if theta_y > 0:
        c, s = np.cos(theta_y), np.sin(theta_y)
        r = np.array([[c, 0, s], [0, 1.0, 0], [-s, 0, c]]) @ r
    new_theta = np.arccos(r[2])
    xzero, yzero = np.isclose(r[0], 0), np.isclose(r[1], 0)
    xneg, ypos = r[0] < 0, r[1] > 0
    new_phi = np.zeros_like(new_theta)
    new_phi[~xzero] = np.arctan(r[1][~xzero] / r[0][~xzero])
    new_phi[xneg & ypos] += np.pi
    new_phi[xneg & ~ypos] -= np.pi
    new_phi[xzero & yzero] = 0.0
    new_phi[xzero & ypos] = np.pi / 2
    new_phi[xzero & ~ypos] = -np.pi / 2
    return new_theta, new_phi % (2 * np.pi)


def offset_zen_az(beam, zen, az):
    """(zen, az) [deg] at which the response of `beam` is evaluated: the pointing-offset branch
    of PixelBeam.gen_beam (beam_model.py:244-256).  Works on reference beam objects as well."""
    theta_x, theta_y = getattr(beam, 'theta_x', 0), getattr(beam, 'theta_y', 0)
    if not (theta_x > 0 or theta_y > 0):
        return zen, az
    nz, na = pointing_offset(utils.tensor2numpy(zen) * D2R, utils.tensor2numpy(az) * D2R,
                             theta_x, theta_y)
    if isinstance(zen, torch.Tensor):
        nz = (torch.as_tensor(nz) / D2R).to(zen.device)
        na = (torch.as_tensor(na) / D2R).to(zen.device)
    else:
        nz, na = nz / D2R, na / D2R
    return nz, na


def cut_sky_fov(sky, cut):
    """sky[..., cut] for an index tensor or slice (beam_model.py:1681-1698)."""
    if isinstance(cut, slice):
        return sky[..., cut]
    if isinstance(cut, np.ndarray):
        cut = torch.as_tensor(cut)
    if not utils.check_devices(cut.device, sky.device):
        cut = cut.to(sky.device)
    return sky.index_select(-1, cut)


def _tukey(n, alpha):
    """Tukey (tapered cosine) window of n points, the definition of scipy.signal.windows.tukey."""
    if alpha <= 0:
        return np.ones(n)
    if alpha >= 1:
        return np.hanning(n)
    k = np.arange(n, dtype=np.float64)
    width = int(np.floor(alpha * (n - 1) / 2.0))
    w = np.ones(n)
    k1, k3 = k[:width + 1], k[n - width - 1:]
    w[:width + 1] = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * k1 / alpha / (n - 1))))
    w[n - width - 1:] = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * k3 / alpha / (n - 1))))
    return w


def beam_edge_taper(zen, mode='gauss', fov=180, device=None, mu=85, sigma=2.5, alpha=0.1):
    """Edge taper of a beam response (beam_model.py:1701-1735): 'gauss' rolls the beam off beyond
    zenith angle mu [deg]; 'tukey' is a 5000-point Tukey window over [-fov/2, fov/2], linearly
    interpolated at zen and zero outside."""
    zen = torch.as_tensor(zen)
    taper = torch.ones(len(zen), device=device, dtype=_float())
    if mode == 'tukey':
        th = np.linspace(-fov / 2, fov / 2, 5000, endpoint=True)
        vals = np.interp(utils.tensor2numpy(zen).astype(np.float64), th, _tukey(5000, alpha),
                         left=0.0, right=0.0)
        taper[:] = torch.as_tensor(vals, device=device)
        return taper
    if mode != 'gauss':
        return taper
    s = zen >= mu
    taper[s] = torch.exp(-0.5 * (zen[s].to(taper.dtype) - mu) ** 2 / sigma ** 2)
    return taper
