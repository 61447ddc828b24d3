"""
Rebuild the golden fixtures (tests/golden/*.npz) as ``bayeslim_b200`` models -- the same
constructor calls a BayesLIM user would write against the reference package -- on a chosen
device and precision.  Shared by the host-logic tests (CPU, emulated kernels) and the GPU
parity tests.
"""
import numpy as np
import torch

import bayeslim_b200 as ba
from tests.oracle_cases import load, bl_list

LOC = (21.42827, -30.72148, 1051.7)


def _t(x, dtype, device):
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


def _inject(rime, name, npix, times, zen_az, device):
    for t, za in zip(times, zen_az):
        rime.telescope.conv_cache[(name, npix, t)] = torch.as_tensor(za, dtype=torch.float64,
                                                                     device=device)


def _array(g, freqs, device, dtype, param=False):
    ants = [int(a) for a in g["ants"]]
    antvecs = _t(g["antvecs"], torch.float64, device)
    array = ba.telescope_model.ArrayModel(ba.utils.AntposDict(ants, antvecs), freqs=freqs,
                                          device=device)
    if param:
        array.set_param('antvecs')
    return ants, array


def build_airy_brute(g, device, dtype=torch.float64):
    """AiryResponse(brute_force=True): no fused builder, the generic torch-response route."""
    return build_point_airy(g, device, dtype, params=("sky", "beam"),
                            response=ba.beam_model.AiryResponse(powerbeam=True, brute_force=True,
                                                                Ntau=int(g["Ntau"])))


def build_point_airy(g, device, dtype=torch.float64, params=("sky", "beam", "antvecs"),
                     response=None):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype, param="antvecs" in params)
    R = ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='powerlaw', f0=float(g["f0"]),
                                      device=device)
    sky = ba.sky_model.PointSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device), R=R,
                                parameter="sky" in params)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs,
                                   R=response if response is not None
                                   else ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                   powerbeam=True, fov=float(g["fov"]), parameter="beam" in params)
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    times = g["times"]
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["bls"]), times, freqs, device=device)
    _inject(rime, sky.name, len(g["ra"]), rime.sim_times, g["zen_az"], device)
    return rime, dict(sky=sky.params, beam=beam.params, antvecs=array.antvecs)


def build_pixel_interp(g, device, dtype=torch.float64, params=("sky", "beam", "antvecs"),
                       interp_mode='linear', bls=None, times=None):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype, param="antvecs" in params)
    sky = ba.sky_model.PixelSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device),
                                float(g["px_area"]),
                                R=ba.sky_model.PixelSkyResponse(freqs.to(dtype), device=device),
                                parameter="sky" in params)
    theta, phi = _t(g["theta_grid"], torch.float64, device), _t(g["phi_grid"], torch.float64, device)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode=interp_mode, theta_grid=theta,
                                    phi_grid=phi, freq_mode='channel', powerbeam=True,
                                    realbeam=True, log=False, device=device)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs, R=R, pol='e',
                                   powerbeam=True, fov=float(g["fov"]), parameter="beam" in params)
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    sim_bls = bl_list(g["bls"]) if bls is None else bls
    rime = ba.RIME(sky, tel, beam, array, sim_bls, g["times"] if times is None else times, freqs,
                   device=device)
    _inject(rime, sky.name, len(g["ra"]), g["times"], g["zen_az"], device)
    return rime, dict(sky=sky.params, beam=beam.params, antvecs=array.antvecs)


def build_pointing_interp(g, device, dtype=torch.float64):
    rime, leaves = build_pixel_interp(g, device, dtype)
    rime.beam.set_pointing_offset(*[float(x) for x in g["offset"]])
    return rime, leaves


def build_pointing_airy(g, device, dtype=torch.float64):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype)
    sky = ba.sky_model.PixelSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device),
                                float(g["px_area"]),
                                R=ba.sky_model.PixelSkyResponse(freqs.to(dtype), device=device),
                                parameter=True)
    beam = ba.beam_model.PixelBeam(torch.ones(1, 1, 1, 1, 1, dtype=dtype, device=device) * 14.0,
                                   freqs, R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                   powerbeam=True, fov=float(g["fov"]), parameter=False,
                                   offset=tuple(float(x) for x in g["offset"]))
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["bls"]), g["times"], freqs, device=device)
    _inject(rime, sky.name, len(g["ra"]), g["times"], g["zen_az"], device)
    return rime, dict(sky=sky.params)


def build_2pol(g, device, dtype=torch.float64):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype)
    sky = ba.sky_model.PointSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device),
                                R=ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='channel',
                                                                device=device), parameter=True)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=True), powerbeam=True,
                                   fov=float(g["fov"]), parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["bls"]), g["times"], freqs, device=device)
    _inject(rime, sky.name, len(g["ra"]), rime.sim_times, g["zen_az"], device)
    return rime, dict(sky=sky.params)


class SkyChain(ba.utils.Module):
    """PixelSky followed by Stokes2Coherency (the documented way to feed a polarised RIME)."""

    def __init__(self, sky):
        super().__init__(name=sky.name)
        self.sky = sky
        self.s2c = ba.sky_model.Stokes2Coherency()
        self.device = sky.device

    def forward(self, prior_cache=None):
        return self.s2c(self.sky(prior_cache=prior_cache))


def build_4pol(g, device, dtype=torch.float64):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype, param=True)
    skymod = ba.sky_model.PixelSky(_t(g["sky_params"], dtype, device),
                                   _t(np.stack([g["ra"], g["dec"]]), torch.float64, device), 1.0,
                                   R=ba.sky_model.PixelSkyResponse(freqs.to(dtype), device=device),
                                   parameter=True)
    sky = SkyChain(skymod)
    theta, phi = _t(g["theta_grid"], torch.float64, device), _t(g["phi_grid"], torch.float64, device)
    R = ba.beam_model.PixelResponse(freqs, 'rect', interp_mode='linear', theta_grid=theta,
                                    phi_grid=phi, freq_mode='channel', powerbeam=False,
                                    realbeam=True, log=False, device=device)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs, R=R, powerbeam=False,
                                   fov=float(g["fov"]), parameter=True)
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["bls"]), g["times"], freqs, device=device)
    _inject(rime, sky.name, len(g["ra"]), rime.sim_times, g["zen_az"], device)
    return rime, dict(sky=skymod.params, beam=beam.params, antvecs=array.antvecs)


def build_multimodel(g, device, dtype=torch.float64):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype)
    sky = ba.sky_model.PointSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device),
                                R=ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='channel',
                                                                device=device), parameter=True)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=False), pol='e',
                                   powerbeam=False, fov=float(g["fov"]), parameter=False,
                                   ant2beam={a: int(m) for a, m in zip(ants, g["ant2beam"])})
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["bls"]), g["times"], freqs, device=device)
    _inject(rime, sky.name, len(g["ra"]), rime.sim_times, g["zen_az"], device)
    return rime, dict(sky=sky.params)


def build_databls(g, device, dtype=torch.float64):
    freqs = _t(g["freqs"], torch.float64, device)
    ants, array = _array(g, freqs, device, dtype)
    sky = ba.sky_model.PointSky(_t(g["sky_params"], dtype, device),
                                _t(np.stack([g["ra"], g["dec"]]), torch.float64, device),
                                R=ba.sky_model.PointSkyResponse(freqs.to(dtype), freq_mode='channel',
                                                                device=device), parameter=True)
    beam = ba.beam_model.PixelBeam(_t(g["beam_params"], dtype, device), freqs,
                                   R=ba.beam_model.AiryResponse(powerbeam=True), pol='e',
                                   powerbeam=True, fov=float(g["fov"]), parameter=False)
    tel = ba.telescope_model.TelescopeModel(LOC, device=device)
    rime = ba.RIME(sky, tel, beam, array, bl_list(g["sim_bls"]), g["times"], freqs,
                   data_bls=bl_list(g["data_bls"]), device=device)
    _inject(rime, sky.name, len(g["ra"]), rime.sim_times, g["zen_az"], device)
    return rime, dict(sky=sky.params)


CASES = {
    "rime_point_airy": (build_point_airy, dict(sky="grad_sky", beam="grad_beam_truncated",
                                               antvecs="grad_antvecs")),
    "rime_pixel_interp": (build_pixel_interp, dict(sky="grad_sky", beam="grad_beam",
                                                   antvecs="grad_antvecs")),
    "rime_2pol": (build_2pol, dict(sky="grad_sky")),
    "rime_airy_brute": (build_airy_brute, dict(sky="grad_sky", beam="grad_beam")),
    "rime_4pol": (build_4pol, dict(sky="grad_sky", beam="grad_beam", antvecs="grad_antvecs")),
    "rime_multimodel": (build_multimodel, dict(sky="grad_sky")),
    "rime_databls": (build_databls, dict(sky="grad_sky")),
}


def run_case(name, device, dtype=torch.float64):
    """Forward + backward (cotangent G of the fixture).  Returns (V, {param: grad}, fixture)."""
    g = load(name)
    build, grads = CASES[name]
    rime, leaves = build(g, device, dtype)
    vd = rime()
    V = vd.data
    G = torch.as_tensor(g["G"]).to(device=device, dtype=V.dtype)
    loss = torch.sum(G.real * V.real + G.imag * V.imag)
    loss.backward()
    out = {k: leaves[k].grad for k in grads}
    return vd, out, g, grads
