"""
Rebuild each golden RIME fixture with the CPU oracle (oracle/rime_oracle.py).

Used by tests/test_oracle.py to pin the oracle against the reference's outputs, and
by the GPU parity tests as the fp64 checker at sizes beyond the fixtures.
"""
import os

import numpy as np
import torch

from oracle import rime_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def tt(x, dtype=torch.float64, grad=False):
    t = torch.as_tensor(np.asarray(x), dtype=dtype).clone()
    if grad:
        t.requires_grad_(True)
    return t


def bl_list(arr):
    return [tuple(int(v) for v in b) for b in np.asarray(arr)]


def real_loss(V, G):
    G = torch.as_tensor(G)
    return torch.sum(G.real * V.real + G.imag * V.imag)


def oracle_point_airy(g, dtype=torch.float64, brute_force=False):
    antvecs = tt(g["antvecs"], dtype, grad=True)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype, grad=True)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    sky = orc.point_sky_response(sky_params, freqs, 'powerlaw', f0=float(g["f0"]))
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    kw = dict(brute_force=True, Ntau=int(g["Ntau"])) if brute_force else {}
    beam_fn = lambda z, a: orc.airy_response(beam_params, z, a, freqs, powerbeam=True, **kw)
    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]))
    return V, dict(sky=sky_params, beam=beam_params, antvecs=antvecs)


def oracle_pointing(g, which, dtype=torch.float64):
    """Pointing-offset fixture: 'interp' (rect-interpolated pixel beam) or 'airy'."""
    antvecs = tt(g["antvecs"], dtype, grad=True)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype, grad=True)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    sky = sky_params * float(g["px_area"])
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    if which == 'interp':
        beam_cache = orc.pixel_response_forward(beam_params, powerbeam=True)

        def beam_fn(z, a):
            inds, wgts = orc.rect_interp_weights(g["theta_grid"], g["phi_grid"], z, a, 'linear')
            return orc.interp_map(beam_cache, inds, wgts.to(dtype))
    else:
        D = torch.ones(1, 1, 1, 1, 1, dtype=dtype) * 14.0
        beam_fn = lambda z, a: orc.airy_response(D, z, a, freqs, powerbeam=True)
    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]),
                         offset=tuple(float(x) for x in g["offset"]))
    return V, dict(sky=sky_params, beam=beam_params, antvecs=antvecs)


def oracle_pixel_interp(g, dtype=torch.float64, interp_mode='linear', grad=True):
    antvecs = tt(g["antvecs"], dtype, grad=grad)
    sky_params = tt(g["sky_params"], dtype, grad=grad)
    beam_params = tt(g["beam_params"], dtype, grad=grad)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    sky = sky_params * float(g["px_area"])
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    beam_cache = orc.pixel_response_forward(beam_params, powerbeam=True)

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(g["theta_grid"], g["phi_grid"], z, a, interp_mode)
        return orc.interp_map(beam_cache, inds, wgts.to(dtype))

    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]))
    return V, dict(sky=sky_params, beam=beam_params, antvecs=antvecs)


def ylm_grid_matrix(g):
    """Ylm at the rect beam grid of the rime_ylm fixture, from the oracle's own recurrence."""
    ph, th = np.meshgrid(np.asarray(g["phi_grid"]), np.asarray(g["theta_grid"]))
    Y = orc.sph_harm_matrix(g["l"], g["m"], np.radians(th.ravel()), np.radians(ph.ravel()))
    return torch.as_tensor(Y)


def oracle_ylm(g, dtype=torch.float64):
    """rime_ylm fixture: a_lm --YlmResponse.forward--> beam_cache --bilinear interp--> RIME."""
    antvecs = tt(g["antvecs"], dtype, grad=True)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype, grad=True)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    sky = sky_params * float(g["px_area"])
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    Ylm = ylm_grid_matrix(g)
    beam_cache = orc.ylm_response_forward(beam_params, Ylm, tt(g["alm_mult"], dtype), powerbeam=True,
                                          beam0=tt(g["beam0"], dtype), comp_params=True)

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(g["theta_grid"], g["phi_grid"], z, a, 'linear')
        return orc.interp_map(beam_cache, inds, wgts.to(dtype))

    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]))
    return V, dict(sky=sky_params, beam=beam_params, antvecs=antvecs), beam_cache


def oracle_alm_sky(g, dtype=torch.float64):
    """rime_alm_sky fixture: a_lm --AlmModel--> pixel sky (+ sky0) x px_area --Airy beam--> RIME."""
    antvecs = tt(g["antvecs"], dtype, grad=True)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    Ylm = torch.as_tensor(orc.sph_harm_matrix(g["l"], g["m"], np.radians(90.0 - g["dec"]),
                                              np.radians(g["ra"])))
    # PixelSkyResponse.__call__ (sky_model.py:674-701): spatial transform, .real, + sky0
    sky = orc.alm_forward(sky_params, Ylm, tt(g["alm_mult"], dtype), real_output=True)
    sky = (sky + tt(g["sky0"], dtype)) * float(g["px_area"])
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    D = torch.ones(1, 1, 1, 1, 1, dtype=dtype) * 14.0
    beam_fn = lambda z, a: orc.airy_response(D, z, a, freqs, powerbeam=True)
    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]))
    return V, dict(sky=sky_params, antvecs=antvecs), Ylm


def oracle_2pol(g, dtype=torch.float64):
    antvecs = tt(g["antvecs"], dtype)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    beam_fn = lambda z, a: orc.airy_response(beam_params, z, a, freqs, powerbeam=True)
    V = orc.rime_forward(sky_params, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]))
    return V, dict(sky=sky_params)


def oracle_4pol(g, dtype=torch.float64):
    antvecs = tt(g["antvecs"], dtype, grad=True)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype, grad=True)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    sky = orc.stokes_to_coherency(sky_params * 1.0)
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    beam_cache = orc.pixel_response_forward(beam_params, powerbeam=False, realbeam=True)

    def beam_fn(z, a):
        inds, wgts = orc.rect_interp_weights(g["theta_grid"], g["phi_grid"], z, a, 'linear')
        return orc.interp_map(beam_cache, inds, wgts.to(dtype))

    V = orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]),
                         powerbeam=False)
    return V, dict(sky=sky_params, beam=beam_params, antvecs=antvecs)


def oracle_multimodel(g, dtype=torch.float64):
    antvecs = tt(g["antvecs"], dtype)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    a2b = {a: int(m) for a, m in zip(ants, g["ant2beam"])}
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    beam_fn = lambda z, a: orc.airy_response(beam_params, z, a, freqs, powerbeam=False)
    V = orc.rime_forward(sky_params, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]),
                         powerbeam=False, ant2beam=a2b)
    return V, dict(sky=sky_params)


def oracle_databls(g, dtype=torch.float64):
    antvecs = tt(g["antvecs"], dtype)
    sky_params = tt(g["sky_params"], dtype, grad=True)
    beam_params = tt(g["beam_params"], dtype)
    freqs = tt(g["freqs"], dtype)
    ants = [int(a) for a in g["ants"]]
    bls = bl_list(g["sim_bls"])
    blvecs = orc.get_blvecs(antvecs, ants, bls)
    zenaz = [(tt(za[0], dtype), tt(za[1], dtype)) for za in g["zen_az"]]
    beam_fn = lambda z, a: orc.airy_response(beam_params, z, a, freqs, powerbeam=True)
    V = orc.rime_forward(sky_params, zenaz, beam_fn, bls, blvecs, freqs, fov=float(g["fov"]),
                         sim2data=torch.as_tensor(g["sim2data"]))
    return V, dict(sky=sky_params)
