"""
CPU oracle for the BayesLIM RIME hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, in plain torch-on-CPU, the arithmetic of the reference's
``rime_model.RIME.forward`` and the model components it calls.  It follows the
reference's own formulation (a materialised (Nbl, Nfreq, Nsrc) complex fringe
tensor that is exponentiated, multiplied with the perceived sky and summed), so

  * torch autograd through it gives the reference's gradients, and
  * timing it on host cores is a faithful stand-in for the reference's CPU path
    (``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module.  The product (``bayeslim_b200``) never does.

Pinning: every function here is checked against outputs of the *unmodified*
reference (imported in the build container through astropy/h5py/healpy stub
modules, see ``tests/golden/make_golden.py``) stored under ``tests/golden/``.
Two pieces cannot be pinned that way because the reference delegates them to
third-party packages that are not installed (and not vendored):

  * ``eq2top`` (astropy ICRS->AltAz; reference ``telescope_model.py:469-502``):
    "parity unpinned".  The hot path consumes (zen, az), so tests and benches
    feed both implementations the same angles from :func:`eq2top_synth`.
  * HEALPix bilinear weights (``healpy.get_interp_weights``, called at
    reference ``utils.py:765-769``): "parity unpinned"; restated from the
    published HEALPix ring-interpolation algorithm in :func:`healpix_interp_weights`.

Reference citations are ``file:line`` relative to ``/root/reference/bayeslim``.
"""
import math

import numpy as np
import torch

C_LIGHT = 2.99792458e8          # telescope_model.py:355 (hard coded in the reference)
D2R = math.pi / 180.0           # utils.py:46


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def make_hex(N, D=15.0):
    """Hexagonal array of 3N^2-3N+1 antennas with spacing D [m].

    Restates utils.py:1943-1962 (``_make_hex``): 2N-1 rows, row i has
    N + min(i, 2N-2-i) antennas, rows offset by half a spacing, centred on the
    mean position.  Returns (ants, antvecs[Nant,3] float64 numpy).
    """
    xs, ys = [], []
    for i in range(2 * N - 1):
        n_in_row = N + min(i, 2 * N - 2 - i)
        start = -0.5 * min(i, 2 * N - 2 - i)
        for j in range(n_in_row):
            xs.append(start + j)
            ys.append(i * math.sin(math.pi / 3))
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    xs -= xs.mean()
    ys -= ys.mean()
    vecs = np.stack([xs, ys, np.zeros_like(xs)], axis=1) * D
    return list(range(len(xs))), vecs


def hera350():
    """Synthetic "HERA-350" of SURVEY section 8(d): hex core (N=11, D=14.6 m, 331 ants)
    plus 19 outriggers at radius 320+12k m, angle 2*pi*k/19."""
    ants, vecs = make_hex(11, D=14.6)
    k = np.arange(19)
    r = 320.0 + 12.0 * k
    ang = 2 * np.pi * k / 19
    out = np.stack([r * np.cos(ang), r * np.sin(ang), np.zeros(19)], axis=1)
    vecs = np.concatenate([vecs, out], axis=0)
    return list(range(len(vecs))), vecs


def cross_baselines(ants):
    """All (i, j) antenna pairs with i < j in list order."""
    return [(ants[i], ants[j]) for i in range(len(ants)) for j in range(i + 1, len(ants))]


def unique_baselines(ants, antvecs, tol=1.0):
    """One representative (i, j) per redundant group of baseline vectors.

    Simplified stand-in for ``telescope_model.build_reds`` (:693-942) that only
    produces the representative list used to build test inputs: two baselines
    are redundant if their vectors (or the negated vector) agree within tol.
    """
    reps, repvecs = [], []
    vecs = np.asarray(antvecs)
    for i in range(len(ants)):
        for j in range(i + 1, len(ants)):
            v = vecs[j] - vecs[i]
            found = False
            for rv in repvecs:
                if np.linalg.norm(v - rv) < tol or np.linalg.norm(v + rv) < tol:
                    found = True
                    break
            if not found:
                reps.append((ants[i], ants[j]))
                repvecs.append(v)
    return reps


def get_blvecs(antvecs, ants, bls):
    """b = antvecs[j] - antvecs[i] for bl (i, j): telescope_model.py:221-239."""
    idx = {a: k for k, a in enumerate(ants)}
    i0 = torch.as_tensor([idx[b[0]] for b in bls], dtype=torch.long)
    i1 = torch.as_tensor([idx[b[1]] for b in bls], dtype=torch.long)
    antvecs = torch.as_tensor(antvecs)
    return antvecs[i1] - antvecs[i0]


def eq2top_synth(time, ra, dec, lat=-30.72148, t0=2458148.0, lst0=1.0):
    """Deterministic stand-in for astropy's ICRS->AltAz.

    Hour-angle rotation at latitude ``lat`` [deg]; ra/dec in degrees, returns
    (zen, az) in degrees, az East of North.  NOT a restatement of astropy; both
    the oracle and the CUDA path are fed the same output.
    """
    ra = np.asarray(ra, dtype=np.float64) * D2R
    dec = np.asarray(dec, dtype=np.float64) * D2R
    phi = lat * D2R
    lst = 2 * np.pi * 1.0027379 * (float(time) - t0) + lst0
    H = lst - ra
    x = -np.cos(dec) * np.sin(H)
    y = np.cos(phi) * np.sin(dec) - np.sin(phi) * np.cos(dec) * np.cos(H)
    z = np.sin(phi) * np.sin(dec) + np.cos(phi) * np.cos(dec) * np.cos(H)
    zen = np.arccos(np.clip(z, -1.0, 1.0)) / D2R
    az = np.mod(np.arctan2(x, y), 2 * np.pi) / D2R
    return zen, az


def shat(zen, az, dtype=torch.float64):
    """Unit pointing vectors (3, Ns) from zen/az in degrees: telescope_model.py:337-343."""
    zen = torch.as_tensor(zen)
    az = torch.as_tensor(az)
    _zen = zen * D2R
    _az = az * D2R
    s = torch.zeros(3, len(zen), dtype=dtype)
    s[0] = torch.sin(_zen) * torch.sin(_az)
    s[1] = torch.sin(_zen) * torch.cos(_az)
    s[2] = torch.cos(_zen)
    return s


def gen_fringe(blvecs, zen, az, freqs, conj=False, dtype=torch.float64):
    """exp(+-2 pi i (b.s) nu / c) as an (Nbl, Nf, Ns) tensor: telescope_model.py:310-358."""
    s = shat(zen, az, dtype=dtype)
    sign = -2j if conj else 2j
    const = torch.as_tensor(freqs, dtype=dtype)[:, None] * (sign * math.pi / C_LIGHT)
    return ((blvecs.to(dtype) @ s)[:, None, :] * const).exp_()


def fov_cut(zen, fov):
    """Indices of sources with zen < fov/2 (strict); all if fov >= 360: beam_model.py:221-224."""
    zen = torch.as_tensor(zen)
    if fov < 360:
        return torch.where(zen < fov / 2)[0]
    return torch.arange(len(zen))


# --------------------------------------------------------------------------
# beams
# --------------------------------------------------------------------------
def j1_trapezoid(x, Ntau=100):
    """J1 from the Bessel integral (1/pi) int_0^pi cos(tau - x sin tau) dtau on Ntau nodes with
    end weights 1/2 (special.py:498-533, the brute_force branch).  Differentiable in x."""
    tau = torch.linspace(0, math.pi, Ntau, dtype=x.dtype)
    h = tau[1] - tau[0]
    acc = torch.zeros_like(x)
    for i in range(Ntau):
        w = 0.5 if i in (0, Ntau - 1) else 1.0
        acc = acc + w * torch.cos(tau[i] - x * torch.sin(tau[i]))
    return acc * h / math.pi


def airy_disk(zen, az, Dew, freqs, Dns=None, freq_ratio=1.0, square=True, brute_force=False,
              Ntau=100):
    """(2 J1(x)/x)^(2|1) with x = pi nu D(az) sin(min(zen, 90deg)) / c, clipped at 1e-10.

    zen, az in RADIANS (beam_model.py:1418-1482).  J1 = torch.special.bessel_j1
    (special.py:535), which has no autograd formula: the gradient that reaches
    Dew/Dns through this function is the truncated one, exactly as in the reference.
    """
    zen = torch.as_tensor(zen).clone()
    az = torch.as_tensor(az)
    zen[zen > math.pi / 2] = math.pi / 2
    if Dns is None:
        diameter = Dew
    else:
        diameter = Dns + torch.abs(torch.sin(az)) ** 2 * (Dew - Dns)
    freqs = torch.as_tensor(freqs)
    x = diameter * torch.sin(zen) * math.pi * freqs.reshape(-1, 1) * freq_ratio / C_LIGHT
    x = x.clip(1e-10)
    j1 = j1_trapezoid(x, Ntau) if brute_force else torch.special.bessel_j1(x)
    beam = 2.0 * j1 / x
    if square:
        beam = beam ** 2
    return beam


def airy_response(params, zen, az, freqs, freq_ratio=1.0, powerbeam=True, brute_force=False,
                  Ntau=100):
    """AiryResponse.__call__ (beam_model.py:956-985): zen/az in DEGREES, params
    (Npol, Nvec, Nmodel, 1, 1|2) -> beam (Npol, Nvec, Nmodel, Nf, Ns)."""
    Dew = params[..., 0:1]
    Dns = params[..., 1:2] if params.shape[-1] > 1 else None
    zen = torch.as_tensor(zen)
    az = torch.as_tensor(az)
    return airy_disk(zen * D2R, az * D2R, Dew, freqs, Dns, freq_ratio, square=powerbeam,
                     brute_force=brute_force, Ntau=Ntau)


def gauss_response(params, zen, az, powerbeam=True):
    """GaussResponse.__call__ (beam_model.py:886-899); zen/az degrees."""
    zen_rad = torch.as_tensor(zen) * D2R
    az_rad = torch.as_tensor(az) * D2R
    srad = torch.sin(zen_rad)
    srad = torch.where(zen_rad > math.pi / 2, torch.ones_like(srad), srad)
    l = srad * torch.sin(az_rad)
    m = srad * torch.cos(az_rad)
    beam = torch.exp(-0.5 * ((l / params[..., 0:1]) ** 2 + (m / params[..., 1:2]) ** 2))
    if not powerbeam:
        beam = torch.sqrt(beam)
    return beam


_S2D = {'nearest': 0, 'linear': 1, 'quadratic': 2, 'cubic': 3}


def _nearest_nodes(grid, xnew, n, wrap):
    """Sorted indices of the n grid nodes nearest to each xnew, as the reference
    picks them (argsort of |grid - x|, utils.py:1003-1004), plus the position of
    xnew relative to the first picked node in units of the grid step."""
    grid = torch.as_tensor(grid, dtype=torch.float64)
    xnew = torch.as_tensor(xnew, dtype=torch.float64)
    N = len(grid)
    dx = grid[1] - grid[0]
    if wrap:
        ext = torch.cat([grid[-n:] - N * dx, grid, grid[:n] + N * dx])
    else:
        ext = grid
    order = torch.argsort(torch.abs(ext - xnew[:, None]), dim=-1)[:, :n]
    nn = torch.sort(order, dim=-1).values
    rel = (xnew - ext[nn[:, 0]]) / dx
    if wrap:
        nn = (nn - n) % N
    return nn, rel


def _lagrange(rel, n):
    """Weights of the degree-(n-1) polynomial through nodes 0..n-1 evaluated at rel.

    The reference solves the monomial least-squares system (utils.py:1084-1116);
    for a square, full-rank design matrix that is exact polynomial interpolation,
    whose weights are the Lagrange basis polynomials."""
    w = []
    for i in range(n):
        num = torch.ones_like(rel)
        den = 1.0
        for j in range(n):
            if j != i:
                num = num * (rel - j)
                den *= (i - j)
        w.append(num / den)
    return torch.stack(w, dim=-1)


def rect_interp_weights(theta_grid, phi_grid, zen, az, interp_mode='linear'):
    """(inds, wgts), each (Ns, Nnn), for interpolation on a uniform (phi, theta) grid.

    Restates PixInterp.get_interp 'rect' branch (utils.py:772-798) with
    bipoly_grid_index (:949-1021, az wraps, zen does not) and
    setup_bipoly_interp (:1024-1116).  Flat index = ix + Nphi*iy, neighbours
    ordered with ix fastest.  zen, az, grids in degrees.
    """
    if ',' in interp_mode:
        deg = [_S2D[s.strip()] for s in interp_mode.split(',')]
    else:
        deg = [_S2D[interp_mode]] * 2
    nx, ny = deg[0] + 1, deg[1] + 1
    xnn, xrel = _nearest_nodes(phi_grid, az, nx, wrap=True)
    ynn, yrel = _nearest_nodes(theta_grid, zen, ny, wrap=False)
    Nphi = len(phi_grid)
    inds = (xnn[:, None, :] + Nphi * ynn[:, :, None]).reshape(len(xrel), -1)
    wx = _lagrange(xrel, nx)
    wy = _lagrange(yrel, ny)
    wgts = (wy[:, :, None] * wx[:, None, :]).reshape(len(xrel), -1)
    return inds, wgts


def interp_map(m, inds, wgts):
    """out[..., s] = sum_i m[..., inds[s, i]] * wgts[s, i]: utils.py:833-841."""
    nearest = m.index_select(-1, inds.reshape(-1)).view(m.shape[:-1] + inds.shape)
    return torch.einsum('...i,...i->...', nearest, wgts.to(nearest.dtype))


# ---- HEALPix (RING) helpers: closed-form restatements of healpy calls -----
def healpix_npix(nside):
    return 12 * nside * nside


def healpix_pixarea(nside):
    return 4 * math.pi / healpix_npix(nside)


def healpix_ring_info(nside):
    """Per-ring (start pixel, n pixels, cos(theta), phi shift flag) for RING ordering."""
    nring = 4 * nside - 1
    start = np.zeros(nring, dtype=np.int64)
    npr = np.zeros(nring, dtype=np.int64)
    z = np.zeros(nring, dtype=np.float64)
    shifted = np.zeros(nring, dtype=bool)
    ncap = 2 * nside * (nside - 1)
    npix = healpix_npix(nside)
    for r in range(1, 4 * nside):
        if r < nside:                       # north cap
            npr[r - 1] = 4 * r
            start[r - 1] = 2 * r * (r - 1)
            z[r - 1] = 1.0 - r * r / (3.0 * nside * nside)
            shifted[r - 1] = True
        elif r <= 3 * nside:                # equatorial belt
            npr[r - 1] = 4 * nside
            start[r - 1] = ncap + (r - nside) * 4 * nside
            z[r - 1] = (2 * nside - r) * 2.0 / (3.0 * nside)
            shifted[r - 1] = ((r - nside) % 2 == 0)
        else:                               # south cap
            rr = 4 * nside - r
            npr[r - 1] = 4 * rr
            start[r - 1] = npix - 2 * rr * (rr + 1)
            z[r - 1] = -1.0 + rr * rr / (3.0 * nside * nside)
            shifted[r - 1] = True
    return start, npr, z, shifted


def healpix_pix2ang(nside):
    """(theta, phi) [rad] of all RING-ordered pixel centres (healpy.pix2ang)."""
    start, npr, z, shifted = healpix_ring_info(nside)
    theta = np.zeros(healpix_npix(nside))
    phi = np.zeros(healpix_npix(nside))
    for r in range(len(start)):
        j = np.arange(npr[r])
        shift = 0.5 if shifted[r] else 0.0
        theta[start[r]:start[r] + npr[r]] = math.acos(z[r])
        phi[start[r]:start[r] + npr[r]] = (j + shift) * 2 * np.pi / npr[r]
    return theta, phi


def healpix_interp_weights(nside, theta, phi):
    """Bilinear RING interpolation: (inds, wgts) of shape (Ns, 4).

    Restates the published HEALPix ``get_interpol`` algorithm that
    ``healpy.get_interp_weights`` wraps (reference call site utils.py:765-769):
    pick the rings above/below theta; in each ring interpolate linearly in phi
    between the two bracketing pixels; combine the rings linearly in theta.
    Beyond the first/last ring, blend the ring pair with the mean of the four
    polar pixels.  theta, phi in radians.  PARITY UNPINNED (healpy absent).
    """
    start, npr, z, shifted = healpix_ring_info(nside)
    ring_theta = np.arccos(z)
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.mod(np.asarray(phi, dtype=np.float64), 2 * np.pi)
    ns = len(theta)
    inds = np.zeros((ns, 4), dtype=np.int64)
    wgts = np.zeros((ns, 4), dtype=np.float64)
    nring = len(start)
    npix = healpix_npix(nside)
    # ring index (0-based) of the last ring with ring_theta <= theta; -1 if above the first ring
    ir1 = np.searchsorted(ring_theta, theta, side='right') - 1

    def ring_pair(r, ph):
        n = npr[r]
        dphi = 2 * np.pi / n
        shift = 0.5 if shifted[r] else 0.0
        t = ph / dphi - shift
        i1 = np.floor(t).astype(np.int64)
        w = t - i1
        i2 = i1 + 1
        i1 = np.mod(i1, n)
        i2 = np.mod(i2, n)
        return start[r] + i1, start[r] + i2, 1.0 - w, w

    for k in range(ns):
        r1 = ir1[k]
        ph = phi[k]
        if r1 < 0:                           # north polar cap above ring 1
            p1, p2, w1, w2 = ring_pair(0, ph)
            wt = theta[k] / ring_theta[0]
            inds[k] = [0, 1, 2, 3]
            wgts[k] = (1 - wt) * 0.25
            # the ring-1 pair pixels are among (0..3): add their weights
            wgts[k, p1] += wt * w1
            wgts[k, p2] += wt * w2
        elif r1 >= nring - 1:                # south polar cap below last ring
            p1, p2, w1, w2 = ring_pair(nring - 1, ph)
            wt = (theta[k] - ring_theta[-1]) / (math.pi - ring_theta[-1])
            base = npix - 4
            inds[k] = [base, base + 1, base + 2, base + 3]
            wgts[k] = wt * 0.25
            wgts[k, p1 - base] += (1 - wt) * w1
            wgts[k, p2 - base] += (1 - wt) * w2
        else:
            a1, a2, wa1, wa2 = ring_pair(r1, ph)
            b1, b2, wb1, wb2 = ring_pair(r1 + 1, ph)
            wt = (theta[k] - ring_theta[r1]) / (ring_theta[r1 + 1] - ring_theta[r1])
            inds[k] = [a1, a2, b1, b2]
            wgts[k] = [(1 - wt) * wa1, (1 - wt) * wa2, wt * wb1, wt * wb2]
    return torch.as_tensor(inds), torch.as_tensor(wgts)


def pixel_response_forward(params, powerbeam=True, realbeam=True, log=False, beam0=None,
                           comp_params=False):
    """PixelResponse.forward for freq_mode='channel', no LM/taper/norm: beam_model.py:750-793."""
    p = params
    if comp_params and not torch.is_complex(p):
        p = torch.view_as_complex(p)
    if powerbeam or realbeam:
        p = p.real
    if log:
        p = torch.exp(p)
    elif powerbeam:
        p = torch.abs(p)
    if beam0 is not None:
        p = p + beam0
    return p


# --------------------------------------------------------------------------
# sky
# --------------------------------------------------------------------------
def point_sky_response(params, freqs, freq_mode='channel', f0=None, log=False):
    """PointSkyResponse.__call__ (sky_model.py:340-366), modes 'channel' and 'powerlaw'."""
    if freq_mode == 'channel':
        out = params
        if log:
            out = torch.exp(out)
        return out
    if freq_mode == 'powerlaw':
        amp = params[..., 0:1, :]
        if log:
            amp = torch.exp(amp)
        return amp * (torch.as_tensor(freqs)[:, None] / f0) ** params[..., 1:2, :]
    raise NotImplementedError(freq_mode)


def stokes_to_coherency(sky):
    """Stokes2Coherency.forward for a (Nstokes, 1, Nf, Ns) input holding
    [I, fQ, fU(, fV)] with Q = I*fQ etc. (sky_model.py:1284-1313)."""
    I = sky[0, 0]
    n = len(sky)
    if n == 1:
        return sky
    Q = I * sky[1, 0]
    U = I * sky[2, 0] if n > 2 else torch.zeros_like(I)
    has_v = n > 3
    dtype = torch.complex128 if (has_v and I.dtype == torch.float64) else (
        torch.complex64 if has_v else I.dtype)
    B = torch.zeros(2, 2, *sky.shape[2:], dtype=dtype)
    B[0, 0] = I + Q
    B[0, 1] = U
    B[1, 0] = U
    B[1, 1] = I - Q
    if has_v:
        V = I * sky[3, 0]
        B[0, 1] = B[0, 1] - 1j * V
        B[1, 0] = B[1, 0] + 1j * V
    return B


# --------------------------------------------------------------------------
# RIME
# --------------------------------------------------------------------------
def apply_beam(beam, sky, bls, powerbeam=True, ant2beam=None):
    """Perceived sky (Npol, Npol, Nbl, Nf, Ns): beam_model.py:273-372.

    beam (Npol, Nvec, Nmodel, Nf, Ns); sky (Nvec, Nvec, Nf, Ns); ant2beam maps
    antenna number -> model index (None: all 0).  Operands of the polarised
    einsum are promoted to a common dtype first (the reference raises a dtype
    error otherwise in torch 2.11 -- SURVEY section 9 item 7; promotion does not
    change the arithmetic).
    """
    Npol, Nvec = beam.shape[0], beam.shape[1]
    a2b = (lambda a: 0) if ant2beam is None else (lambda a: ant2beam[a])
    bl2mp = {bl: (a2b(bl[0]), a2b(bl[1])) for bl in bls}
    modelpairs = sorted(set(bl2mp.values()))
    i1 = torch.as_tensor([mp[0] for mp in modelpairs])
    i2 = torch.as_tensor([mp[1] for mp in modelpairs])
    beam1 = torch.index_select(beam, 2, i1)
    beam2 = torch.index_select(beam, 2, i2)
    sky = sky[:, :, None]
    if Npol == 1 and Nvec == 1:
        if powerbeam:
            psky = beam1 * sky
        else:
            psky = (beam1 * beam2.conj()) * sky
    elif Npol == 2 and powerbeam:
        psky = torch.zeros(2, 1, *torch.broadcast_shapes(beam1.shape[2:], sky.shape[2:]),
                           dtype=torch.result_type(beam1, sky))
        psky[0] = beam1[0, 0] * sky[0, 0]
        psky[1] = beam1[1, 0] * sky[0, 0]
    else:
        dt = torch.result_type(beam1, sky)
        psky = torch.einsum("ab...,bc...,dc...->ad...", beam1.to(dt), sky.to(dt),
                            beam2.conj().to(dt))
    mp_idx = torch.as_tensor([modelpairs.index(bl2mp[bl]) for bl in bls])
    return torch.index_select(psky, 2, mp_idx)


def prod_and_sum(beam, cut_sky, bls, blvecs, zen, az, freqs, powerbeam=True,
                 ant2beam=None, conj=False, bl_chunk=None):
    """sum_s fringe * psky -> (Npol, Npol, Nbl, Nf): rime_model.py:391-440.

    bl_chunk bounds the materialised (Nbl, Nf, Ns) tensor (the reference's own
    remedy is baseline minibatching, rime_model.py:55-57); the arithmetic per
    baseline is unchanged.
    """
    Nbl = len(bls)
    bl_chunk = bl_chunk or Nbl
    out = []
    for b0 in range(0, Nbl, bl_chunk):
        sl = slice(b0, min(b0 + bl_chunk, Nbl))
        psky = apply_beam(beam, cut_sky, bls[sl], powerbeam=powerbeam, ant2beam=ant2beam)
        fringe = gen_fringe(blvecs[sl], zen, az, freqs, conj=conj,
                            dtype=blvecs.dtype)
        out.append(torch.sum(fringe * psky, dim=-1))
    return torch.cat(out, dim=2)


def pointing_offset(zen, az, theta_x=0.0, theta_y=0.0):
    """(zen, az) [deg] at which the beam response is evaluated for a beam with a small-angle
    pointing offset: beam_model.py:244-256 + pointing_offset :1631-1678 (unit vector
    (sin t cos p, sin t sin p, cos t), rotation about x-hat by theta_x then about y-hat by
    theta_y, each only when > 0 (`rotation`, :1514-1545), back to angles with new_phi in
    [0, 2 pi)).  Non-differentiable (numpy), as in the reference."""
    if not (theta_x > 0 or theta_y > 0):
        return zen, az
    t = np.asarray(torch.as_tensor(zen).detach().cpu().double()) * D2R
    p = np.asarray(torch.as_tensor(az).detach().cpu().double()) * D2R
    r = np.array([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)])
    if theta_x > 0:
        c, s = math.cos(theta_x), math.sin(theta_x)
        r = np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]]) @ r
    if theta_y > 0:
        c, s = math.cos(theta_y), math.sin(theta_y)
        r = np.array([[c, 0, s], [0, 1.0, 0], [-s, 0, c]]) @ r
    nt = np.arccos(r[2])
    xzero, yzero = np.isclose(r[0], 0), np.isclose(r[1], 0)
    xneg, ypos = r[0] < 0, r[1] > 0
    nphi = np.zeros_like(nt)
    nphi[~xzero] = np.arctan(r[1][~xzero] / r[0][~xzero])
    nphi[xneg & ypos] += np.pi
    nphi[xneg & ~ypos] -= np.pi
    nphi[xzero & yzero] = 0.0
    nphi[xzero & ypos] = np.pi / 2
    nphi[xzero & ~ypos] = -np.pi / 2
    nphi = nphi % (2 * np.pi)
    return torch.as_tensor(nt / D2R), torch.as_tensor(nphi / D2R)


def rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=180.0, powerbeam=True,
                 ant2beam=None, sim2data=None, bl_chunk=None, offset=(0.0, 0.0)):
    """The reference time loop (rime_model.py:326-368) for one sky component.

    sky      : (Nvec, Nvec, Nf, Npix) coherency / (1,1,Nf,Npix) Stokes-I tensor
    zenaz    : list over times of (zen, az) [deg] for all Npix sources
    beam_fn  : callable (zen_cut, az_cut) -> (Npol, Nvec, Nmodel, Nf, Ns) beam
    returns V: (Npol, Npol, Nbl|Ndata, Nt, Nf)
    """
    vis = []
    for zen, az in zenaz:
        zen = torch.as_tensor(zen)
        az = torch.as_tensor(az)
        cut = fov_cut(zen, fov)
        zc, ac = zen[cut], az[cut]
        beam = beam_fn(*pointing_offset(zc, ac, *offset))     # the fringe keeps (zc, ac)
        cut_sky = sky.index_select(-1, cut)
        v = prod_and_sum(beam, cut_sky, bls, blvecs, zc, ac, freqs, powerbeam=powerbeam,
                         ant2beam=ant2beam, bl_chunk=bl_chunk)
        if sim2data is not None:
            v = torch.index_select(v, 2, sim2data)
        vis.append(v)
    return torch.stack(vis, dim=3)


# --------------------------------------------------------------------------
# imaging: the adjoint of the same operator
# --------------------------------------------------------------------------
def imaging_matrix(blvecs, zen, az, freqs, beam=None):
    """A (Nbl, Nf, Ns) = conj(fringe) * beam: VisMapper.build_A, imaging.py:289-296."""
    A = gen_fringe(blvecs, zen, az, freqs, conj=True)
    if beam is not None:
        A = A * beam.to(A.dtype)[None]
    return A


def vismapper_make_map(v, w, blvecs, zenaz, freqs, npix, beam_fn=None, fov=180.0, method='A2w',
                       clip=1e-8):
    """VisMapper.make_map with contract='diag' (imaging.py:362-478) restated.

    v (Nbl, Nt, Nf) complex visibilities, w (Nbl, Nt, Nf) weights or None (= 1), zenaz list over
    times of (zen, az) [deg] of all Npix pixels, beam_fn(zen_cut, az_cut) -> (Nf, Ns) power beam
    or None.  With a beam the cut is the beam's strict zen < fov/2 (beam_model.py:221-224),
    without it zen <= fov/2 (imaging.py:285).  Returns (maps, P, D), each (Nf, Npix)."""
    freqs = torch.as_tensor(freqs, dtype=torch.float64)
    nf = len(freqs)
    maps = torch.zeros(nf, npix, dtype=torch.float64)
    P = torch.zeros(nf, npix, dtype=torch.float64)
    Aw = torch.zeros(nf, 1 if method == 'w' else npix, dtype=torch.float64)
    for t, (zen, az) in enumerate(zenaz):
        zen, az = torch.as_tensor(zen), torch.as_tensor(az)
        if beam_fn is not None:
            cut = fov_cut(zen, fov)
            beam = beam_fn(zen[cut], az[cut])
        else:
            cut = torch.where(zen <= fov / 2)[0]
            beam = None
        A = imaging_matrix(blvecs, zen[cut], az[cut], freqs, beam)
        wt = torch.ones(v.shape[0], nf, dtype=torch.float64) if w is None else w[:, t].double()
        vt = v[:, t].to(A.dtype)
        maps[:, cut] += torch.einsum('vfp,vf->fp', A, vt * wt).real          # imaging.py:736
        P[:, cut] += (wt[..., None] * A.abs().pow(2)).sum(0).real           # imaging.py:850
        if method == 'w':
            Aw += wt.sum(0)[:, None]                                         # imaging.py:460
        elif method == 'Aw':
            Aw[:, cut] += (wt[:, :, None] * A.abs()).sum(0)                  # imaging.py:462
        else:
            Aw[:, cut] += (wt[:, :, None] * A.pow(2).real).sum(0)            # imaging.py:464
    D = 1 / Aw.clip(clip)                                                    # imaging.py:469
    return maps * D, P * D, D


def vismapper_compute_Am(maps, blvecs, zenaz, freqs, beam_fn=None, fov=180.0):
    """VisMapper.compute_Am (imaging.py:480-540): v[m, b, t, f] = sum_p conj(A[b, f, p]) maps[m, f, p]."""
    freqs = torch.as_tensor(freqs, dtype=torch.float64)
    out = []
    for zen, az in zenaz:
        zen, az = torch.as_tensor(zen), torch.as_tensor(az)
        if beam_fn is not None:
            cut = fov_cut(zen, fov)
            beam = beam_fn(zen[cut], az[cut])
        else:
            cut = torch.where(zen <= fov / 2)[0]
            beam = None
        A = imaging_matrix(blvecs, zen[cut], az[cut], freqs, beam)
        out.append(torch.einsum("vfp,mfp->mvf", A.conj(), maps[..., cut].to(A.dtype)))
    return torch.stack(out, dim=2)


# --------------------------------------------------------------------------
# gain application
# --------------------------------------------------------------------------
def apply_cal(vis, gains, g1_idx, g2_idx, cal_2pol=False, cov=None, undo=False):
    """calibration._apply_cal for complex visibilities (calibration.py:2412-2487):
    1pol / 2pol: vout[p][p] = g1[p][p] conj(g2[p][p]) vis[p][p] with the off-diagonal outputs zeroed
    (linalg.diag_matmul, linalg.py:116-149), cov_out = |g1 conj g2|^2 cov (:2470-2476);
    4pol: vout[a][d] = sum_{b,c} g1[a][b] vis[b][c] conj(g2[d][c]) (:2485).
    undo (diagonal modes): gains -> 1 / gains on the diagonal, zero elsewhere (linalg.diag_inv)."""
    polmode = '1pol' if tuple(vis.shape[:2]) == (1, 1) else ('2pol' if cal_2pol else '4pol')
    if undo:
        assert polmode != '4pol', "the reference's 4pol undo does not run (torch.pinv)"
        inv = torch.zeros_like(gains)
        for p in range(gains.shape[0]):
            inv[p, p] = 1 / gains[p, p]
        gains = inv
    g1 = gains.index_select(2, g1_idx)
    g2 = gains.index_select(2, g2_idx)
    cov_out = cov
    if polmode in ('1pol', '2pol'):
        G = g1 * g2.conj()
        vout = torch.zeros_like(vis)
        for p in range(vis.shape[0]):
            vout[p, p] = G[p, p] * vis[p, p]
        if cov is not None:
            GG = (G * G.conj()).real
            cov_out = torch.zeros_like(cov)
            for p in range(vis.shape[0]):
                cov_out[p, p] = GG[p, p] * cov[p, p]
    else:
        vout = torch.einsum("ab...,bc...,dc...->ad...", g1, vis, g2.conj())
    return vout, cov_out


def jones_gains(params, param_type, freqs=None, refant_idx=None):
    """JonesModel's parameter -> gain map for channel-mode parameters (calibration.py:587-597
    fix_refant_phs with mode 'rephase' :2541-2574, then JonesResponse.params2complex :816-865 over
    params2complex :215-251).  Returns (rephased params, complex gains)."""
    p = params
    if refant_idx is not None:
        ref = p[:, :, refant_idx:refant_idx + 1]
        if param_type == 'com':
            p = p / torch.exp(1j * torch.angle(ref).detach())
        elif param_type in ('dly', 'phs'):
            p = p - ref
        elif param_type == 'amp_phs':
            p = torch.stack([p[..., 0], p[..., 1] - ref[..., 1]], dim=-1)
    if param_type == 'com':
        g = p
    elif param_type == 'real':
        g = p + 0j
    elif param_type == 'amp':
        g = torch.exp(p) + 0j
    elif param_type == 'phs':
        g = torch.exp(1j * p)
    elif param_type == 'amp_phs':
        g = torch.exp(p[..., 0] + 1j * p[..., 1])
    elif param_type == 'dly':
        g = torch.exp(2j * math.pi * p * (torch.as_tensor(freqs, dtype=p.dtype) / 1e9))
    else:
        raise ValueError(param_type)
    return p, g


# --------------------------------------------------------------------------
# apparent-place eq2top -- PARITY UNPINNED against astropy
# --------------------------------------------------------------------------
def eq2top_apparent(jd_utc, ra_deg, dec_deg, lon_deg, lat_deg):
    """ICRS/J2000 (ra, dec) -> (zen, az) [deg] in the classical spherical-trigonometry form
    (Meeus, Astronomical Algorithms, ch. 21-23): rigorous IAU 1976 precession of (alpha, delta),
    first-order nutation (23.1) and annual aberration (23.2, eccentricity terms dropped) corrections,
    apparent sidereal time, then altitude / azimuth (13.5, 13.6).  Stands in for the astropy call of
    telescope_model.py:469-502, which cannot be pinned here (astropy absent, no golden values in
    the reference); the product builds the same transformation as one rotation matrix
    (telescope_model.icrs_to_enu), so the two derivations check each other to second order in the
    20-arcsecond corrections."""
    asec = np.pi / 180 / 3600
    d2r = np.pi / 180
    a0 = np.asarray(ra_deg, dtype=np.float64) * d2r
    d0 = np.asarray(dec_deg, dtype=np.float64) * d2r
    T = (jd_utc + 69.184 / 86400.0 - 2451545.0) / 36525.0
    zeta = (2306.2181 * T + 0.30188 * T ** 2 + 0.017998 * T ** 3) * asec
    z = (2306.2181 * T + 1.09468 * T ** 2 + 0.018203 * T ** 3) * asec
    th = (2004.3109 * T - 0.42665 * T ** 2 - 0.041833 * T ** 3) * asec
    A = np.cos(d0) * np.sin(a0 + zeta)
    B = np.cos(th) * np.cos(d0) * np.cos(a0 + zeta) - np.sin(th) * np.sin(d0)
    C = np.sin(th) * np.cos(d0) * np.cos(a0 + zeta) + np.cos(th) * np.sin(d0)
    al = np.arctan2(A, B) + z
    de = np.arcsin(np.clip(C, -1, 1))
    Om = (125.04452 - 1934.136261 * T) * d2r
    Ls = (280.4665 + 36000.7698 * T) * d2r
    Lm = (218.3165 + 481267.8813 * T) * d2r
    Ms = (357.52772 + 35999.050340 * T) * d2r
    Mm = (134.96298 + 477198.867398 * T) * d2r
    dpsi = (-17.1996 * np.sin(Om) - 1.3187 * np.sin(2 * Ls) - 0.2274 * np.sin(2 * Lm)
            + 0.2062 * np.sin(2 * Om) + 0.1426 * np.sin(Ms) + 0.0712 * np.sin(Mm)
            - 0.0517 * np.sin(2 * Ls + Ms) - 0.0386 * np.sin(2 * Lm - Om)
            - 0.0301 * np.sin(2 * Lm + Mm)) * asec
    deps = (9.2025 * np.cos(Om) + 0.5736 * np.cos(2 * Ls) + 0.0977 * np.cos(2 * Lm)
            - 0.0895 * np.cos(2 * Om) + 0.0054 * np.cos(Ms) + 0.0224 * np.cos(2 * Ls + Ms)
            + 0.0200 * np.cos(2 * Lm - Om) + 0.0129 * np.cos(2 * Lm + Mm)) * asec
    eps0 = (84381.448 - 46.8150 * T - 0.00059 * T ** 2 + 0.001813 * T ** 3) * asec
    eps = eps0 + deps
    lam = Ls + (1.914602 - 0.004817 * T) * d2r * np.sin(Ms) + 0.019993 * d2r * np.sin(2 * Ms)
    kap = 20.49552 * asec
    # nutation (Meeus 23.1) and aberration (23.2) corrections of the mean place of date
    da = ((np.cos(eps0) + np.sin(eps0) * np.sin(al) * np.tan(de)) * dpsi - np.cos(al) * np.tan(de) * deps
          - kap * (np.cos(al) * np.cos(lam) * np.cos(eps0) + np.sin(al) * np.sin(lam)) / np.cos(de))
    dd = (np.sin(eps0) * np.cos(al) * dpsi + np.sin(al) * deps
          - kap * (np.cos(lam) * np.cos(eps0) * (np.tan(eps0) * np.cos(de) - np.sin(al) * np.sin(de))
                   + np.cos(al) * np.sin(de) * np.sin(lam)))
    al, de = al + da, de + dd
    d = jd_utc - 2451545.0
    Tu = d / 36525.0
    gmst = (280.46061837 + 360.98564736629 * d + 0.000387933 * Tu ** 2 - Tu ** 3 / 38710000.0) * d2r
    H = gmst + dpsi * np.cos(eps) + lon_deg * d2r - al
    phi = lat_deg * d2r
    sin_alt = np.sin(phi) * np.sin(de) + np.cos(phi) * np.cos(de) * np.cos(H)
    zen = np.arccos(np.clip(sin_alt, -1, 1)) / d2r
    az = np.mod(np.arctan2(-np.cos(de) * np.sin(H),
                           np.sin(de) * np.cos(phi) - np.cos(de) * np.cos(H) * np.sin(phi)), 2 * np.pi) / d2r
    return zen, az


# --------------------------------------------------------------------------
# spherical-harmonic beam
# --------------------------------------------------------------------------
def sph_harm_matrix(l, m, theta, phi):
    """Orthonormal Y_lm(theta, phi) (Ncoeff, Npix), complex128, theta / phi in RADIANS:
    sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(cos theta) exp(i m phi) with the Condon-Shortley
    phase -- what gen_sph2pix(method='sphere') returns (sph_harm.py:255-475).  Independent of
    scipy: normalised associated Legendre functions by the standard three-term recurrence
    in l started from P_m^m."""
    l = np.asarray(l).astype(int)
    m = np.asarray(m).astype(int)
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    x, s = np.cos(theta), np.sin(theta)
    out = np.zeros((len(l), len(theta)), dtype=np.complex128)
    for i, (li, mi) in enumerate(zip(l, m)):
        ma = abs(mi)
        # normalised P_ma^ma = (-1)^ma sqrt((2ma+1)!! / (4 pi (2ma)!!)) sin^ma
        pmm = np.full_like(x, math.sqrt(1.0 / (4 * math.pi)))
        for k in range(1, ma + 1):
            pmm = -pmm * s * math.sqrt((2 * k + 1) / (2.0 * k))
        if li == ma:
            p = pmm
        else:
            pm1 = x * math.sqrt(2 * ma + 3) * pmm
            p0, p1 = pmm, pm1
            for ll in range(ma + 2, li + 1):
                a = math.sqrt((4.0 * ll * ll - 1) / (ll * ll - ma * ma))
                b = math.sqrt(((ll - 1.0) ** 2 - ma * ma) / (4.0 * (ll - 1) ** 2 - 1))
                p0, p1 = p1, a * (x * p1 - b * p0)
            p = p1
        y = p * np.exp(1j * ma * phi)
        if mi < 0:
            y = (-1) ** ma * np.conj(y)
        out[i] = y
    return out


def alm_forward(params, Ylm, alm_mult=None, real_output=False):
    """AlmModel.forward_alm, non-separable branch (sph_harm.py:1344-1373): optional
    multiplication of the coefficients, einsum "...i,ij->...j", real part if real_output."""
    if torch.is_complex(Ylm) and not torch.is_complex(params):
        params = torch.view_as_complex(params)
    if alm_mult is not None:
        params = params * alm_mult
    out = torch.einsum("...i,ij->...j", params.to(torch.result_type(params, Ylm)), Ylm)
    return out.real if real_output else out


def alm_forward_separable(params, Theta, Phi, alm_mult=None, real_output=False):
    """AlmModel.forward_alm, separable branch (sph_harm.py:1354-1362): (..., Ncoeff) ->
    (..., Ntheta * Nphi)."""
    if torch.is_complex(Phi) and not torch.is_complex(params):
        params = torch.view_as_complex(params)
    if alm_mult is not None:
        params = params * alm_mult
    x = torch.einsum("ct,...c->...tc", Theta.to(params.dtype), params)
    x = torch.einsum("...tc,cp->...tp", x, Phi)
    out = x.reshape(x.shape[:-2] + (Theta.shape[1] * Phi.shape[1],))
    return out.real if real_output else out


def ylm_response_forward(params, Ylm, alm_mult=None, powerbeam=True, realbeam=True, log=False,
                         beam0=None, comp_params=False):
    """YlmResponse.forward (beam_model.py:1166-1233) for freq_mode 'channel', no taper / norm:
    a_lm (Npol, Nvec, Nmodel, Nf, Ncoeff[, 2]) -> beam map (Npol, Nvec, Nmodel, Nf, Npix)."""
    if comp_params and not torch.is_complex(params):
        params = torch.view_as_complex(params)
    realbeam = True if powerbeam else realbeam
    beam = alm_forward(params, Ylm, alm_mult, real_output=realbeam)
    if log:
        beam = torch.exp(beam)
    elif powerbeam:
        beam = torch.abs(beam)
    if beam0 is not None:
        beam = beam + beam0
    return beam
