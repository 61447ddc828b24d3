"""
Host-side mirror of the reference's telescope_model (bayeslim/telescope_model.py):
``TelescopeModel`` (sky -> topocentric angles, cached) and ``ArrayModel`` (antenna
layout, baseline vectors, redundancy bookkeeping, fringe).

Differences that matter:
  * ``TelescopeModel.eq2top`` answers from ``conv_cache`` (the reference's own cache, keyed
    exactly as rime_model.py:345 builds the key) and, on a miss, from astropy's ICRS -> AltAz
    when astropy is importable (the reference's own call).  Without astropy (this image) it
    warns once and computes apparent places itself (IAU 1976 precession, IAU 1980 nutation
    leading terms, annual aberration, apparent sidereal time; arcsecond-grade, parity with
    astropy is unpinned): the per-time rotation is built on the host, the
    per-source part runs in the CUDA kernel b200rime_eq2top when the model lives on a GPU;
    ``eq2top_fn`` overrides both.
  * ``ArrayModel.gen_fringe`` keeps the reference signature and semantics for callers
    outside the RIME (imaging), but ``rime_model.RIME`` never calls it: the fringe is
    generated inside the CUDA kernels (ops.fringe_sum).
"""
import copy
import itertools

import numpy as np
import torch

from . import utils
from .utils import _float, D2R


class TelescopeModel:
    """Telescope location + equatorial -> topocentric conversion cache
    (telescope_model.py:20-137)."""

    def __init__(self, location, tloc=None, device=None, dtype=None, eq2top_fn=None):
        self.location = location
        self.tloc = tloc
        self.dtype = dtype
        self.conv_cache = {}
        self.device = device
        self.eq2top_fn = eq2top_fn

    def hash(self, time, ra):
        return (time, len(ra))

    def clear_cache(self, key=None):
        if key is None:
            self.conv_cache = {}
        else:
            del self.conv_cache[key]

    def eq2top(self, time, ra, dec, store=False, key=None):
        """(zen, az) [deg] stacked as a (2, Nsrc) tensor; az is East of North."""
        key = key if key is not None else self.hash(time, ra)
        if key in self.conv_cache:
            return self.conv_cache[key]
        dev = torch.device(self.device) if self.device is not None else None
        if self.eq2top_fn is None and not _have_astropy() and dev is not None and dev.type == 'cuda':
            # the per-source part on the device
            _warn_no_astropy()
            angs = eq2top_device(self.location, time, ra, dec, dev)
            if self.dtype is not None:
                angs = angs.to(self.dtype)
        else:
            ra_n, dec_n = utils.tensor2numpy(ra), utils.tensor2numpy(dec)
            fn = self.eq2top_fn if self.eq2top_fn is not None else eq2top
            angs = fn(self.location, time, ra_n, dec_n)
            angs = torch.as_tensor(np.asarray(angs), device=self.device, dtype=self.dtype)
        if store:
            self.conv_cache[key] = angs
        return angs

    def push(self, device):
        dtype = isinstance(device, torch.dtype)
        if dtype:
            self.dtype = device
        else:
            self.device = device
        for k in self.conv_cache:
            self.conv_cache[k] = utils.push(self.conv_cache[k], device)


def JD2LST(jd, longitude):
    """Approximate local (mean) sidereal time [rad] at east longitude [deg]: GMST polynomial
    of the IAU 1982 model.  Stand-in for the astropy call of telescope_model.py:671-690."""
    d = np.asarray(jd, dtype=np.float64) - 2451545.0
    gmst_hours = 18.697374558 + 24.06570982441908 * d
    return np.mod(gmst_hours * 15.0 + longitude, 360.0) * D2R


_WARNED_NO_ASTROPY = False


def _astropy_eq2top(location, time, ra, dec):
    """The reference's own conversion (telescope_model.py:469-502): astropy ICRS -> AltAz."""
    from astropy import units
    from astropy.coordinates import AltAz, EarthLocation, ICRS
    from astropy.time import Time
    if not isinstance(location, EarthLocation):
        alt = location[2] if len(location) > 2 else 0.0
        location = EarthLocation.from_geodetic(location[0], location[1], alt)
    altaz = AltAz(location=location, obstime=Time(time, format='jd'))
    out = ICRS(ra=np.asarray(ra) * units.deg, dec=np.asarray(dec) * units.deg).transform_to(altaz)
    return out.zen.deg, out.az.deg


def _have_astropy():
    try:
        import astropy.coordinates  # noqa: F401
        return True
    except ImportError:
        return False


def _warn_no_astropy():
    global _WARNED_NO_ASTROPY
    if not _WARNED_NO_ASTROPY:
        import warnings
        warnings.warn("bayeslim_b200: astropy is not installed; eq2top computes apparent places "
                      "itself (IAU 1976 precession, IAU 1980 nutation leading terms, annual "
                      "aberration; no UT1-UTC, polar motion or refraction: arcsecond-grade). Inject "
                      "TelescopeModel.conv_cache or eq2top_fn for astropy-grade angles.")
        _WARNED_NO_ASTROPY = True


def eq2top(location, time, ra, dec):
    """(ra, dec) [deg] -> (zen, az) [deg] at `location` (lon, lat[, alt]) and Julian date `time`.
    With astropy installed this is the reference's ICRS -> AltAz transformation
    (telescope_model.py:469-502).  Without it (this image) it falls back, with a warning, to
    eq2top_apparent."""
    try:
        return _astropy_eq2top(location, time, ra, dec)
    except ImportError:
        _warn_no_astropy()
        return eq2top_apparent(location, time, ra, dec)


def _rot(axis, ang):
    """Passive rotation matrix R_axis(ang) (IAU SOFA convention: rotates the frame by +ang)."""
    c, s = np.cos(ang), np.sin(ang)
    if axis == 1:
        return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])
    if axis == 2:
        return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])


def icrs_to_enu(location, jd_utc):
    """(m9, v3): the 3 x 3 rotation from ICRS/J2000 unit vectors to local (East, North, Up) at
    `location` (east lon, lat [deg]) and UTC Julian date, and the observer's velocity / c in the
    same frame (annual aberration).  Precession: IAU 1976 (zeta, z, theta); nutation: the nine
    largest terms of the IAU 1980 series; sidereal time: IAU 1982 GMST (UT1 = UTC) plus the
    equation of the equinoxes.  Frame bias, polar motion, diurnal aberration and refraction are
    left out (each < 0.4 arcsec)."""
    asec = np.pi / 180.0 / 3600.0
    jd_tt = jd_utc + 69.184 / 86400.0
    T = (jd_tt - 2451545.0) / 36525.0
    zeta = (2306.2181 * T + 0.30188 * T ** 2 + 0.017998 * T ** 3) * asec
    z = (2306.2181 * T + 1.09468 * T ** 2 + 0.018203 * T ** 3) * asec
    theta = (2004.3109 * T - 0.42665 * T ** 2 - 0.041833 * T ** 3) * asec
    P = _rot(3, -z) @ _rot(2, theta) @ _rot(3, -zeta)
    d2r = np.pi / 180.0
    Om = (125.04452 - 1934.136261 * T) * d2r          # lunar node
    Ls = (280.4665 + 36000.7698 * T) * d2r            # mean longitude of the Sun
    Lm = (218.3165 + 481267.8813 * T) * d2r           # ... of the Moon
    Ms = (357.52772 + 35999.050340 * T) * d2r         # mean anomaly of the Sun
    Mm = (134.96298 + 477198.867398 * T) * d2r        # ... of the Moon
    dpsi = (-17.1996 * np.sin(Om) - 1.3187 * np.sin(2 * Ls) - 0.2274 * np.sin(2 * Lm)
            + 0.2062 * np.sin(2 * Om) + 0.1426 * np.sin(Ms) + 0.0712 * np.sin(Mm)
            - 0.0517 * np.sin(2 * Ls + Ms) - 0.0386 * np.sin(2 * Lm - Om)
            - 0.0301 * np.sin(2 * Lm + Mm)) * asec
    deps = (9.2025 * np.cos(Om) + 0.5736 * np.cos(2 * Ls) + 0.0977 * np.cos(2 * Lm)
            - 0.0895 * np.cos(2 * Om) + 0.0054 * np.cos(Ms) + 0.0224 * np.cos(2 * Ls + Ms)
            + 0.0200 * np.cos(2 * Lm - Om) + 0.0129 * np.cos(2 * Lm + Mm)) * asec
    eps0 = (84381.448 - 46.8150 * T - 0.00059 * T ** 2 + 0.001813 * T ** 3) * asec
    N = _rot(1, -(eps0 + deps)) @ _rot(3, -dpsi) @ _rot(1, eps0)
    d = jd_utc - 2451545.0
    Tu = d / 36525.0
    gmst = (280.46061837 + 360.98564736629 * d + 0.000387933 * Tu ** 2 - Tu ** 3 / 38710000.0) * d2r
    gast = gmst + dpsi * np.cos(eps0 + deps)
    lon, lat = location[0] * d2r, location[1] * d2r
    # true equator of date -> local meridian frame (x' to the meridian, y' east, z' pole)
    Rl = _rot(3, gast + lon)
    # -> East, North, Up
    H = np.array([[0.0, 1.0, 0.0],
                  [-np.sin(lat), 0.0, np.cos(lat)],
                  [np.cos(lat), 0.0, np.sin(lat)]])
    m9 = H @ Rl @ N @ P
    # annual aberration: Earth's velocity is 90 deg behind the Sun's apparent longitude
    lam = Ls + (1.914602 - 0.004817 * T) * d2r * np.sin(Ms) + 0.019993 * d2r * np.sin(2 * Ms)
    kappa = 20.49552 * asec
    v_date = kappa * np.array([np.sin(lam), -np.cos(lam) * np.cos(eps0), -np.cos(lam) * np.sin(eps0)])
    v3 = P.T @ v_date                                  # mean equator of date -> J2000
    return np.ascontiguousarray(m9, dtype=np.float64), np.ascontiguousarray(v3, dtype=np.float64)


def eq2top_apparent(location, time, ra, dec):
    """Host (numpy) evaluation of icrs_to_enu's transformation: (ra, dec) [deg] -> (zen, az) [deg]."""
    m9, v3 = icrs_to_enu(location, float(time))
    ra = np.asarray(ra, dtype=np.float64) * D2R
    dec = np.asarray(dec, dtype=np.float64) * D2R
    p = np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)]) + v3[:, None]
    p /= np.linalg.norm(p, axis=0)
    e, n, u = m9 @ p
    # atan2 form: arccos(u) loses half the digits near the zenith
    return np.arctan2(np.hypot(e, n), u) / D2R, np.mod(np.arctan2(e, n), 2 * np.pi) / D2R


def eq2top_device(location, time, ra, dec, device):
    """The same on the GPU: (2, Nsrc) float64 tensor [zen, az] (b200rime_eq2top kernel)."""
    import ctypes
    from . import _lib
    m9, v3 = icrs_to_enu(location, float(time))
    ra_t = torch.as_tensor(ra).to(device=device, dtype=torch.float64).contiguous()
    dec_t = torch.as_tensor(dec).to(device=device, dtype=torch.float64).contiguous()
    out = torch.empty(2, ra_t.numel(), dtype=torch.float64, device=device)
    dp = ctypes.POINTER(ctypes.c_double)
    with torch.cuda.device(device):
        rc = _lib.lib.b200rime_eq2top_f64(
            ctypes.c_void_p(ra_t.data_ptr()), ctypes.c_void_p(dec_t.data_ptr()), ra_t.numel(),
            m9.ctypes.data_as(dp), v3.ctypes.data_as(dp), ctypes.c_void_p(out[0].data_ptr()),
            ctypes.c_void_p(out[1].data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    if rc != 0:
        raise _lib.B200RimeError(_lib.lib.b200rime_last_error().decode())
    return out


def eq2top_rigid(location, time, ra, dec):
    """Rigid rotation of (ra, dec) [deg] to (zen, az) [deg] at `location` (lon, lat[, alt]) and
    Julian date `time`.  Not astropy's ICRS->AltAz (telescope_model.py:469-502): no precession,
    nutation, aberration or refraction."""
    lon, lat = location[0], location[1]
    lst = JD2LST(time, lon)
    ra = np.asarray(ra, dtype=np.float64) * D2R
    dec = np.asarray(dec, dtype=np.float64) * D2R
    phi = lat * D2R
    H = lst - ra
    x = -np.cos(dec) * np.sin(H)
    y = np.cos(phi) * np.sin(dec) - np.sin(phi) * np.cos(dec) * np.cos(H)
    z = np.sin(phi) * np.sin(dec) + np.cos(phi) * np.cos(dec) * np.cos(H)
    zen = np.arccos(np.clip(z, -1, 1)) / D2R
    az = np.mod(np.arctan2(x, y), 2 * np.pi) / D2R
    return zen, az


class ArrayModel(utils.Module, utils.AntposDict):
    """Antenna layout, baseline vectors and the interferometric fringe
    (telescope_model.py:140-466)."""

    def __init__(self, antpos, freqs=None, device=None, cache_s=True, cache_depth=None,
                 redtol=1.0, name=None, **kwargs):
        utils.Module.__init__(self, name=name)
        if isinstance(antpos, utils.AntposDict):
            ants, antvecs = antpos.ants, antpos.antvecs
        else:
            ants, antvecs = list(antpos.keys()), list(antpos.values())
        utils.AntposDict.__init__(self, ants, antvecs)
        self.cache_s = cache_s
        self.clear_cache()
        self.redtol = redtol
        self.device = device
        self.cache_depth = cache_depth
        self.set_freqs(freqs)
        (self.reds, self.redvecs, self.bl2red, self.bls, self.redlens, self.redangs,
         self.redtags) = build_reds(self, redtol=redtol, **kwargs)
        if device:
            self.push(device)

    # nn.Module defines __getitem__-free attribute access; route dict-style access to AntposDict
    def __getitem__(self, key):
        if isinstance(key, str):
            return utils.Module.__getitem__(self, key)
        return utils.AntposDict.__getitem__(self, key)

    def __setitem__(self, key, value):
        if isinstance(key, str):
            return utils.Module.__setitem__(self, key, value)
        return utils.AntposDict.__setitem__(self, key, value)

    def get_antpos(self, ant):
        return utils.AntposDict.__getitem__(self, ant)

    def get_blvecs(self, bls):
        """Baseline vectors b = antvecs[j] - antvecs[i] (ENU, metres), differentiable w.r.t.
        antvecs (telescope_model.py:221-239).  One gather + subtract instead of a per-baseline
        Python stack."""
        if isinstance(bls, tuple) or isinstance(bls[0], (int, np.integer)):
            bls = [tuple(bls)]
        i0 = torch.as_tensor([self._ant_idx[int(b[0])] for b in bls], dtype=torch.long,
                             device=self.antvecs.device)
        i1 = torch.as_tensor([self._ant_idx[int(b[1])] for b in bls], dtype=torch.long,
                             device=self.antvecs.device)
        return self.antvecs.index_select(0, i1) - self.antvecs.index_select(0, i0)

    def set_freqs(self, freqs):
        """Frequencies [Hz] of the fringe.  Kept in float64 whatever the session dtype: they are
        geometry (the reference casts to _float(), telescope_model.py:301, which in float32
        sessions jitters a 100-200 MHz grid by +-8 Hz -- 1e-4 rad of phase on a 1 km baseline)."""
        self.freqs = freqs
        if self.freqs is not None:
            self.freqs = torch.as_tensor(self.freqs, device=self.device).to(torch.float64)

    def set_freq_index(self, idx=None):
        self._freq_idx = idx

    def clear_cache(self, depth=None):
        if depth is None:
            self.cache = {}
        else:
            utils.clear_cache_depth(self.cache, depth)

    def gen_fringe(self, blvecs, zen, az, conj=False):
        """Materialised fringe exp(+-2 pi i (b . s) nu / c) of shape (Nbls, Nfreqs, Npix);
        zen, az in degrees (telescope_model.py:310-358).  API compatibility for callers
        outside the RIME -- the RIME hot path generates the fringe inside its kernels."""
        key = utils.arr_hash(zen)
        if not self.cache_s or key not in self.cache:
            _zen, _az = zen * D2R, az * D2R
            s = torch.zeros(3, len(zen), dtype=_float(), device=self.device)
            s[0] = torch.sin(_zen) * torch.sin(_az)
            s[1] = torch.sin(_zen) * torch.cos(_az)
            s[2] = torch.cos(_zen)
            if self.cache_s:
                self.cache[key] = s
        else:
            s = self.cache[key]
        sign = -2j if conj else 2j
        freqs = self.freqs
        if getattr(self, '_freq_idx', None) is not None:
            freqs = freqs[self._freq_idx]
        const = freqs[:, None] * (sign * np.pi / 2.99792458e8)
        return ((blvecs @ s)[:, None, :] * const).exp_()

    def push(self, device):
        dtype = isinstance(device, torch.dtype)
        utils.AntposDict.push(self, device)
        if self.freqs is not None:
            self.freqs = self.freqs.to(device)
        if not dtype:
            self.device = device
        for k in self.cache:
            if isinstance(self.cache[k], torch.Tensor):
                self.cache[k] = self.cache[k].to(device)

    def get_bls(self, uniq_bls=False, keep_autos=True, min_len=None, max_len=None, min_EW=None,
                max_EW=None, min_NS=None, max_NS=None, min_deg=None, max_deg=None, xants=None):
        """All physical baselines (or one per redundant group), optionally selected on the
        baseline vector (telescope_model.py:373-460)."""
        lens = np.asarray(self.redlens, dtype=float)
        angs = np.asarray(self.redangs, dtype=float)
        vecs = np.abs(np.asarray([utils.tensor2numpy(v) for v in self.redvecs]).reshape(len(lens), -1))
        keep = np.ones(len(lens), dtype=bool)
        if not keep_autos:
            autos = np.isclose(lens, 0, atol=self.redtol)
            if autos.any():
                keep[np.where(autos)[0][0]] = False
        if min_len is not None:
            keep &= lens >= min_len
        if max_len is not None:
            keep &= lens <= max_len
        if min_EW is not None:
            keep &= vecs[:, 0] >= min_EW
        if max_EW is not None:
            keep &= vecs[:, 0] <= max_EW
        if min_NS is not None:
            keep &= vecs[:, 1] >= min_NS
        if max_NS is not None:
            keep &= vecs[:, 1] <= max_NS
        if min_deg is not None:
            keep &= angs >= min_deg
        if max_deg is not None:
            keep &= angs <= max_deg
        reds = [list(self.reds[i]) for i in np.where(keep)[0]]
        if uniq_bls:
            reds = [red[:1] for red in reds]
        bls = [bl for red in reds for bl in red]
        if xants is not None:
            bls = [bl for bl in bls if bl[0] not in xants and bl[1] not in xants]
        return bls

    def to_antpos(self):
        return utils.AntposDict(copy.deepcopy(self.ants), self.antvecs.detach().clone())


def build_reds(antpos, bls=None, red_bls=None, redtol=1.0, min_len=None, max_len=None,
               min_EW_len=None, exclude_reds=None, skip_reds=False, norm_vec=False,
               use_blnums=False, use_2d=False, fcluster=False, red_info=None):
    """Sort baselines into redundant groups (telescope_model.py:693-942).

    Same outputs and ordering conventions as the reference: groups are formed greedily in
    baseline order (first vector within `redtol` wins), then sorted by length + angle*redtol/180;
    returns (reds, redvecs, bl2red, bls, redlens, redangs, redtags).  The group search is a
    vectorised distance test per baseline instead of a Python double loop."""
    if red_info is not None:
        return red_info
    if use_blnums or fcluster:
        raise NotImplementedError("use_blnums / fcluster are not part of the RIME path")
    ants = list(antpos.keys())
    if bls is None:
        bls = [(a, a) for a in ants] + list(itertools.combinations(ants, 2))
    idx = {a: i for i, a in enumerate(ants)}
    pos = utils.tensor2numpy(antpos.antvecs if hasattr(antpos, 'antvecs')
                             else torch.as_tensor(np.asarray(list(antpos.values()))))
    pos = np.asarray(pos, dtype=np.float64)
    i0 = np.asarray([idx[b[0]] for b in bls])
    i1 = np.asarray([idx[b[1]] for b in bls])
    vec = pos[i1] - pos[i0]
    if use_2d:
        vec = vec[:, :2]
    lens = np.linalg.norm(vec, axis=1)
    if norm_vec:
        vec = np.zeros_like(vec)
        vec[:, 0] = lens
    keep = np.ones(len(bls), dtype=bool)
    if min_len is not None:
        keep &= lens > min_len
    if max_len is not None:
        keep &= lens < max_len
    if min_EW_len is not None:
        keep &= np.abs(vec[:, 0]) > min_EW_len
    if exclude_reds is not None:
        ex = np.asarray([pos[idx[b[1]]] - pos[idx[b[0]]] for b in exclude_reds])
        if use_2d:
            ex = ex[:, :2]
        for e in ex:
            keep &= ~((np.linalg.norm(vec - e, axis=1) < redtol) |
                      (np.linalg.norm(vec + e, axis=1) < redtol))
    sel = np.where(keep)[0]
    bls = [bls[i] for i in sel]
    vec, lens = vec[sel], lens[sel]

    if skip_reds:
        group = np.arange(len(bls))
        rvec = vec.copy()
    else:
        group = np.zeros(len(bls), dtype=np.int64)
        rvec = np.zeros_like(vec)
        ngroup = 0
        for b in range(len(bls)):
            g = -1
            if ngroup:
                hit = np.where(np.linalg.norm(rvec[:ngroup] - vec[b], axis=1) < redtol)[0]
                if len(hit):
                    g = hit[0]
            if g < 0:
                g = ngroup
                rvec[g] = vec[b]
                ngroup += 1
            group[b] = g
        rvec = rvec[:ngroup]
    rlens = np.linalg.norm(rvec, axis=1)
    rangs = np.degrees(np.arctan2(rvec[:, 1], rvec[:, 0]))
    rangs = np.where(rvec[:, 1] < 0, rangs + 180.0, rangs)
    rangs = np.where(np.abs(rvec[:, 1]) < redtol, 0.0, rangs)
    members = [[] for _ in range(len(rvec))]
    for b, g in enumerate(group):
        members[g].append(bls[b])
    if red_bls is not None:
        order = []
        for rbl in red_bls:
            for i, red in enumerate(members):
                if rbl in red or utils.conjbl(rbl) in red:
                    order.append(i)
                    break
    else:
        order = np.argsort(rlens + rangs * redtol / 180, kind='stable')
    reds = [sorted(members[i]) for i in order]
    redvecs = [torch.as_tensor(rvec[i]) for i in order]
    redlens = [float(rlens[i]) for i in order]
    redangs = [float(rangs[i]) for i in order]
    redtags = ["{:03.0f}_{:03.0f}".format(rlens[i], rangs[i]) for i in order]
    all_bls = [bl for red in reds for bl in red]
    bl2red = {}
    if not skip_reds:
        for i, red in enumerate(reds):
            for bl in red:
                bl2red[bl] = i
    return reds, redvecs, bl2red, all_bls, redlens, redangs, redtags
