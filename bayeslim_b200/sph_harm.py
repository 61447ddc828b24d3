"""
Spherical-harmonic forward model a_lm -> f(theta, phi) (reference
sph_harm.py:1244-1745 ``AlmModel``) on the CUDA library: the product with the Ylm matrix is the
tensor-core complex GEMM ``b200rime_cgemm_f32`` (float64 sessions: ``b200rime_cgemm_f64``), its
adjoint the same kernel with the conjugate-transposed operand.  There is no torch / CPU route
for the product.

In scope: integer-degree harmonics on the sphere (``gen_sph2pix(method='sphere')``, the case the
beam models use), user-supplied Ylm matrices of any kind, separable (Theta, Phi) grids, the Ylm
cache and multigrid forward.  Out of scope (host-side precompute of the reference, minutes of
mpmath per matrix, off the hot path): non-integer-degree 'stripe' / 'cap' harmonics, Ylm file
IO, least-squares inverses, the spherical Fourier-Bessel classes.
"""
import numpy as np
import torch

from . import ops, utils


def gen_lm(lmax, real_field=True):
    """(2, Ncoeff) array of (l, m), m-major, healpy Alm.getlm order (sph_harm.py:14-39)."""
    m0 = 0 if real_field else -lmax
    return np.array([[l, m] for m in range(m0, lmax + 1) for l in range(abs(m), lmax + 1)]).T


def gen_sph2pix(theta, phi, l, m, separable=False, method='sphere', theta_crit=None, device=None,
                real=False, m_phasor=False, renorm=False, **kwargs):
    """Ylm matrix (Ncoeff, Npix) -- or (Theta, Phi) if separable -- for theta, phi in RADIANS,
    orthonormal harmonics sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_lm(cos theta) exp(i m phi)
    (sph_harm.py:255-475).  Returns (Ylm, norm, alm_mult) like the reference; alm_mult doubles
    the m > 0 modes when the negative orders were truncated."""
    if method != 'sphere' or renorm:
        raise NotImplementedError("only integer-degree full-sphere harmonics are generated here; "
                                  "pass a precomputed Ylm to setup_Ylm / set_Ylm for the others")
    from scipy import special
    l = np.atleast_1d(np.asarray(l))
    m = np.atleast_1d(np.asarray(m))
    if not (np.allclose(l, np.round(l)) and np.allclose(m, np.round(m))):
        raise NotImplementedError("non-integer degree harmonics are not generated here")
    li, mi = np.round(l).astype(int)[:, None], np.round(m).astype(int)[:, None]
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    if hasattr(special, 'sph_harm_y'):
        H = special.sph_harm_y(li, mi, theta[None, :], 0.0).real
    else:
        H = special.sph_harm(mi, li, 0.0, theta[None, :]).real
    Phi = np.exp(1j * mi * phi[None, :])
    if m_phasor:
        Phi = Phi * np.exp(1j * phi[None, :])
    dtype = utils._float() if real else utils._cfloat()
    cast = (lambda x: torch.as_tensor(np.real(x) if real else x, dtype=dtype, device=device))
    Y = (cast(H), cast(Phi)) if separable else cast(H * Phi)
    alm_mult = torch.ones(len(l), dtype=utils._float())
    if not np.any(m < 0) and not real:
        alm_mult[torch.as_tensor(m.ravel() > 0)] *= 2
    if m_phasor and not real:
        alm_mult[torch.as_tensor(np.isclose(m.ravel(), 0))] *= 2
    return Y, torch.ones(len(l)), alm_mult


class AlmModel:
    """f(theta, phi) = sum_lm Y_lm(theta, phi) a_lm: params (..., Ncoeff) -> (..., Npix)
    (sph_harm.py:1244-1745).  Ylm matrices are cached by the hash of their theta array; each
    matrix is packed once for the CUDA product (ops.AlmPlan) the first time it is used."""

    def __init__(self, l, m, default_kw=None, real_output=False, LM=None):
        self.l, self.m = l, m
        self.device = None
        self.default_kw = {} if default_kw is None else default_kw
        self.real_output = real_output
        self.LM = LM
        self.clear_Ylm_cache()
        self.clear_multigrid()

    def __call__(self, params, **kwargs):
        return self.forward_alm(params, **kwargs)

    # ------------------------------------------------------------------ the product
    def _plan(self, Y):
        key = (Y.data_ptr(), tuple(Y.shape), Y.dtype, Y._version)
        if key not in self._plans:
            if len(self._plans) > 16:
                self._plans.clear()
            self._plans[key] = (ops.AlmPlan(Y), Y)       # keep Y alive: the key is its address
        return self._plans[key][0]

    def forward_alm(self, params, Ylm=None, alm_mult=None, ignoreLM=False):
        """a_lm -> map (sph_harm.py:1289-1373)."""
        if self.LM is not None and not ignoreLM:
            params = self.LM(params)
        if Ylm is None and self.multigrid is not None:
            outs = []
            for h in self.multigrid:
                c = self.Ylm_cache[h]
                outs.append(self.forward_alm(params, Ylm=c['Ylm'], alm_mult=c['alm_mult'],
                                             ignoreLM=True))
            out = torch.cat(outs, dim=-1)
            if self._multigrid_idx is not None:
                out = torch.index_select(out, -1, self._multigrid_idx)
            return out
        if Ylm is None:
            Ylm, alm_mult, separable = self.Ylm, self.alm_mult, self.separable
        else:
            separable = isinstance(Ylm, (list, tuple))
        ycomplex = torch.is_complex(Ylm[1] if separable else Ylm)
        if ycomplex and not torch.is_complex(params):
            params = utils.viewcomp(params)
        if alm_mult is not None:
            params = params * alm_mult.to(params.device)
        if separable:
            # out[.., t, p] = sum_c (a_c Theta[c, t]) Phi[c, p]: the scaling by Theta is
            # elementwise, the sum over modes is the CUDA product with Phi
            Theta, Phi = Ylm
            x = params.unsqueeze(-2) * Theta.T.to(params.device)            # (..., Nt, Ncoeff)
            out = ops.alm_forward(x, self._plan(Phi), real_out=self.real_output)
            return out.reshape(out.shape[:-2] + (Theta.shape[1] * Phi.shape[1],))
        return ops.alm_forward(params, self._plan(Ylm), real_out=self.real_output)

    # ------------------------------------------------------------------ Ylm bookkeeping
    @staticmethod
    def setup_angs(theta, phi, separable):
        if separable:
            phi_arr, theta_arr = np.meshgrid(np.asarray(phi), np.asarray(theta), copy=False)
            return theta_arr.ravel(), phi_arr.ravel()
        return theta, phi

    def setup_Ylm(self, theta, phi, Ylm=None, alm_mult=None, separable=False, generate=False,
                  cache=True, h=None, **kwargs):
        """Attach (and cache) the transform matrices of a set of angles [deg]; generate them
        when asked (sph_harm.py:1408-1494)."""
        self.theta, self.phi = theta, phi
        if separable:
            self.theta_grid, self.phi_grid = theta, phi
            self.theta, self.phi = self.setup_angs(theta, phi, separable)
        if Ylm is None and generate:
            kw = dict(self.default_kw)
            kw.update(kwargs)
            th, ph = (self.theta_grid, self.phi_grid) if separable else (self.theta, self.phi)
            Ylm, _, alm_mult = gen_sph2pix(_np(th) * utils.D2R, _np(ph) * utils.D2R, self.l, self.m,
                                           separable=separable, device=self.device, **kw)
        self.Ylm, self.alm_mult, self.separable = Ylm, alm_mult, separable
        if separable and Ylm is not None:
            assert isinstance(Ylm, (tuple, list))
        if cache:
            angs = (self.theta_grid, self.phi_grid) if separable else (theta, phi)
            self.set_Ylm(Ylm, angs, alm_mult=alm_mult, h=h)

    def get_Ylm(self, theta, phi, separable=False, h=None):
        """Cached (Ylm, alm_mult) of these angles, generating them on a miss
        (sph_harm.py:1496-1547; a miss in the reference attaches Ylm = None)."""
        h = h if h is not None else utils.arr_hash(theta)
        if h in self.Ylm_cache:
            c = self.Ylm_cache[h]
            Ylm, alm_mult = c['Ylm'], c['alm_mult']
            theta, phi = c['angs']
            if separable:
                self.theta_grid, self.phi_grid = theta, phi
                theta, phi = self.setup_angs(theta, phi, separable)
        else:
            self.setup_Ylm(theta, phi, cache=True, h=h, separable=separable, generate=True)
            Ylm, alm_mult = self.Ylm, self.alm_mult
            theta, phi = self.theta, self.phi
        self.Ylm, self.alm_mult, self.separable = Ylm, alm_mult, separable
        self.theta, self.phi = theta, phi
        return Ylm, alm_mult

    def set_Ylm(self, Ylm, angs, alm_mult=None, h=None):
        h = h if h is not None else utils.arr_hash(angs[0])
        self.Ylm_cache[h] = dict(Ylm=Ylm, angs=angs, separable=isinstance(Ylm, (tuple, list)),
                                 alm_mult=alm_mult)
        return h

    def clear_Ylm_cache(self):
        self.Ylm_cache = {}
        self._plans = {}

    def setup_multigrid_forward(self, thetas, phis, Ylms, alm_mults, idx=None):
        self.multigrid = [self.set_Ylm(Y, (th, ph), alm_mult=a)
                          for th, ph, Y, a in zip(thetas, phis, Ylms, alm_mults)]
        self._multigrid_idx = idx

    def clear_multigrid(self):
        self.multigrid = None
        self._multigrid_idx = None

    def push(self, device):
        isdtype = isinstance(device, torch.dtype)

        def mv(Y):
            if isinstance(Y, (tuple, list)):
                return tuple(utils.push(y, device) for y in Y)
            return utils.push(Y, device)
        if getattr(self, 'Ylm', None) is not None:
            self.Ylm = mv(self.Ylm)
        if getattr(self, 'alm_mult', None) is not None:
            self.alm_mult = utils.push(self.alm_mult, device)
        if self._multigrid_idx is not None and not isdtype:
            self._multigrid_idx = self._multigrid_idx.to(device)
        for c in self.Ylm_cache.values():
            c['Ylm'] = mv(c['Ylm'])
            if c['alm_mult'] is not None:
                c['alm_mult'] = utils.push(c['alm_mult'], device)
        self._plans = {}
        if self.LM is not None:
            self.LM.push(device)
        if not isdtype:
            self.device = device


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
