"""
Visibility imaging, mirroring the reference's ``imaging.VisMapper`` (bayeslim/imaging.py:12-716)
for the operations that are the adjoint / forward of the RIME fringe sum:

    make_map      m = D A^T w v           (imaging.py:362-478, make_map :717-736)
    compute_Am    v = conj(A) m           (imaging.py:480-540, compute_Am :755-774)
    compute_Pm    P m = D A^T w conj(A) m (imaging.py:542-611, compute_Pm :777-815)

with A[b, f, p] = conj(fringe_b(f, p)) * beam(f, p) (build_A, imaging.py:251-296).  The reference
materialises A (Nbls, Nfreqs, Npix) per time; here A never exists: the sums over baselines run in
the CUDA adjoint kernels (``ops.fringe_adjoint`` = the backward-to-sky kernel of the RIME, the
antenna-factorised one when the baseline set fills its antenna-pair tiles) and the sums over pixels
in the forward kernels (``ops.fringe_sum``), all times of the mapper in one launch.  Only
single-polarisation imaging with an antenna-independent beam, as in the reference.

Not mirrored: the dense (Npix, Npix) PSF matrix (``contract=None``), ``deconvolve_map`` and the
``build_A`` cache (there is no A to cache).  Tensors must live on a CUDA device.
"""
import numpy as np
import torch

from . import ops, telescope_model
from .dataset import MapData


class VisMapper:
    """Dirty maps from visibilities: m = D A^T W y (see the module docstring)."""

    def __init__(self, vd, ra, dec, beam=None, fov=180, dtype=None, cache_A=False, **kwargs):
        """vd: VisData (metadata + visibilities); ra, dec [deg] of the map pixels; beam:
        PixelBeam included in A; fov used when beam is None; dtype: torch.float32 / float64 of
        the kernels (default: that of the visibilities).  kwargs go to ArrayModel."""
        self.vd = vd
        self.telescope = vd.telescope
        kwargs.setdefault('skip_reds', True)
        self.array = telescope_model.ArrayModel(vd.antpos, vd.freqs, device=vd.data.device, **kwargs)
        self.ra, self.dec = ra, dec
        self.Npix = len(ra)
        self.device = vd.data.device
        self.dtype = dtype
        self.beam = beam
        self.fov = beam.fov if beam is not None else fov
        self._freqs = torch.as_tensor(vd.freqs)
        self._times = np.asarray(torch.as_tensor(vd.times).cpu())
        self._bls = list(vd.bls)
        self.set_freq_inds()
        self.set_time_inds()
        self.set_bl_inds()
        self.cache_A = cache_A
        self.clear_cache()
        self.set_normalization()

    # ------------------------------------------------------------------ selections
    def clear_cache(self):
        self.A = {}
        self.D = None
        self._geom = None

    def set_freq_inds(self, freq_inds=None, freqs=None):
        assert not (freqs is not None and freq_inds is not None)
        if freqs is not None:
            f = np.atleast_1d(np.asarray(torch.as_tensor(freqs).cpu(), dtype=np.float64))
            ref = np.asarray(self._freqs.cpu(), dtype=np.float64)
            freq_inds = [int(np.argmin(np.abs(ref - x))) for x in f]
        if freq_inds is None:
            freq_inds = list(range(len(self._freqs)))
        elif isinstance(freq_inds, slice):
            freq_inds = list(range(len(self._freqs)))[freq_inds]
        self.freq_inds = [int(i) for i in np.atleast_1d(np.asarray(freq_inds))]
        self.freqs = self._freqs[self.freq_inds]
        self.Nfreqs = len(self.freq_inds)
        self.clear_cache()

    def set_time_inds(self, time_inds=None, times=None):
        assert not (times is not None and time_inds is not None)
        if times is not None:
            tt = np.atleast_1d(np.asarray(times, dtype=np.float64))
            time_inds = [int(np.where(np.isclose(self._times, t, atol=1e-10, rtol=1e-13))[0][0])
                         for t in tt]
        if time_inds is None:
            time_inds = list(range(len(self._times)))
        elif isinstance(time_inds, slice):
            time_inds = list(range(len(self._times)))[time_inds]
        self.time_inds = [int(i) for i in np.atleast_1d(np.asarray(time_inds))]
        self.times = self._times[self.time_inds]
        self.Ntimes = len(self.time_inds)
        self.clear_cache()

    def set_bl_inds(self, bl_inds=None, bls=None):
        assert not (bls is not None and bl_inds is not None)
        if bls is not None:
            bl_inds = [self._bls.index(tuple(int(a) for a in bl)) for bl in bls]
        if bl_inds is None:
            bl_inds = list(range(len(self._bls)))
        elif isinstance(bl_inds, slice):
            bl_inds = list(range(len(self._bls)))[bl_inds]
        self.bl_inds = [int(i) for i in np.atleast_1d(np.asarray(bl_inds))]
        self.bls = [self._bls[i] for i in self.bl_inds]
        self.Nbls = len(self.bls)
        self.blvecs = self.array.get_blvecs(self.bls)
        self.clear_cache()

    def set_normalization(self, method='A2w', icov=None, clip=1e-8):
        """'w': D = 1 / sum w;  'Aw': D = 1 / sum w |A|;  'A2w': D = 1 / sum w Re(A^2)
        (imaging.py:228-249, :459-464)."""
        assert method in ['w', 'Aw', 'A2w']
        self.method = method
        self.icov = icov
        self.D = None
        self.clip = clip

    # ------------------------------------------------------------------ geometry / operands
    def _rdtype(self):
        if self.dtype is not None:
            return self.dtype
        return ops._real(self.vd.data.dtype)

    def _geometry(self):
        """Packed source axis of all imaged times, per-time FOV cut and beam (imaging.py:270-287)."""
        if self._geom is not None:
            return self._geom
        dev = self.device
        zens, azs, cuts, beams = [], [], [], []
        fidx = torch.as_tensor(self.freq_inds, device=dev)
        for time in self.times:
            za = self.telescope.eq2top(float(time), self.ra, self.dec, store=True)
            zen, az = torch.as_tensor(za[0]).to(dev), torch.as_tensor(za[1]).to(dev)
            if self.beam is not None:
                beam, cut, zen, az = self.beam.gen_beam(zen, az)
                beam = beam.detach()[0, 0, 0].index_select(0, fidx.to(beam.device)).to(dev)
                if not self.beam.powerbeam:
                    beam = beam ** 2
                if not isinstance(cut, torch.Tensor):
                    cut = torch.arange(self.Npix, device=dev)
            else:
                beam = None
                cut = torch.where(zen <= self.fov / 2)[0]
                zen, az = zen[cut], az[cut]
            zens.append(zen.double())
            azs.append(az.double())
            cuts.append(cut.to(dev))
            beams.append(beam)
        geom = ops.Geometry(zens, azs, dev)
        f64 = torch.as_tensor(self.array.freqs).detach().to(dev, torch.float64)[fidx].contiguous()
        blv = self.blvecs.detach().to(dev)
        blmax = float(blv.double().norm(dim=1).max()) if len(blv) else 0.0
        rdt = self._rdtype()
        tiling = None
        if rdt == torch.float32:
            try:
                rows = self.array._ant_idx
                til = ops.AntTiling([rows[int(b[0])] for b in self.bls],
                                    [rows[int(b[1])] for b in self.bls],
                                    len(self.array.antvecs), dev)
                tiling = til if til.usable else None
            except (KeyError, AttributeError):
                tiling = None
        self._geom = dict(geom=geom, cuts=cuts, beams=beams, f64=f64, blv=blv, tiling=tiling,
                          uniform=ops.freqs_uniform(f64, blmax, rdt),
                          uniform2=ops.freqs_uniform(f64, 2 * blmax, rdt))
        return self._geom

    def _adjoint(self, G, twice=False):
        """per-time list of (nplane, Nf, Ns_t): sum_b Re(conj(F_b) G_b); twice: fringe of 2 b."""
        g = self._geometry()
        scale = 2.0 if twice else 1.0
        D = ops.fringe_adjoint(G, g['geom'], g['f64'], self.Nfreqs, blvecs=scale * g['blv'],
                               antvecs=scale * self.array.antvecs.detach().to(self.device),
                               tiling=g['tiling'], conj=False,
                               uniform=g['uniform2'] if twice else g['uniform'])
        return ops.unpack_planes(g['geom'], D, self.Nfreqs)

    def _select(self, data):
        """(Npol, Npol, Nbls, Ntimes, Nfreqs) -> (Nbls, Ntimes, Nfreqs) of the selections."""
        dev = self.device
        x = data[0, 0]
        x = x.index_select(0, torch.as_tensor(self.bl_inds, device=dev))
        x = x.index_select(1, torch.as_tensor(self.time_inds, device=dev))
        return x.index_select(2, torch.as_tensor(self.freq_inds, device=dev))

    def build_v(self, vd=None):
        """Visibilities to image, (Nmaps, Nbls, Ntimes, Nfreqs) (imaging.py:298-326)."""
        vd = self.vd if vd is None else vd
        if isinstance(vd, (list, tuple)):
            return torch.stack([self._select(v.data.to(self.device)) for v in vd])
        if isinstance(vd, torch.Tensor):
            return torch.stack([self._select(v) for v in vd]) if vd.ndim > 5 else self._select(vd)[None]
        return self._select(vd.data.to(self.device))[None]

    def build_w(self):
        """Visibility weights (Nbls, Ntimes, Nfreqs) real: self.icov, else vd.icov, else 1
        (imaging.py:328-360).  Only weights of the data's shape (cov_axis None)."""
        icov = self.icov if self.icov is not None else getattr(self.vd, 'icov', None)
        if icov is None:
            return torch.ones(self.Nbls, self.Ntimes, self.Nfreqs, dtype=self._rdtype(),
                              device=self.device)
        if getattr(self.vd, 'cov_axis', None) is not None and self.icov is None:
            raise NotImplementedError("VisMapper: only diagonal weights of the data's shape")
        return self._select(torch.as_tensor(icov).to(self.device)).real.to(self._rdtype())

    # ------------------------------------------------------------------ operators
    @torch.no_grad()
    def make_map(self, vd=None, return_P=True, contract='diag'):
        """Dirty maps (..., Nfreqs, Npix) summed over the imaged times and normalised, and the
        diagonal of the PSF matrix (imaging.py:362-478)."""
        assert self.method is not None, "First run set_normalization()"
        if return_P and contract != 'diag':
            raise NotImplementedError("VisMapper.make_map: only contract='diag' is provided "
                                      "(the dense PSF matrix is not part of the CUDA path)")
        single = not isinstance(vd if vd is not None else self.vd, (list, tuple, torch.Tensor)) or \
            (isinstance(vd, torch.Tensor) and vd.ndim <= 5)
        g = self._geometry()
        rdt = self._rdtype()
        cdt = ops._cplx(rdt)
        dev = self.device
        v = self.build_v(vd).to(cdt)                                     # (Nmaps, Nbl, Nt, Nf)
        w = self.build_w()                                               # (Nbl, Nt, Nf)
        Nmaps = v.shape[0]
        maps = torch.zeros(Nmaps, self.Nfreqs, self.Npix, dtype=rdt, device=dev)
        dirty = self._adjoint(v * w[None].to(cdt))
        wsum = w.sum(0)                                                  # (Nt, Nf)
        if self.method == 'w':
            Aw = torch.zeros(self.Nfreqs, 1, dtype=rdt, device=dev)
        else:
            Aw = torch.zeros(self.Nfreqs, self.Npix, dtype=rdt, device=dev)
        if self.method == 'A2w':
            # sum_b w Re(A^2) = beam^2 sum_b w Re(conj(F_b)^2), and F_b^2 is the fringe of 2 b
            w2 = self._adjoint(w[None].to(cdt), twice=True)
        P = torch.zeros(self.Nfreqs, self.Npix, dtype=rdt, device=dev) if return_P else None
        for i in range(self.Ntimes):
            cut, beam = g['cuts'][i], g['beams'][i]
            b1 = beam.to(rdt) if beam is not None else torch.ones(1, 1, dtype=rdt, device=dev)
            maps[:, :, cut] += dirty[i] * b1[None]
            if return_P:
                P[:, cut] += (b1.abs() ** 2 * wsum[i][:, None]).expand(self.Nfreqs, len(cut))
            if self.method == 'w':
                Aw += wsum[i][:, None]
            elif self.method == 'Aw':
                Aw[:, cut] += (b1.abs() * wsum[i][:, None]).expand(self.Nfreqs, len(cut))
            else:
                Aw[:, cut] += w2[i][0] * b1 ** 2
        self.D = 1 / Aw.clip(self.clip)
        maps *= self.D
        if return_P:
            P *= self.D
        return (maps[0] if single else maps), P

    @torch.no_grad()
    def compute_Am(self, maps):
        """Visibilities conj(A) @ maps of shape ([Nmaps,] Nbls, Ntimes, Nfreqs): the RIME forward
        of the maps through the mapper's beam (imaging.py:480-540)."""
        m2t = lambda m: m.data if isinstance(m, MapData) else m
        if isinstance(maps, (list, tuple)):
            maps = torch.stack([m2t(m) for m in maps])
        maps = m2t(maps)
        single = maps.ndim <= 2
        maps = maps.reshape((-1,) + tuple(maps.shape[-2:])).to(self.device)
        g = self._geometry()
        rdt = self._rdtype()
        planes = []
        for i in range(self.Ntimes):
            cut, beam = g['cuts'][i], g['beams'][i]
            x = maps[:, :, cut].to(rdt)
            planes.append((x * beam.to(rdt)[None]) if beam is not None else x)
        A = ops.pack_planes(g['geom'], [p.contiguous() for p in planes])
        til = g['tiling']
        if til is not None:
            V = ops.fringe_sum_ant(A, self.array.antvecs.detach().to(self.device), til, g['geom'],
                                   g['f64'], self.Nfreqs, conj=False)
        else:
            V = ops.fringe_sum(A, g['blv'], g['geom'], g['f64'], self.Nfreqs, conj=False,
                               uniform=g['uniform'])
        return V[0] if single else V

    @torch.no_grad()
    def compute_Pm(self, maps, D=None):
        """P @ maps = D A^T w (conj(A) maps), (Nmaps, Nfreqs, Npix) (imaging.py:542-611)."""
        v = self.compute_Am(maps)
        v = v[None] if v.ndim == 3 else v
        g = self._geometry()
        rdt = self._rdtype()
        w = self.build_w()
        dirty = self._adjoint(v * w[None].to(v.dtype))
        out = torch.zeros(v.shape[0], self.Nfreqs, self.Npix, dtype=rdt, device=self.device)
        for i in range(self.Ntimes):
            cut, beam = g['cuts'][i], g['beams'][i]
            out[:, :, cut] += dirty[i] * (beam.to(rdt)[None] if beam is not None else 1.0)
        if D is not None:
            out *= D
        return out

