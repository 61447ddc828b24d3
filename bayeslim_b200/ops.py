"""
torch.autograd bindings of the C-ABI CUDA library (include/b200rime.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape.  All
arithmetic of the hot path runs in libb200rime.so; every Function below raises if it is
handed a non-CUDA tensor (there is no CPU path).

Objects
-------
Geometry      packed source axis of one time group (per-time FOV-cut sources padded to 128,
              float64 unit vectors, work-unit tables)
pack_planes   row-major (nplane, Nf, Ns_t) perceived-sky planes  -> tiled layout A
build_interp  fused PixInterp.interp x cut_sky_fov x beam*sky      -> A   (reference
              utils.py:833-841, beam_model.py:1696, :341)
build_airy    fused airy_disk x cut_sky_fov x beam*sky             -> A   (beam_model.py:1464-1480)
fringe_sum    A, blvecs -> V (nplane, Nbl, Nt, Nf): gen_fringe + multiply + sum
              (telescope_model.py:351-356, rime_model.py:426-429) and its adjoints
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

D2R = math.pi / 180.0

# upper bound on the per-launch partial-visibility workspace [bytes]
VPART_BUDGET = 6 << 30
# longest run of sources accumulated in float32 registers before a float64 reduction
UNIT_MAX_SRC = 8192
UNIT_MIN_SRC = 256
# grid.z limit of a launch (units per launch)
MAX_GRID_UNITS = 32768


def _sfx(dtype):
    if dtype in (torch.float32, torch.complex64):
        return "f32"
    if dtype in (torch.float64, torch.complex128):
        return "f64"
    raise TypeError("b200rime kernels exist for float32/complex64 and float64/complex128 only")


def _real(dtype):
    return torch.float32 if _sfx(dtype) == "f32" else torch.float64


def _cplx(dtype):
    return torch.complex64 if _sfx(dtype) == "f32" else torch.complex128


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bayeslim_b200: tensors must live on a CUDA device "
                               "(the hot path has no CPU implementation)")


def nchunks(nfreq, dtype):
    kc = _lib.KC[_sfx(dtype)]
    return (nfreq + kc - 1) // kc


_launches = 0


def launch_count():
    """Number of libb200rime kernel launches issued by this process (bench.py reports it)."""
    return _launches


def _call(name, sfx, *args):
    """Launch b200rime_<name>_<sfx>.  Tensor arguments are passed as raw device pointers
    (None -> NULL); the trailing stream argument is appended here."""
    global _launches
    _launches += 1
    conv = [ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor)
            else (ctypes.c_void_p(0) if a is None else a) for a in args]
    _lib.call(name, sfx, *conv, _stream())


def freqs_uniform(freqs, blmax, dtype):
    """Host check that `freqs` (float64 tensor) is equally spaced inside every chunk to better
    than 2e-8 cycles of fringe phase on the longest baseline, i.e. that the rotation
    recurrence is exact to rounding.  Otherwise the kernels evaluate every channel directly."""
    f = freqs.detach().double().cpu().numpy()
    kc = _lib.KC[_sfx(dtype)]
    worst = 0.0
    for k0 in range(0, len(f), kc):
        c = f[k0:k0 + kc]
        if len(c) < 3:
            continue
        df = (c[-1] - c[0]) / (len(c) - 1)
        mid = c[kc // 2] if kc // 2 < len(c) else c[0] + (kc // 2) * df
        fit = mid + (np.arange(len(c)) - kc // 2) * df
        worst = max(worst, float(np.abs(c - fit).max()))
    return worst * max(float(blmax), 1.0) / 2.99792458e8 < 2e-8


class Geometry:
    """Packed source axis for a group of times (device resident, built once and cached).

    zen_list/az_list: per-time float64 tensors [deg] of the sources INSIDE the FOV.
    """

    def __init__(self, zen_list, az_list, device):
        self.device = torch.device(device)
        pad = _lib.SRC_PAD
        self.ns = [int(len(z)) for z in zen_list]
        self.ns_pad = [((n + pad - 1) // pad) * pad for n in self.ns]
        self.toff = [0]
        for n in self.ns_pad:
            self.toff.append(self.toff[-1] + n)
        self.S = self.toff[-1]
        self.nt = len(self.ns)
        shat = torch.zeros(max(self.S, 1), 4, dtype=torch.float64, device=self.device)
        for t, (zen, az) in enumerate(zip(zen_list, az_list)):
            if self.ns[t] == 0:
                continue
            z = zen.to(self.device, torch.float64) * D2R
            a = az.to(self.device, torch.float64) * D2R
            sl = slice(self.toff[t], self.toff[t] + self.ns[t])
            sz = torch.sin(z)
            shat[sl, 0] = sz * torch.sin(a)   # east      (telescope_model.py:341)
            shat[sl, 1] = sz * torch.cos(a)   # north     (:342)
            shat[sl, 2] = torch.cos(z)        # up        (:343)
        self.shat = shat
        tt = np.zeros(max(self.S // pad, 1), dtype=np.int32)
        for t in range(self.nt):
            tt[self.toff[t] // pad:self.toff[t + 1] // pad] = t
        self.tile_time = torch.as_tensor(tt, device=self.device)
        self._unit_cache = {}

    def units(self, nbl, nchunk, sm_count):
        """Work units (time, s_begin, s_end) for the baseline-owned kernels.

        A unit is at most UNIT_MAX_SRC sources (bounds the float32 accumulation length; unit
        partials are then summed in float64) and small enough that the grid fills the GPU for
        at least ~8 waves when the baseline x chunk grid alone does not."""
        key = (nbl, nchunk, sm_count)
        if key in self._unit_cache:
            return self._unit_cache[key]
        tile = _lib.SRC_TILE
        base = ((nbl + 127) // 128) * nchunk
        target = 24 * sm_count
        want = max(self.nt, -(-target // max(base, 1)))
        per = max(self.S, 1) / want
        unit = int(min(max(math.ceil(per / tile) * tile, UNIT_MIN_SRC), UNIT_MAX_SRC))
        rows, ubeg = [], [0]
        for t in range(self.nt):
            s0, s1 = self.toff[t], self.toff[t + 1]
            while s0 < s1:
                e = min(s0 + unit, s1)
                rows.append((t, s0, e, 0))
                s0 = e
            ubeg.append(len(rows))
        units = torch.as_tensor(np.asarray(rows, dtype=np.int32).reshape(-1, 4), device=self.device)
        out = (units, ubeg)
        self._unit_cache[key] = out
        return out


_SM_COUNT = {}


def sm_count(device):
    idx = torch.device(device).index or 0
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


# ----------------------------------------------------------------------------- pack
class _Pack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geom, *planes):
        _need_cuda(*planes)
        x0 = planes[0]
        dtype, sfx = x0.dtype, _sfx(x0.dtype)
        nplane, nfreq = x0.shape[0], x0.shape[1]
        kc = _lib.KC[sfx]
        nchunk = nchunks(nfreq, dtype)
        A = torch.empty(nplane, nchunk, max(geom.S, 1), kc, dtype=dtype, device=x0.device)
        for t, X in enumerate(planes):
            X = X.contiguous()
            assert X.shape == (nplane, nfreq, geom.ns[t])
            for p in range(nplane):
                _call("pack", sfx, X[p], geom.ns[t], nfreq, geom.ns[t], geom.ns_pad[t],
                      geom.toff[t], geom.S, A[p])
        ctx.geom, ctx.nfreq, ctx.nplane = geom, nfreq, nplane
        return A

    @staticmethod
    def backward(ctx, dA):
        geom, nfreq, nplane = ctx.geom, ctx.nfreq, ctx.nplane
        dA = dA.contiguous()
        sfx = _sfx(dA.dtype)
        grads = []
        for t in range(geom.nt):
            X = torch.empty(nplane, nfreq, geom.ns[t], dtype=dA.dtype, device=dA.device)
            for p in range(nplane):
                _call("unpack", sfx, dA[p], geom.ns[t], nfreq, geom.ns[t], geom.toff[t],
                      geom.S, X[p])
            grads.append(X)
        return (None,) + tuple(grads)


def pack_planes(geom, planes):
    """planes: list over times of real (nplane, Nf, Ns_t) tensors -> A (nplane, nchunk, S, KC)."""
    return _Pack.apply(geom, *planes)


# ----------------------------------------------------------------------------- fused builders
class InterpRecord:
    """Per-time tables of the interpolated-beam builder (device): FOV cut indices into the sky,
    neighbour indices/weights into the beam map, and their CSR transpose for the adjoint."""

    def __init__(self, cut, inds, wgts, npix_beam, dtype, device):
        self.cut = cut.to(device=device, dtype=torch.int32).contiguous()
        self.inds = inds.to(device=device, dtype=torch.int32).contiguous()
        self.wgts = wgts.to(device=device, dtype=dtype).contiguous()
        self.nnn = int(inds.shape[1]) if inds.ndim == 2 else 1
        self.ns = int(len(cut))
        self.npix_beam = npix_beam
        self._csr = None

    def csr(self):
        """CSR transpose (beam pixel -> [(source, weight)]), built on first backward."""
        if self._csr is None:
            flat = self.inds.reshape(-1).long()
            order = torch.argsort(flat, stable=True)
            counts = torch.bincount(flat, minlength=self.npix_beam)
            rowptr = torch.zeros(self.npix_beam + 1, dtype=torch.int32, device=flat.device)
            rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            col = (order // self.nnn).to(torch.int32).contiguous()
            val = self.wgts.reshape(-1)[order].contiguous()
            self._csr = (rowptr, col, val)
        return self._csr


class _BuildInterp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sky, bmap, geom, recs):
        _need_cuda(sky, bmap)
        dtype, sfx = sky.dtype, _sfx(sky.dtype)
        sky = sky.contiguous()
        bmap = bmap.to(dtype).contiguous()
        nfreq = sky.shape[0]
        assert bmap.shape[0] == nfreq
        kc = _lib.KC[sfx]
        A = torch.empty(1, nchunks(nfreq, dtype), max(geom.S, 1), kc, dtype=dtype, device=sky.device)
        for t, r in enumerate(recs):
            _call("build_interp", sfx, bmap, bmap.shape[1], r.inds, r.wgts, r.nnn,
                  sky, sky.shape[1], r.cut, nfreq, r.ns, geom.ns_pad[t], geom.toff[t],
                  geom.S, A[0])
        ctx.save_for_backward(sky, bmap)
        ctx.geom, ctx.recs = geom, recs
        return A

    @staticmethod
    def backward(ctx, dA):
        sky, bmap = ctx.saved_tensors
        geom, recs = ctx.geom, ctx.recs
        dA = dA.contiguous()
        sfx = _sfx(sky.dtype)
        nfreq = sky.shape[0]
        need_sky, need_beam = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dsky = torch.zeros_like(sky) if need_sky else None
        dbmap = torch.zeros_like(bmap) if need_beam else None
        nsmax = max([r.ns for r in recs] + [1])
        dBI = torch.empty(nfreq, nsmax, dtype=sky.dtype, device=sky.device) if need_beam else None
        for t, r in enumerate(recs):
            if r.ns == 0:
                continue
            _call("build_interp_bwd", sfx, dA[0], bmap, bmap.shape[1], r.inds, r.wgts,
                  r.nnn, sky, sky.shape[1], r.cut, nfreq, r.ns, geom.toff[t], geom.S,
                  dsky, dBI, nsmax)
            if need_beam:
                rowptr, col, val = r.csr()
                _call("interp_transpose", sfx, dBI, nsmax, rowptr, col, val,
                      r.npix_beam, nfreq, dbmap, bmap.shape[1])
        return dsky, dbmap, None, None


def build_interp(sky, bmap, geom, recs):
    """sky (Nf, Npix), bmap (Nf, Npb) -> A (1, nchunk, S, KC)."""
    return _BuildInterp.apply(sky, bmap, geom, recs)


class AiryRecord:
    """Per-time tables of the Airy builder: cut, sin(min(zen, 90deg)) and sin(az)^2."""

    def __init__(self, cut, zen_deg, az_deg, dtype, device):
        self.cut = cut.to(device=device, dtype=torch.int32).contiguous()
        zen = torch.clamp(zen_deg.to(device, torch.float64) * D2R, max=math.pi / 2)
        az = az_deg.to(device, torch.float64) * D2R
        self.sinzen = torch.sin(zen).to(dtype).contiguous()
        self.sin2az = (torch.abs(torch.sin(az)) ** 2).to(dtype).contiguous()
        self.ns = int(len(cut))


class _BuildAiry(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sky, diam, geom, recs, freqs64, freq_ratio, square, full_grad):
        _need_cuda(sky)
        dtype, sfx = sky.dtype, _sfx(sky.dtype)
        sky = sky.contiguous()
        nfreq = sky.shape[0]
        d = diam.detach().double().cpu().reshape(-1)
        Dew = float(d[0])
        asym = d.numel() > 1
        Dns = float(d[1]) if asym else Dew
        kc = _lib.KC[sfx]
        A = torch.empty(1, nchunks(nfreq, dtype), max(geom.S, 1), kc, dtype=dtype, device=sky.device)
        for t, r in enumerate(recs):
            _call("build_airy", sfx, Dew, Dns, float(freq_ratio), int(square), r.sinzen,
                  r.sin2az if asym else None, freqs64, sky, sky.shape[1], r.cut,
                  nfreq, r.ns, geom.ns_pad[t], geom.toff[t], geom.S, A[0], None, 0)
        ctx.save_for_backward(sky, freqs64)
        ctx.meta = (geom, recs, Dew, Dns, asym, float(freq_ratio), int(square), int(full_grad),
                    diam.shape, diam.dtype, diam.device)
        return A

    @staticmethod
    def backward(ctx, dA):
        sky, freqs64 = ctx.saved_tensors
        geom, recs, Dew, Dns, asym, ratio, square, full_grad, dshape, ddtype, ddev = ctx.meta
        dA = dA.contiguous()
        sfx = _sfx(sky.dtype)
        nfreq = sky.shape[0]
        need_sky, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dsky = torch.zeros_like(sky) if need_sky else None
        gD = torch.zeros(2, dtype=torch.float64, device=sky.device)
        for t, r in enumerate(recs):
            if r.ns == 0:
                continue
            dD = None
            if need_d:
                nb = _lib.lib.b200rime_airy_bwd_blocks(nfreq, r.ns)
                dD = torch.zeros(nb, 2, dtype=torch.float64, device=sky.device)
            _call("build_airy_bwd", sfx, dA[0], Dew, Dns, ratio, square, full_grad, r.sinzen,
                  r.sin2az if asym else None, freqs64, sky, sky.shape[1], r.cut,
                  nfreq, r.ns, geom.toff[t], geom.S, dsky, dD)
            if need_d:
                gD = gD + dD.sum(0)
        gdiam = None
        if need_d:
            g = gD if asym else gD[:1]
            gdiam = g.to(device=ddev, dtype=ddtype).reshape(dshape)
        return dsky, gdiam, None, None, None, None, None, None


def build_airy(sky, diam, geom, recs, freqs64, freq_ratio=1.0, square=True, full_grad=False):
    """sky (Nf, Npix), diam tensor with 1 (D) or 2 (Dew, Dns) elements -> A (1, nchunk, S, KC)."""
    return _BuildAiry.apply(sky, diam, geom, recs, freqs64, freq_ratio, square, full_grad)


# ----------------------------------------------------------------------------- fringe sum
def _blv4(blvecs, device):
    b = torch.zeros(blvecs.shape[0], 4, dtype=torch.float64, device=device)
    b[:, :3] = blvecs.detach().to(device=device, dtype=torch.float64)
    return b


class _FringeSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, blvecs, geom, freqs64, nfreq, conj, uniform):
        _need_cuda(A, freqs64)
        dtype, sfx = A.dtype, _sfx(A.dtype)
        A = A.contiguous()
        dev = A.device
        nplane, nchunk = A.shape[0], A.shape[1]
        kc = _lib.KC[sfx]
        nfp = nchunk * kc
        nbl, nt = blvecs.shape[0], geom.nt
        blv = _blv4(blvecs, dev)
        V = torch.zeros(nplane, nbl, nt, nfreq, dtype=_cplx(dtype), device=dev)
        if nbl > 0 and nt > 0 and geom.S > 0:
            units, ubeg = geom.units(nbl, nchunk, sm_count(dev))
            esize = 8 if sfx == "f32" else 16
            per_unit = nbl * nfp * esize
            max_units = min(max(1, VPART_BUDGET // per_unit), MAX_GRID_UNITS)
            # sub-batches of whole times whose unit partials fit the workspace budget
            t0 = 0
            batches = []
            while t0 < nt:
                t1 = t0 + 1
                while t1 < nt and ubeg[t1 + 1] - ubeg[t0] <= max_units:
                    t1 += 1
                batches.append((t0, t1))
                t0 = t1
            nu_max = max(ubeg[b] - ubeg[a] for a, b in batches)
            vpart = torch.empty(nu_max, nbl, nfp, 2, dtype=dtype, device=dev)
            Vr = torch.view_as_real(V)
            for (ta, tb) in batches:
                u0, u1 = ubeg[ta], ubeg[tb]
                ub = torch.as_tensor(np.asarray(ubeg[ta:tb + 1], dtype=np.int32) - u0, device=dev)
                for p in range(nplane):
                    _call("fringe_sum_fwd", sfx, A[p], geom.shat, blv, freqs64,
                          units[u0:], u1 - u0, nbl, nfreq, geom.S, int(conj), int(uniform),
                          vpart)
                    _call("reduce_units", sfx, vpart, ub, tb - ta, nbl, nfreq,
                          Vr[p, :, ta:], nt * nfreq, nfreq, 1, 1.0, 0.0, 0)
        need_bl = ctx.needs_input_grad[1]
        ctx.save_for_backward(A if need_bl else None, blv, freqs64)
        ctx.meta = (geom, nfreq, int(conj), int(uniform), blvecs.dtype, blvecs.device, A.shape)
        return V

    @staticmethod
    def backward(ctx, G):
        A, blv, freqs64 = ctx.saved_tensors
        geom, nfreq, conj, uniform, bdtype, bdev, ashape = ctx.meta
        need_A, need_bl = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dev = G.device
        rdtype = _real(G.dtype)
        sfx = _sfx(G.dtype)
        nplane, nchunk, _, kc = ashape
        nfp = nchunk * kc
        nbl, nt = blv.shape[0], geom.nt
        dA = torch.zeros(ashape, dtype=rdtype, device=dev) if need_A else None
        dbl = torch.zeros(nbl, 3, dtype=torch.float64, device=dev) if need_bl else None
        if nbl > 0 and nt > 0 and geom.S > 0:
            Gr = torch.view_as_real(G.contiguous())
            # cotangent in kernel layout [t][chunk][baseline][KC] (zero padded in frequency): a
            # tile of baselines is one contiguous block for the TMA engine
            Gp = torch.zeros(nt, nchunk, nbl, kc, 2, dtype=rdtype, device=dev)
            Gv = Gp.permute(2, 0, 1, 3, 4)                     # (nbl, nt, nchunk, kc, 2) view
            pad = nfp - nfreq
            for p in range(nplane):
                src = Gr[p] if pad == 0 else torch.nn.functional.pad(Gr[p], (0, 0, 0, pad))
                Gv.copy_(src.reshape(nbl, nt, nchunk, kc, 2))
                if need_A:
                    _call("fringe_sum_bwd_sky", sfx, Gp, geom.shat, blv, freqs64,
                          geom.tile_time, nbl, nt, nfreq, geom.S, conj, uniform, dA[p])
                if need_bl:
                    units, ubeg = geom.units(nbl, nchunk, sm_count(dev))
                    per_unit = nchunk * nbl * 32
                    step = min(max(1, (VPART_BUDGET // 4) // per_unit), MAX_GRID_UNITS)
                    for u0 in range(0, units.shape[0], step):
                        nun = min(step, units.shape[0] - u0)
                        part = torch.empty(nun, nchunk, nbl, 4, dtype=torch.float64, device=dev)
                        _call("fringe_sum_bwd_bl", sfx, Gp, A[p], geom.shat, blv,
                              freqs64, units[u0:], nun, nbl, nt, nfreq, geom.S, conj, uniform,
                              part)
                        dbl = dbl + part.sum(dim=(0, 1))[:, :3]
        gbl = dbl.to(device=bdev, dtype=bdtype) if need_bl else None
        return dA, gbl, None, None, None, None, None


def fringe_sum(A, blvecs, geom, freqs64, nfreq, conj=False, uniform=True):
    """A (nplane, nchunk, S, KC) real, blvecs (Nbl, 3) -> V (nplane, Nbl, Nt, Nf) complex:
    V[p, b, t, f] = sum_s A[p, f, s] exp(+-2 pi i (b . shat_s) nu_f / c)."""
    return _FringeSum.apply(A, blvecs, geom, freqs64, nfreq, conj, uniform)
