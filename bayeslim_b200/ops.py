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


def _stream(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


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
    dev = None
    for a in args:
        for t in (a if isinstance(a, (list, tuple)) else (a,)):
            if isinstance(t, torch.Tensor):
                if dev is None:
                    dev = t.device
                elif t.device != dev:
                    raise RuntimeError("b200rime_%s: tensors live on different devices (%s, %s)"
                                       % (name, dev, t.device))

    def one(a):
        if isinstance(a, torch.Tensor):
            return ctypes.c_void_p(a.data_ptr())
        if isinstance(a, (list, tuple)):      # host table of device pointers (None -> NULL)
            tab = (ctypes.c_void_p * len(a))(*[t.data_ptr() if t is not None else 0 for t in a])
            return ctypes.cast(tab, ctypes.c_void_p)
        return ctypes.c_void_p(0) if a is None else a
    conv = [one(a) for a in args]
    # launch on the device that owns the tensors and on torch's current stream OF THAT DEVICE
    # (not on whatever device happens to be current in the process)
    with torch.cuda.device(dev):
        _lib.call(name, sfx, *conv, _stream(dev))


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
        base = ((nbl + 63) // 64) * nchunk // 2 + 1
        target = 24 * sm_count
        want = max(self.nt, -(-target // max(base, 1)))
        per = max(self.S, 1) / want
        unit = int(min(max(math.ceil(per / tile) * tile, UNIT_MIN_SRC), UNIT_MAX_SRC))
        rows, ubeg = [], [0]
        for t in range(self.nt):
            s0, s1 = self.toff[t], self.toff[t + 1]
            n = -(-(s1 - s0) // unit)                       # units of this time ...
            if n:
                step = -(-(s1 - s0) // (n * tile)) * tile   # ... of (nearly) equal length
                while s0 < s1:
                    e = min(s0 + step, s1)
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
def _packed_cut(geom, cuts, npix):
    """(cut, pos): cut int32 [S] = sky pixel of every packed source (-1 for padding entries);
    pos int32 [Nt][Npix] = packed index of pixel p at time t (-1 outside the FOV)."""
    dev = geom.device
    cut = torch.full((max(geom.S, 1),), -1, dtype=torch.int32, device=dev)
    pos = torch.full((geom.nt, npix), -1, dtype=torch.int32, device=dev)
    for t, c in enumerate(cuts):
        n = geom.ns[t]
        if n == 0:
            continue
        c = c.to(dev)
        cut[geom.toff[t]:geom.toff[t] + n] = c.to(torch.int32)
        pos[t, c.long()] = torch.arange(geom.toff[t], geom.toff[t] + n, dtype=torch.int32,
                                        device=dev)
    return cut.contiguous(), pos.contiguous()


class InterpTable:
    """Device tables of the interpolated-beam builder for one time group, over the packed
    source axis: FOV cut into the sky, neighbour indices / weights into the beam map, and their
    CSR transpose for the adjoint (built on first backward)."""

    def __init__(self, geom, cuts, inds_list, wgts_list, npix_sky, npix_beam, dtype):
        dev = geom.device
        S = max(geom.S, 1)
        self.nnn = int(inds_list[0].shape[1]) if len(inds_list) and inds_list[0].ndim == 2 else 1
        self.npix_sky, self.npix_beam = npix_sky, npix_beam
        self.cut, self.pos = _packed_cut(geom, cuts, npix_sky)
        self.inds = torch.zeros(S, self.nnn, dtype=torch.int32, device=dev)
        self.wgts = torch.zeros(S, self.nnn, dtype=dtype, device=dev)
        for t, (i, w) in enumerate(zip(inds_list, wgts_list)):
            n = geom.ns[t]
            if n:
                sl = slice(geom.toff[t], geom.toff[t] + n)
                self.inds[sl] = i.to(dev).reshape(n, self.nnn).to(torch.int32)
                self.wgts[sl] = w.to(dev).reshape(n, self.nnn).to(dtype)
        self._csr = None

    def csr(self):
        if self._csr is None:
            live = (self.cut >= 0).repeat_interleave(self.nnn)
            flat = self.inds.reshape(-1).long()[live]
            src = torch.arange(self.inds.shape[0], device=flat.device).repeat_interleave(self.nnn)[live]
            w = self.wgts.reshape(-1)[live]
            order = torch.argsort(flat, stable=True)
            counts = torch.bincount(flat, minlength=self.npix_beam)
            rowptr = torch.zeros(self.npix_beam + 1, dtype=torch.int32, device=flat.device)
            rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            self._csr = (rowptr, src[order].to(torch.int32).contiguous(), w[order].contiguous())
        return self._csr


class _BuildInterp(torch.autograd.Function):
    """A = interp(bmap) * gather(sky); either operand may be None (factor 1)."""

    @staticmethod
    def forward(ctx, sky, bmap, geom, tab):
        ref = sky if sky is not None else bmap
        _need_cuda(sky, bmap)
        dtype, sfx = ref.dtype, _sfx(ref.dtype)
        sky = sky.contiguous() if sky is not None else None
        bmap = bmap.to(dtype).contiguous() if bmap is not None else None
        nfreq = ref.shape[0]
        kc = _lib.KC[sfx]
        S = max(geom.S, 1)
        A = torch.empty(1, nchunks(nfreq, dtype), S, kc, dtype=dtype, device=ref.device)
        ctx.bT = None
        if geom.S > 0 and bmap is not None:
            # channel-major copy of the beam map (zero padded to whole chunks): neighbour reads
            # become full 128-byte lines instead of 32 scattered pixels per load
            nfp = A.shape[1] * kc
            bT = torch.zeros(bmap.shape[1], nfp, dtype=dtype, device=ref.device)
            bT[:, :nfreq] = bmap.t()
            _call("build_interp_t", sfx, bT, nfp, tab.inds, tab.wgts, tab.nnn, sky,
                  sky.shape[1] if sky is not None else 0, tab.cut, nfreq, geom.S, geom.S, 0,
                  geom.S, A[0])
            if sfx == "f32" and tab.nnn == 4 and sky is not None:
                ctx.bT = bT             # 0.13 GB at C3: kept for the channel-major backward
        elif geom.S > 0:
            _call("build_interp", sfx, None, 0, tab.inds, tab.wgts, tab.nnn, sky, sky.shape[1],
                  tab.cut, nfreq, geom.S, geom.S, 0, geom.S, A[0])
        else:
            A.zero_()
        ctx.save_for_backward(sky, bmap)
        ctx.geom, ctx.tab, ctx.nfreq = geom, tab, nfreq
        return A

    @staticmethod
    def backward(ctx, dA):
        sky, bmap = ctx.saved_tensors
        geom, tab, nfreq = ctx.geom, ctx.tab, ctx.nfreq
        dA = dA.contiguous()
        sfx = _sfx(dA.dtype)
        need_sky = sky is not None and ctx.needs_input_grad[0]
        need_beam = bmap is not None and ctx.needs_input_grad[1]
        dsky = torch.zeros_like(sky) if need_sky else None
        dbmap = torch.zeros_like(bmap) if need_beam else None
        if geom.S > 0 and (need_sky or need_beam):
            S = geom.S
            dIs = torch.empty(nfreq, S, dtype=dA.dtype, device=dA.device) if need_sky else None
            dBI = torch.empty(nfreq, S, dtype=dA.dtype, device=dA.device) if need_beam else None
            if ctx.bT is not None:
                _call("build_interp_bwd_t", sfx, dA[0], ctx.bT, ctx.bT.shape[1], tab.inds, tab.wgts,
                      sky, sky.shape[1], tab.cut, nfreq, S, 0, S, dBI, S, dIs)
            else:
                _call("build_interp_bwd", sfx, dA[0], bmap, bmap.shape[1] if bmap is not None else 0,
                      tab.inds, tab.wgts, tab.nnn, sky, sky.shape[1] if sky is not None else 0,
                      tab.cut, nfreq, S, 0, S, None, dBI, S, dIs)
            if need_sky:
                _call("gather_times", sfx, dIs, S, tab.pos, geom.nt, tab.npix_sky, nfreq, dsky,
                      sky.shape[1])
            if need_beam:
                rowptr, col, val = tab.csr()
                _call("interp_transpose", sfx, dBI, S, rowptr, col, val, tab.npix_beam, nfreq,
                      dbmap, bmap.shape[1])
        return dsky, dbmap, None, None


def build_interp(sky, bmap, geom, tab):
    """sky (Nf, Npix), bmap (Nf, Npb) -> A (1, nchunk, S, KC); one launch for all times.
    sky=None: interpolated beam only; bmap=None: FOV-gathered sky only."""
    return _BuildInterp.apply(sky, bmap, geom, tab)


class AiryTable:
    """Device tables of the Airy builder over the packed source axis: cut,
    sin(min(zen, 90deg)) and sin(az)^2."""

    def __init__(self, geom, cuts, zen_list, az_list, npix_sky, dtype):
        dev = geom.device
        S = max(geom.S, 1)
        self.npix_sky = npix_sky
        self.cut, self.pos = _packed_cut(geom, cuts, npix_sky)
        self.sinzen = torch.zeros(S, dtype=dtype, device=dev)
        self.sin2az = torch.zeros(S, dtype=dtype, device=dev)
        for t, (z, a) in enumerate(zip(zen_list, az_list)):
            n = geom.ns[t]
            if n:
                sl = slice(geom.toff[t], geom.toff[t] + n)
                zen = torch.clamp(z.to(dev, torch.float64) * D2R, max=math.pi / 2)
                az = a.to(dev, torch.float64) * D2R
                self.sinzen[sl] = torch.sin(zen).to(dtype)
                self.sin2az[sl] = (torch.abs(torch.sin(az)) ** 2).to(dtype)


class _BuildAiry(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sky, diam, geom, tab, freqs64, freq_ratio, square, full_grad):
        _need_cuda(sky)
        dtype, sfx = sky.dtype, _sfx(sky.dtype)
        sky = sky.contiguous()
        nfreq = sky.shape[0]
        # the diameters stay on the device (no host synchronisation; CUDA-graph capturable)
        asym = diam.numel() > 1
        d = diam.detach().double().reshape(-1)
        ddev = torch.stack([d[0], d[1] if asym else d[0]]).to(sky.device).contiguous()
        Dew = Dns = 0.0
        kc = _lib.KC[sfx]
        S = max(geom.S, 1)
        A = torch.empty(1, nchunks(nfreq, dtype), S, kc, dtype=dtype, device=sky.device)
        if geom.S > 0:
            _call("build_airy", sfx, Dew, Dns, ddev, float(freq_ratio), int(square), tab.sinzen,
                  tab.sin2az if asym else None, freqs64, sky, sky.shape[1], tab.cut, nfreq, geom.S,
                  geom.S, 0, geom.S, A[0], None, 0)
        ctx.save_for_backward(sky, freqs64, ddev)
        ctx.meta = (geom, tab, Dew, Dns, asym, float(freq_ratio), int(square), int(full_grad),
                    diam.shape, diam.dtype, diam.device)
        return A

    @staticmethod
    def backward(ctx, dA):
        sky, freqs64, diam_dev = ctx.saved_tensors
        geom, tab, Dew, Dns, asym, ratio, square, full_grad, dshape, ddtype, ddev = ctx.meta
        dA = dA.contiguous()
        sfx = _sfx(sky.dtype)
        nfreq = sky.shape[0]
        need_sky, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dsky = torch.zeros_like(sky) if need_sky else None
        gdiam = None
        if geom.S > 0 and (need_sky or need_d):
            S = geom.S
            dIs = torch.empty(nfreq, S, dtype=sky.dtype, device=sky.device) if need_sky else None
            dD = None
            if need_d:
                nb = _lib.lib.b200rime_airy_bwd_blocks(nfreq, S)
                dD = torch.zeros(nb, 2, dtype=torch.float64, device=sky.device)
            _call("build_airy_bwd", sfx, dA[0], Dew, Dns, diam_dev, ratio, square, full_grad, tab.sinzen,
                  tab.sin2az if asym else None, freqs64, sky, sky.shape[1], tab.cut, nfreq, S, 0, S,
                  None, dD, dIs, S)
            if need_sky:
                _call("gather_times", sfx, dIs, S, tab.pos, geom.nt, tab.npix_sky, nfreq, dsky,
                      sky.shape[1])
            if need_d:
                gD = dD.sum(0)
                g = gD if asym else gD[:1]
                gdiam = g.to(device=ddev, dtype=ddtype).reshape(dshape)
        elif need_d:
            gdiam = torch.zeros(dshape, dtype=ddtype, device=ddev)
        return dsky, gdiam, None, None, None, None, None, None


def build_airy(sky, diam, geom, tab, freqs64, freq_ratio=1.0, square=True, full_grad=False):
    """sky (Nf, Npix), diam tensor with 1 (D) or 2 (Dew, Dns) elements -> A (1, nchunk, S, KC);
    one launch for all times."""
    return _BuildAiry.apply(sky, diam, geom, tab, freqs64, freq_ratio, square, full_grad)


# ----------------------------------------------------------------------------- Jones sandwich
class _JonesSandwich(torch.autograd.Function):
    """P[a][d] = sum_{b,c} J1[a][b] C[b][c] J2[d][c] on real tiled planes (jones_sandwich kernels).
    Inputs: 4 planes of J1, 4 of J2 (the same tensors when both antennas share a beam model),
    4 of C, index 2 * row + col; output (4, *plane shape)."""

    @staticmethod
    def forward(ctx, same, *planes):
        _need_cuda(*planes)
        planes = [x.contiguous() for x in planes]
        j1, j2, c = planes[0:4], planes[4:8], planes[8:12]
        ref = j1[0]
        sfx = _sfx(ref.dtype)
        out = torch.empty((4,) + tuple(ref.shape), dtype=ref.dtype, device=ref.device)
        n = ref.numel()
        if n % 4:
            raise ValueError("jones_sandwich needs plane sizes that are multiples of 4 (tiled layout)")
        _call("jones_sandwich", sfx, j1, j2, c, n, [out[m] for m in range(4)])
        ctx.save_for_backward(*planes)
        ctx.same = bool(same)
        return out

    @staticmethod
    def backward(ctx, dP):
        planes = ctx.saved_tensors
        j1, j2, c = list(planes[0:4]), list(planes[4:8]), list(planes[8:12])
        dP = dP.contiguous()
        sfx = _sfx(dP.dtype)
        n = j1[0].numel()
        need1 = any(ctx.needs_input_grad[1:5])
        need2 = any(ctx.needs_input_grad[5:9]) and not ctx.same
        needc = any(ctx.needs_input_grad[9:13])
        g1 = torch.empty_like(dP) if need1 or (ctx.same and any(ctx.needs_input_grad[5:9])) else None
        g2 = torch.empty_like(dP) if need2 else None
        gc = torch.empty_like(dP) if needc else None
        tab = lambda g: [g[m] for m in range(4)] if g is not None else None
        _call("jones_sandwich_bwd", sfx, [dP[m] for m in range(4)], j1, j2, c, n, int(ctx.same),
              tab(g1), tab(g2), tab(gc))
        un = lambda g: [g[m] for m in range(4)] if g is not None else [None] * 4
        # same: the summed Jones gradient goes to the first operand; the second (the same
        # tensors) gets none, autograd adds nothing twice
        return (None,) + tuple(un(g1)) + tuple(un(g2)) + tuple(un(gc))


def jones_sandwich(J1, J2, C):
    """J1, J2: 2 x 2 real Jones planes (a (2, 2, *shape) tensor or nested lists of plane tensors;
    J2 may be J1), C: 2 x 2 real coherency planes -> P (2, 2, *shape) = J1 C J2^T
    (beam_model.py:363)."""
    same = J2 is J1
    flat = lambda X: [X[a][b] for a in range(2) for b in range(2)]
    out = _JonesSandwich.apply(same, *(flat(J1) + flat(J2) + flat(C)))
    return out.reshape((2, 2) + tuple(out.shape[1:]))


# ----------------------------------------------------------------------------- fringe sum
def _blv4(blvecs, device):
    b = torch.zeros(blvecs.shape[0], 4, dtype=torch.float64, device=device)
    b[:, :3] = blvecs.detach().to(device=device, dtype=torch.float64)
    return b


def _unit_batches(ubeg, nt, bytes_per_unit):
    """Launch plan whose unit partials fit the workspace budget (and the grid limit):
    [(t0, t1, u0, u1, accumulate)] -- whole times t0..t1 with units u0..u1 where they fit; a time
    with more units than the budget allows is split into several launches of that one time, each
    reduced into the output with accumulate = 1 after the first.  Also returns the largest
    number of units in one launch."""
    max_units = min(max(1, VPART_BUDGET // max(bytes_per_unit, 1)), MAX_GRID_UNITS)
    t0, plan = 0, []
    while t0 < nt:
        n0 = ubeg[t0 + 1] - ubeg[t0]
        if n0 > max_units:
            for u in range(ubeg[t0], ubeg[t0 + 1], max_units):
                plan.append((t0, t0 + 1, u, min(u + max_units, ubeg[t0 + 1]), int(u > ubeg[t0])))
            t0 += 1
            continue
        t1 = t0 + 1
        while t1 < nt and ubeg[t1 + 1] - ubeg[t0] <= max_units:
            t1 += 1
        plan.append((t0, t1, ubeg[t0], ubeg[t1], 0))
        t0 = t1
    return plan, max(u1 - u0 for _, _, u0, u1, _ in plan)


def _final_flags(plan):
    """True for the launches after which their times are complete (a time split over several
    launches is complete after the last of them)."""
    return [not (k + 1 < len(plan) and plan[k + 1][4] == 1) for k in range(len(plan))]


class _ChisqEpilogue:
    """State of the fused likelihood epilogue (reduce_units_chisq): data D and weights W in the
    layout of V (nplane, nbl, nt, nfreq), and the float64 block partials collected per launch."""

    def __init__(self, data, icov, rdtype):
        self.D = torch.view_as_real(data.contiguous())
        self.W = icov.to(rdtype).contiguous() if icov is not None else None
        self.parts = []

    def reduce(self, sfx, vpart, ub, ntimes, nbl, nfreq, Vr, p, ta, nt, accum, final):
        if not final:
            _call("reduce_units", sfx, vpart, ub, ntimes, nbl, nfreq, Vr[p, :, ta:], nt * nfreq,
                  nfreq, 1, 1.0, 0.0, accum)
            return
        nb = _lib.lib.b200rime_chisq_blocks(nbl, nfreq)
        part = torch.empty(ntimes, nb, dtype=torch.float64, device=vpart.device)
        _call("reduce_units_chisq", sfx, vpart, ub, ntimes, nbl, nfreq, Vr[p, :, ta:],
              self.D[p, :, ta:], self.W[p, :, ta:] if self.W is not None else None, nt * nfreq,
              nfreq, 1, accum, part)
        self.parts.append(part)

    def total(self, device):
        if not self.parts:
            return torch.zeros((), dtype=torch.float64, device=device)
        return torch.stack([q.sum() for q in self.parts]).sum()


_UBEG_CACHE = {}      # small index tables, uploaded once (no host-to-device copy per step)


def _batch_ubeg(ubeg, ta, tb, u0, u1, device):
    """Unit offsets of the times ta..tb of a launch, relative to its first unit u0 (a split time
    has the single range [0, u1 - u0))."""
    key = (tuple(ubeg[ta:tb + 1]), u0, u1, str(device))
    if key not in _UBEG_CACHE:
        if len(_UBEG_CACHE) > 256:
            _UBEG_CACHE.clear()
        rel = np.clip(np.asarray(ubeg[ta:tb + 1], dtype=np.int64), u0, u1) - u0
        _UBEG_CACHE[key] = torch.as_tensor(rel.astype(np.int32), device=device)
    return _UBEG_CACHE[key]


def _run_fwd_baseline(A, blv, geom, freqs64, nfreq, conj, uniform, epilogue=None):
    """Launches of the baseline-owned forward: V (nplane, nbl, nt, nfreq) complex, or -- with a
    _ChisqEpilogue -- the cotangent 2 W (V - D) in its place."""
    dtype, sfx = A.dtype, _sfx(A.dtype)
    dev = A.device
    nplane, nchunk = A.shape[0], A.shape[1]
    nfp = nchunk * _lib.KC[sfx]
    nbl, nt = blv.shape[0], geom.nt
    V = torch.zeros(nplane, nbl, nt, nfreq, dtype=_cplx(dtype), device=dev)
    if nbl > 0 and nt > 0 and geom.S > 0:
        units, ubeg = geom.units(nbl, nchunk, sm_count(dev))
        esize = 8 if sfx == "f32" else 16
        batches, nu_max = _unit_batches(ubeg, nt, nbl * nfp * esize)
        final = _final_flags(batches)
        vpart = torch.empty(nu_max, nbl, nfp, 2, dtype=dtype, device=dev)
        Vr = torch.view_as_real(V)
        for (ta, tb, u0, u1, accum), fin in zip(batches, final):
            ub = _batch_ubeg(ubeg, ta, tb, u0, u1, dev)
            for p in range(nplane):
                _call("fringe_sum_fwd", sfx, A[p], geom.shat, blv, freqs64,
                      units[u0:], u1 - u0, nbl, nfreq, geom.S, int(conj), int(uniform),
                      vpart)
                if epilogue is not None:
                    epilogue.reduce(sfx, vpart, ub, tb - ta, nbl, nfreq, Vr, p, ta, nt, accum, fin)
                else:
                    _call("reduce_units", sfx, vpart, ub, tb - ta, nbl, nfreq,
                          Vr[p, :, ta:], nt * nfreq, nfreq, 1, 1.0, 0.0, accum)
    elif epilogue is not None and V.numel():
        raise RuntimeError("fused chi-square needs at least one source inside the field of view")
    return V


class _FringeSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, blvecs, geom, freqs64, nfreq, conj, uniform):
        _need_cuda(A, freqs64)
        A = A.contiguous()
        blv = _blv4(blvecs, A.device)
        V = _run_fwd_baseline(A, blv, geom, freqs64, nfreq, conj, uniform)
        need_bl = ctx.needs_input_grad[1]
        ctx.save_for_backward(A if need_bl else None, blv, freqs64)
        ctx.meta = (geom, nfreq, int(conj), int(uniform), blvecs.dtype, blvecs.device, A.shape)
        return V

    @staticmethod
    def backward(ctx, G):
        A, blv, freqs64 = ctx.saved_tensors
        geom, nfreq, conj, uniform, bdtype, bdev, ashape = ctx.meta
        need_A, need_bl = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dA, dbl = _fringe_backward(G, A, blv, geom, freqs64, nfreq, conj, uniform, ashape, need_A,
                                   need_bl)
        gbl = dbl.to(device=bdev, dtype=bdtype) if need_bl else None
        return dA, gbl, None, None, None, None, None


def _fringe_backward(G, A, blv, geom, freqs64, nfreq, conj, uniform, ashape, need_A, need_bl):
    """Adjoints of the fringe sum for a cotangent G (nplane, Nbl, Nt, Nf) complex:
    dA[p, k, s] = sum_b Re(conj(F_bsk) G[p, b, t(s), k]) in the tiled layout and, with the saved
    perceived sky A, dL/d(blvecs) (Nbl, 3) float64."""
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
    return dA, dbl


# ----------------------------------------------------------- antenna-factorised fringe sum
ANT_FWD_MIN_FILL = 0.55     # wanted pairs / computed pairs below which the baseline-owned K1 wins
ANT_BWD_MIN_FILL = 0.40     # Nbl / na_pad^2 below which K2 + K3 win
ANT_H_BUDGET = 4 << 30      # bytes of Hermitian cotangent matrix per backward sub-batch of times


def _xpos(a):
    """Position of antenna slot a (0..63) inside a 64-wide operand row (include/b200rime.h)."""
    return (((a >> 1) & 3) << 4) | ((a >> 3) << 1) | (a & 1)


class AntTiling:
    """Tiling of a baseline list over 64 x 64 blocks of antenna pairs for the antenna-factorised
    kernels.  i_idx / j_idx: antenna row (into antvecs) of the first / second antenna of every
    baseline, whose vector is antvecs[j] - antvecs[i] (telescope_model.py:221-239)."""

    def __init__(self, i_idx, j_idx, na, device):
        T = _lib.ANT_TILE
        i = np.asarray(i_idx, dtype=np.int64)
        j = np.asarray(j_idx, dtype=np.int64)
        nbl = len(i)
        self.nbl, self.na = nbl, int(na)
        self.nblk = max(1, -(-self.na // T))
        self.na_pad = self.nblk * T
        mq = max(16, _lib.ANT_STAGE)
        self.nm_pad = -(-self.na // mq) * mq          # partner axis of the backward, padded to 16
        # output rows the backward computes: a last block of <= 32 antennas runs half-width
        self.bwd_rows = self.na_pad - (T // 2 if self.na - (self.nblk - 1) * T <= T // 2 else 0)
        bi, bj = i // T, j // T
        swap = (bi > bj) | ((bi == bj) & (i > j))        # fold onto the upper triangle
        x = np.where(swap, j, i)
        y = np.where(swap, i, j)
        key = (x // T) * self.nblk + (y // T)
        uniq, tid = np.unique(key, return_inverse=True)
        self.ntile = len(uniq)
        ar = np.arange(T)
        tile_ant = -np.ones((self.ntile, 2 * T), dtype=np.int32)
        for n, kq in enumerate(uniq):
            xb, yb = divmod(int(kq), self.nblk)
            ax, ay = xb * T + ar, yb * T + ar
            tile_ant[n, :T] = np.where(ax < self.na, ax, -1)
            tile_ant[n, T:] = np.where(ay < self.na, ay, -1)
        flat = tid * T * T + (x % T) * T + (y % T)
        self.unique = len(np.unique(flat)) == nbl         # a pair listed twice cannot be tiled
        tile_bl = -np.ones(self.ntile * T * T, dtype=np.int32)
        tile_bl[flat] = (np.arange(nbl) << 1) | swap
        # tiles whose second antennas all sit in slots 0..31 keep one of the two consumer warps
        # idle: they go last so that the four pipelines of a CTA get equal work
        tb = tile_bl.reshape(self.ntile, T, T)
        both = (tb[:, :, T // 2:] >= 0).any(axis=(1, 2))
        order = np.concatenate([np.nonzero(both)[0], np.nonzero(~both)[0]]).astype(np.int32)
        cost = both.sum() + 0.5 * (~both).sum()
        self.pair_slots = float(cost * T * T)            # antenna pairs the forward computes
        self.fill_fwd = nbl / max(self.pair_slots, 1)
        self.fill_bwd = nbl / float(self.na_pad ** 2)
        self.usable = (self.unique and nbl > 0 and self.fill_fwd >= ANT_FWD_MIN_FILL
                       and self.fill_bwd >= ANT_BWD_MIN_FILL)
        dev = torch.device(device)
        self.tile_ant = torch.as_tensor(tile_ant, device=dev)
        self.tile_bl = torch.as_tensor(tile_bl.reshape(self.ntile, T, T), device=dev)
        self.tile_order = torch.as_tensor(order, device=dev)
        self.i = torch.as_tensor(i, device=dev)
        self.j = torch.as_tensor(j, device=dev)
        # scatter coordinates of the Hermitian cotangent matrix (see include/b200rime.h):
        # row a -> (block, position in block), column m -> (stage, index in stage)
        st = _lib.ANT_STAGE
        pos = np.asarray([_xpos(a) for a in range(T)], dtype=np.int64)
        t = lambda v: torch.as_tensor(v, device=dev)
        self.h_ji = (t(j // T), t(i // st), t(i % st), t(pos[j % T]))     # H[a = j][m = i] = G
        cross = i != j
        self.h_cross = t(np.nonzero(cross)[0])
        self.h_auto = t(np.nonzero(~cross)[0])
        ic, jc = i[cross], j[cross]
        self.h_ij = (t(ic // T), t(jc // st), t(jc % st), t(pos[ic % T]))  # H[a = i][m = j] = conj G
        self.h_ij_all = (t(i // T), t(j // st), t(j % st), t(pos[i % T]))
        self.h_jgt = t(np.nonzero(j > i)[0])      # pairs stored at H[j][i] in the lower triangle
        self.h_igt = t(np.nonzero(i > j)[0])      # pairs stored at H[i][j] (conjugated)

    def antv4(self, antvecs):
        out = torch.zeros(self.na_pad, 4, dtype=torch.float64, device=self.tile_ant.device)
        out[:self.na, :3] = antvecs.detach().to(out.device, torch.float64)
        return out

    def hermitian_cotangent(self, G, nfp, lower_only=False):
        """G (nbl, nt, nf) complex64 -> Hp in the kernel layout (float32 view).  lower_only: the
        doubled lower triangle (a > m), enough for dL/dA alone (include/b200rime.h)."""
        st, T = _lib.ANT_STAGE, _lib.ANT_TILE
        nbl, nt, nf = G.shape
        if nfp != nf:
            G = torch.nn.functional.pad(torch.view_as_real(G), (0, 0, 0, nfp - nf))
            G = torch.view_as_complex(G)
        Gq = G.permute(1, 2, 0)                                   # (nt, nfp, nbl)
        H = torch.zeros(nt, nfp, self.nblk, self.nm_pad // st, st, T, dtype=G.dtype,
                        device=G.device)
        if lower_only:
            # pair (i, j): H[j][i] = 2 G if j > i, else H[i][j] = 2 conj(G)
            for idx, sel, conj in ((self.h_ji, self.h_jgt, False), (self.h_ij_all, self.h_igt, True)):
                if len(sel) == 0:
                    continue
                b, ms, r, pa = [v.index_select(0, sel) for v in idx]
                val = 2 * Gq.index_select(2, sel)
                H[:, :, b, ms, r, pa] = val.conj() if conj else val
        else:
            b, ms, r, pa = self.h_ji
            H[:, :, b, ms, r, pa] = Gq
            b, ms, r, pa = self.h_ij
            H[:, :, b, ms, r, pa] = Gq.index_select(2, self.h_cross).conj()
        if len(self.h_auto):
            b, ms, r, pa = [v.index_select(0, self.h_auto) for v in self.h_ji]
            H[:, :, b, ms, r, pa] = (2 * Gq.index_select(2, self.h_auto).real).to(G.dtype)
        return torch.view_as_real(H)


TC_MIN_FILL = 0.12        # wanted pairs / computed pairs below which the tensor-core items lose


class TcTiling:
    """Items of the tensor-core fringe-sum kernel for a baseline list: blocks of 128 first
    antennas against ranges of at most 256 second antennas, and the antenna-pair -> baseline
    table (include/b200rime.h, tcfringe_fwd).  Pairs are folded onto i <= j (a baseline listed
    as (j, i) is computed as (i, j) and conjugated on output)."""

    def __init__(self, i_idx, j_idx, na, device):
        M, NMAX = _lib.TC_ROWS, _lib.TC_COLS_MAX
        i = np.asarray(i_idx, dtype=np.int64)
        j = np.asarray(j_idx, dtype=np.int64)
        nbl = len(i)
        self.nbl, self.na = nbl, int(na)
        self.ldp = max(32, -(-self.na // 32) * 32)
        swap = i > j
        x = np.where(swap, j, i)
        y = np.where(swap, i, j)
        flat = x * self.ldp + y
        self.unique = len(np.unique(flat)) == nbl
        pair = -np.ones(self.ldp * self.ldp, dtype=np.int32)
        pair[flat] = (np.arange(nbl) << 1) | swap
        items, slots = [], 0
        for ib in range(-(-self.na // M)):
            sel = (x // M) == ib
            if not sel.any():
                continue
            lo = int(y[sel].min() // 32) * 32
            hi = int(-(-(y[sel].max() + 1) // 32)) * 32
            npiece = -(-(hi - lo) // NMAX)
            per = -(-(hi - lo) // (npiece * 32)) * 32
            j0 = lo
            while j0 < hi:
                n = min(per, hi - j0)
                items.append((ib * M, j0, n, 0))
                slots += M * n
                j0 += n
        self.nitems = len(items)
        self.pair_slots = float(slots)
        self.fill = nbl / max(self.pair_slots, 1.0)
        self.usable = self.unique and nbl > 0 and self.fill >= TC_MIN_FILL
        dev = torch.device(device)
        self.items = torch.as_tensor(np.asarray(items, dtype=np.int32).reshape(-1, 4), device=dev)
        self.pair_bl = torch.as_tensor(pair.reshape(self.ldp, self.ldp), device=dev)
        # backward (include/b200rime.h, tcfringe_bwd): items of 128 output antennas, partner axis
        # padded to 16
        self.nitem_bwd = -(-self.na // M)
        self.apad = self.nitem_bwd * M
        self.nm_pad = -(-self.na // 16) * 16
        self.bwd_usable = self.usable and self.na <= 512
        self.i = torch.as_tensor(i, device=dev)
        self.j = torch.as_tensor(j, device=dev)
        self.cross = torch.as_tensor(np.nonzero(i != j)[0], device=dev)
        self.auto = torch.as_tensor(np.nonzero(i == j)[0], device=dev)
        self.jgt = torch.as_tensor(np.nonzero(j > i)[0], device=dev)
        self.igt = torch.as_tensor(np.nonzero(i > j)[0], device=dev)
        # stages of 16 partner antennas m that hold cotangent entries, per item of 128 antennas a
        # [lo, hi): full matrix (both triangles) and doubled lower triangle (a > m).  A baseline
        # group that covers only some antenna blocks skips the empty stages.
        def stage_ranges(a_idx, m_idx):
            out = np.zeros((self.nitem_bwd, 2), dtype=np.int32)
            for ib in range(self.nitem_bwd):
                sel = (a_idx // M) == ib
                if sel.any():
                    out[ib] = (m_idx[sel].min() // 16, m_idx[sel].max() // 16 + 1)
            return out
        both_a, both_m = np.concatenate([j, i]), np.concatenate([i, j])
        self.mrange_full = torch.as_tensor(stage_ranges(both_a, both_m), device=dev)
        self.mrange_lower = torch.as_tensor(stage_ranges(y, x), device=dev)
        self.bwd_stages_full = int((self.mrange_full[:, 1] - self.mrange_full[:, 0]).sum())
        self.bwd_stages_lower = int((self.mrange_lower[:, 1] - self.mrange_lower[:, 0]).sum())

    def antv4(self, antvecs):
        out = torch.zeros(self.apad, 4, dtype=torch.float64, device=self.items.device)
        out[:self.na, :3] = antvecs.detach().to(out.device, torch.float64)[:self.na]
        return out

    def cotangent_operand(self, G, nfp, lower_only=False):
        """G (nbl, nt, nf) complex64 -> (Hq, hscale): the Hermitian cotangent matrix
        H[a][m] = G_b for b = (m, a), conj(G_b) for b = (a, m), 2 Re G_b for autos (lower_only: the
        doubled lower triangle a > m), scaled by the power of two hscale into float16 range,
        split into float16 hi / lo parts and laid out as the stacked UMMA B operands of
        tcfringe_bwd, the three-half buffers (-Hi ; Hr ; Hi) whose rows 0..255 / 128..383 are the
        operands M = (-Hi ; Hr) / P = (Hr ; Hi):
        [nt][Nfp][item][stage of 16 m][hi | lo][3 halves][16 row groups][2 k groups][8 rows][8 k]."""
        nbl, nt, nf = G.shape
        if G.dtype != torch.complex64:
            G = G.to(torch.complex64)
        if G.stride(2) != 1 or G.stride(1) != nf:
            G = G.contiguous()
        # max |H| <= 2 max(|Re G|, |Im G|) when entries are doubled (autos, lower triangle)
        dbl = 2.0 if (lower_only or len(self.auto)) else 1.0
        amax = (torch.view_as_real(G).abs().amax() * dbl).clamp_min(1e-30)
        hscale = torch.exp2(14.0 - torch.floor(torch.log2(amax))).to(torch.float32).reshape(1)
        Hq = torch.empty(nt, nfp, self.nitem_bwd, self.nm_pad // 16, 6, 16, 2, 8, 8,
                         dtype=torch.float16, device=G.device)
        _call("tc_pack_cotangent", "f32", G, G.stride(0), self.pair_bl, self.ldp, nt, nf, self.na,
              self.nm_pad, int(lower_only), hscale, Hq)
        return Hq, hscale


def tc_scale(A):
    """Per-plane power of two that brings max|A| into [2^14, 2^15): the float16 range of the
    tensor-core operands A E (device tensor, no host synchronisation)."""
    amax = A.detach().abs().reshape(A.shape[0], -1).amax(dim=1).clamp_min(1e-30)
    return torch.exp2(14.0 - torch.floor(torch.log2(amax))).to(torch.float32).contiguous()


def _run_fwd_ant(A, antv, geom, freqs64, nfreq, conj, tiling, tc, epilogue=None):
    """Launches of the antenna-factorised forward (tensor-core items when tc is given)."""
    dev = A.device
    nplane, nchunk = A.shape[0], A.shape[1]
    nfp = nchunk * _lib.KC["f32"]
    plan = tiling if tiling is not None else tc
    nbl, nt = plan.nbl, geom.nt
    V = torch.zeros(nplane, nbl, nt, nfreq, dtype=torch.complex64, device=dev)
    if nbl > 0 and nt > 0 and geom.S > 0:
        units, ubeg = geom.units(nbl, nchunk, sm_count(dev))
        batches, nu_max = _unit_batches(ubeg, nt, nbl * nfp * 8)
        final = _final_flags(batches)
        # pairs outside the tiling keep their zero: every wanted pair has exactly one owner
        vpart = torch.empty(nu_max, nbl, nfp, 2, dtype=torch.float32, device=dev)
        Vr = torch.view_as_real(V)
        if tc is not None:
            # channel-major copy: a stage's 16 sky values become one contiguous TMA copy
            ascale = tc_scale(A)
            acm = A.permute(0, 1, 3, 2).reshape(nplane, nfp, A.shape[2]).contiguous()
        for (ta, tb, u0, u1, accum), fin in zip(batches, final):
            ub = _batch_ubeg(ubeg, ta, tb, u0, u1, dev)
            for p in range(nplane):
                if tc is not None:
                    _call("tcfringe_fwd", "f32", acm[p], ascale[p:p + 1], geom.shat, antv,
                          freqs64, units[u0:], u1 - u0, tc.items, tc.nitems, tc.pair_bl,
                          tc.ldp, tc.na, nbl, nfreq, geom.S, int(conj), vpart)
                else:
                    _call("antfringe_fwd", "f32", A[p], geom.shat, antv, freqs64, units[u0:],
                          u1 - u0, tiling.tile_ant, tiling.tile_bl, tiling.tile_order,
                          tiling.ntile, nbl, nfreq, geom.S, int(conj), vpart)
                if epilogue is not None:
                    epilogue.reduce("f32", vpart, ub, tb - ta, nbl, nfreq, Vr, p, ta, nt, accum, fin)
                else:
                    _call("reduce_units", "f32", vpart, ub, tb - ta, nbl, nfreq,
                          Vr[p, :, ta:], nt * nfreq, nfreq, 1, 1.0, 0.0, accum)
    elif epilogue is not None and V.numel():
        raise RuntimeError("fused chi-square needs at least one source inside the field of view")
    return V


def _ant_route_backward(G, A, antv, geom, freqs64, nfreq, conj, tiling, tc, need_A, need_r):
    if tc is not None and (tc.bwd_usable or tiling is None):
        return _tc_backward(G, A, antv, geom, freqs64, nfreq, conj, tc, A.shape, need_A, need_r)
    return _ant_backward(G, A, antv, geom, freqs64, nfreq, conj, tiling, A.shape, need_A, need_r)


class _AntFringeSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, antvecs, geom, freqs64, nfreq, conj, tiling, tc=None):
        _need_cuda(A, freqs64)
        if A.dtype != torch.float32:
            raise TypeError("fringe_sum_ant is float32 only")
        A = A.contiguous()
        antv = (tiling if tiling is not None else tc).antv4(antvecs)
        V = _run_fwd_ant(A, antv, geom, freqs64, nfreq, conj, tiling, tc)
        ctx.save_for_backward(A, antv, freqs64)
        ctx.meta = (geom, nfreq, int(conj), tiling, antvecs.dtype, antvecs.device, antvecs.shape,
                    tc)
        return V

    @staticmethod
    def backward(ctx, G):
        A, antv, freqs64 = ctx.saved_tensors
        geom, nfreq, conj, tiling, adtype, adev, ashape_ant, tc = ctx.meta
        need_A, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dA, dr = _ant_route_backward(G, A, antv, geom, freqs64, nfreq, conj, tiling, tc, need_A,
                                     need_r)
        gant = None
        if need_r:
            gant = torch.zeros(ashape_ant, dtype=torch.float64, device=G.device)
            na = (tiling if tiling is not None else tc).na
            gant[:na] = dr[:na]
            gant = gant.to(device=adev, dtype=adtype)
        return dA, gant, None, None, None, None, None, None


class _FringeChisq(torch.autograd.Function):
    """chisq = sum icov |V - data|^2 with V the fringe sum of A, fused: the unit reduction writes
    the cotangent 2 icov (V - data) instead of V (reduce_units_chisq), the backward kernels start
    from it.  vecs: baseline vectors (baseline-owned kernels) or antenna positions (tiling / tc)."""

    @staticmethod
    def forward(ctx, A, vecs, geom, freqs64, nfreq, conj, uniform, tiling, tc, data, icov):
        _need_cuda(A, freqs64, data, icov)
        A = A.contiguous()
        ant = tiling is not None or tc is not None
        if ant and A.dtype != torch.float32:
            raise TypeError("the antenna-factorised kernels are float32 only")
        epi = _ChisqEpilogue(data.to(_cplx(A.dtype)), icov, A.dtype)
        if ant:
            v4 = (tiling if tiling is not None else tc).antv4(vecs)
            G = _run_fwd_ant(A, v4, geom, freqs64, nfreq, conj, tiling, tc, epilogue=epi)
        else:
            v4 = _blv4(vecs, A.device)
            G = _run_fwd_baseline(A, v4, geom, freqs64, nfreq, conj, uniform, epilogue=epi)
        if tuple(G.shape) != tuple(data.shape):
            raise ValueError("data %s does not match the visibilities %s" % (tuple(data.shape),
                                                                              tuple(G.shape)))
        ctx.save_for_backward(A, v4, freqs64, G)
        ctx.meta = (geom, nfreq, int(conj), int(uniform), tiling, tc, vecs.dtype, vecs.device,
                    vecs.shape)
        ctx.mark_non_differentiable(G)
        return epi.total(A.device), G

    @staticmethod
    def backward(ctx, gchi, _gG):
        A, v4, freqs64, G = ctx.saved_tensors
        geom, nfreq, conj, uniform, tiling, tc, vdtype, vdev, vshape = ctx.meta
        need_A, need_v = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if tiling is not None or tc is not None:
            dA, dv = _ant_route_backward(G, A, v4, geom, freqs64, nfreq, conj, tiling, tc, need_A,
                                         need_v)
        else:
            dA, dv = _fringe_backward(G, A, v4, geom, freqs64, nfreq, conj, uniform, A.shape,
                                      need_A, need_v)
        if need_A:
            dA = dA * gchi.to(dA.dtype)
        gv = None
        if need_v:
            gv = torch.zeros(vshape, dtype=torch.float64, device=G.device)
            n = min(vshape[0], dv.shape[0])
            gv[:n] = dv[:n, :3] * gchi.to(torch.float64)
            gv = gv.to(device=vdev, dtype=vdtype)
        return dA, gv, None, None, None, None, None, None, None, None, None


def fringe_chisq(A, vecs, geom, freqs64, nfreq, data, icov=None, conj=False, uniform=True,
                 tiling=None, tc=None):
    """(chisq, G): chisq = sum icov |V - data|^2 (float64 scalar, differentiable with respect to A
    and vecs) and the cotangent G = 2 icov (V - data), for V = fringe_sum(A, ...) -- the fused
    form of optim.LogProb.forward_chisq (optim.py:1012-1024) on the visibilities of the RIME.
    The residual is V - data = G / (2 icov)."""
    return _FringeChisq.apply(A, vecs, geom, freqs64, nfreq, conj, uniform, tiling, tc, data, icov)


def _ant_backward(G, A, antv, geom, freqs64, nfreq, conj, tiling, ashape, need_A, need_r):
    """Antenna-factorised adjoints for a cotangent G (nplane, Nbl, Nt, Nf) complex64: dA in the
    tiled layout and dL/d(antenna positions) (na_pad, 3) float64 (needs the perceived sky A)."""
    dev = G.device
    nplane, nchunk, S, kc = ashape
    nfp = nchunk * kc
    nt = geom.nt
    dA = torch.zeros(ashape, dtype=torch.float32, device=dev) if need_A else None
    dr = torch.zeros(tiling.na_pad, 3, dtype=torch.float64, device=dev) if need_r else None
    if tiling.nbl > 0 and nt > 0 and geom.S > 0 and (need_A or need_r):
        units, ubeg = geom.units(tiling.nbl, nchunk, sm_count(dev))
        per_time = nfp * tiling.na_pad * tiling.nm_pad * 8
        tstep = max(1, ANT_H_BUDGET // per_time)
        G = G.contiguous()
        for p in range(nplane):
            dApart = (torch.zeros(tiling.nblk, nchunk, S, kc, dtype=torch.float32, device=dev)
                      if need_A else None)
            for ta in range(0, nt, tstep):
                tb = min(nt, ta + tstep)
                u0, u1 = ubeg[ta], ubeg[tb]
                if u1 == u0:
                    continue
                Hp = tiling.hermitian_cotangent(G[p, :, ta:tb], nfp, lower_only=not need_r)
                un = units[u0:u1].clone()
                un[:, 0] -= ta
                drpart = (torch.zeros(u1 - u0, nfp, 2, tiling.na_pad, 4, dtype=torch.float64,
                                      device=dev) if need_r else None)
                _call("antfringe_bwd", "f32", Hp, A[p] if need_r else None, geom.shat, antv,
                      freqs64, un, u1 - u0, tiling.na, tiling.na_pad, tiling.nm_pad, nfreq,
                      geom.S, conj, dApart, drpart)
                if need_r:
                    dr = dr + drpart.sum(dim=(0, 1, 2))[:, :3]
                del Hp
            if need_A:
                dA[p] = dApart.sum(0)
    return dA, dr


TC_H_BUDGET = 3 << 30       # bytes of dense cotangent matrix per backward sub-batch of times


def _tc_backward(G, A, antv, geom, freqs64, nfreq, conj, tc, ashape, need_A, need_r):
    """Tensor-core adjoints (tcfringe_bwd) for a cotangent G (nplane, Nbl, Nt, Nf) complex64: dA in
    the tiled layout and dL/d(antenna positions) (rows, 3) float64 (needs the perceived sky A)."""
    dev = G.device
    nplane, nchunk, S, kc = ashape
    nfp = nchunk * kc
    nt = geom.nt
    dA = torch.zeros(ashape, dtype=torch.float32, device=dev) if need_A else None
    dr = torch.zeros(antv.shape[0], 3, dtype=torch.float64, device=dev) if need_r else None
    if tc.nbl > 0 and nt > 0 and geom.S > 0 and (need_A or need_r):
        units, ubeg = geom.units(tc.nbl, nchunk, sm_count(dev))
        per_time = nfp * tc.apad * tc.nm_pad * 8
        tstep = max(1, TC_H_BUDGET // per_time)
        G = G.contiguous()
        for p in range(nplane):
            dAcm = (torch.zeros(tc.nitem_bwd, nfp, S, dtype=torch.float32, device=dev)
                    if need_A else None)
            acm = A[p].permute(0, 2, 1).reshape(nfp, S).contiguous() if need_r else None
            for ta in range(0, nt, tstep):
                tb = min(nt, ta + tstep)
                u0, u1 = ubeg[ta], ubeg[tb]
                if u1 == u0:
                    continue
                Hq, hscale = tc.cotangent_operand(G[p, :, ta:tb], nfp, lower_only=not need_r)
                un = units[u0:u1].clone()
                un[:, 0] -= ta
                drpart = (torch.zeros(u1 - u0, nfp, 4, tc.apad, 4, dtype=torch.float32, device=dev)
                          if need_r else None)
                _call("tcfringe_bwd", "f32", Hq, hscale, acm, geom.shat, antv, freqs64, un,
                      u1 - u0, tc.nitem_bwd, tc.na, tc.nm_pad,
                      tc.mrange_full if need_r else tc.mrange_lower, nfreq, geom.S, conj,
                      dAcm, drpart)
                if need_r:
                    dr[:tc.na] += drpart.sum(dim=(0, 1, 2), dtype=torch.float64)[:tc.na, :3]
                del Hq
            if need_A:
                dA[p] = dAcm.sum(0).reshape(nchunk, kc, S).permute(0, 2, 1)
    return dA, dr


def fringe_sum_ant(A, antvecs, tiling, geom, freqs64, nfreq, conj=False, tc=None):
    """Same sum as fringe_sum for the baselines (tiling.i, tiling.j) of antenna positions
    antvecs (Na, 3), through the antenna-factorised float32 kernels; gradients flow to A and
    straight to antvecs.  tc (TcTiling of the same baseline list): the sum and its adjoints run
    on the tensor cores (tcgen05) instead of the FP32 pipes; tiling may then be None."""
    return _AntFringeSum.apply(A, antvecs, geom, freqs64, nfreq, conj, tiling, tc)


def fringe_adjoint(G, geom, freqs64, nfreq, blvecs=None, antvecs=None, tiling=None, conj=False,
                   uniform=True):
    """The adjoint (imaging) operator of fringe_sum, without autograd: for visibilities
    G (nplane, Nbl, Nt, Nf) complex returns D (nplane, nchunk, S, KC) real with
    D[p, k, s] = sum_b Re(conj(F_bsk) G[p, b, t(s), k]) -- the dirty-map sum of the reference's
    imaging.make_map (imaging.py:736, A = conj(fringe), imaging.py:293).  With `tiling` (and
    antenna positions) the antenna-factorised kernel walks one triangle of the Hermitian matrix."""
    _need_cuda(G, freqs64)
    nchunk = nchunks(nfreq, G.dtype)
    ashape = (G.shape[0], nchunk, max(geom.S, 1), _lib.KC[_sfx(G.dtype)])
    if tiling is not None and G.dtype == torch.complex64:
        return _ant_backward(G, None, tiling.antv4(antvecs), geom, freqs64, nfreq, int(conj),
                             tiling, ashape, True, False)[0]
    blv = _blv4(blvecs, G.device)
    return _fringe_backward(G, None, blv, geom, freqs64, nfreq, int(conj), int(uniform), ashape,
                            True, False)[0]


def unpack_planes(geom, A, nfreq):
    """Tiled A (nplane, nchunk, S, KC) -> list over times of row-major (nplane, Nf, Ns_t)."""
    _need_cuda(A)
    sfx = _sfx(A.dtype)
    A = A.contiguous()
    out = []
    for t in range(geom.nt):
        X = torch.zeros(A.shape[0], nfreq, geom.ns[t], dtype=A.dtype, device=A.device)
        if geom.ns[t]:
            for p in range(A.shape[0]):
                _call("unpack", sfx, A[p], geom.ns[t], nfreq, geom.ns[t], geom.toff[t], geom.S,
                      X[p])
        out.append(X)
    return out


def fringe_sum(A, blvecs, geom, freqs64, nfreq, conj=False, uniform=True):
    """A (nplane, nchunk, S, KC) real, blvecs (Nbl, 3) -> V (nplane, Nbl, Nt, Nf) complex:
    V[p, b, t, f] = sum_s A[p, f, s] exp(+-2 pi i (b . shat_s) nu_f / c)."""
    return _FringeSum.apply(A, blvecs, geom, freqs64, nfreq, conj, uniform)


# ----------------------------------------------------------------------------- a_lm -> map
def _pow2_scale(x):
    """Power of two that brings max|x| into [2^14, 2^15) (device float32 [1], no host sync)."""
    amax = x.detach().abs().amax().clamp_min(1e-30).to(torch.float32)
    return torch.exp2(14.0 - torch.floor(torch.log2(amax))).reshape(1).contiguous()


def _re_im(x):
    """(re view, im view or None, element stride in floats) of a real / complex tensor."""
    if x.is_complex():
        v = torch.view_as_real(x)
        return v[..., 0], v[..., 1], 2
    return x, None, 1


class AlmPlan:
    """The constant operand of AlmModel.forward_alm (sph_harm.py:1289-1373): the Ylm matrix
    (Ncoeff, Npix) of one set of angles, prepared once for both directions of the product.

    float32 sessions: packed for the tensor-core GEMM (b200rime_cgemm_pack_b_f32) as
    Y[pixel][mode] for the forward map and conj(Y)[mode][pixel] for the adjoint (12 bytes per
    complex entry each: float16 hi + lo of (re; im; -re)).  float64 sessions keep the matrix as
    it is (b200rime_cgemm_f64 reads strided operands)."""

    def __init__(self, Ylm):
        _need_cuda(Ylm)
        assert Ylm.ndim == 2
        self.ncoeff, self.npix = int(Ylm.shape[0]), int(Ylm.shape[1])
        self.complex = Ylm.is_complex()
        self.f64 = _real(Ylm.dtype) == torch.float64
        self.device = Ylm.device
        self.Ylm = Ylm.detach().contiguous()
        if self.f64:
            return
        Y = self.Ylm
        re, im, w = _re_im(Y)
        self.scale = _pow2_scale(Y)
        C, P = self.ncoeff, self.npix
        self.Bq_fwd = torch.empty(_lib.lib.b200rime_cgemm_b_bytes(P, C), dtype=torch.uint8,
                                  device=self.device)
        _call("cgemm_pack_b", "f32", re, im, w, w * P, P, C, self.scale, 0, self.Bq_fwd)
        self.Bq_adj = torch.empty(_lib.lib.b200rime_cgemm_b_bytes(C, P), dtype=torch.uint8,
                                  device=self.device)
        _call("cgemm_pack_b", "f32", re, im, w * P, w, C, P, self.scale, 1, self.Bq_adj)
        self.Ylm = None              # the packed copies are all the GEMM needs


def _ksplit(M, N, K, device):
    """Split of the k axis that fills the SMs when there are few output tiles; every split keeps
    at least 8 stages of 16."""
    tiles = ((M + _lib.TC_ROWS - 1) // _lib.TC_ROWS) * ((N + _lib.TC_COLS_MAX - 1) // _lib.TC_COLS_MAX)
    nkst = (K + 15) // 16
    if tiles >= sm_count(device):
        return 1
    return int(max(1, min(nkst // 8, (2 * sm_count(device)) // tiles, 64)))


def _cgemm_f32(X, Bq, bscale, M, N, K, x_conj, real_out):
    """out (M, N) = X (M, K) . Y^T through the packed Y; X real or complex64, row-major."""
    dev = X.device
    re, im, w = _re_im(X.contiguous())
    sa = _pow2_scale(X)
    Aq = torch.empty(_lib.lib.b200rime_cgemm_a_bytes(M, K), dtype=torch.uint8, device=dev)
    # the MMA pairing forms conj(X) Y: pack -Im X for the plain product
    _call("cgemm_pack_a", "f32", re, im, w * K, w, M, K, sa, 0 if x_conj else 1, Aq)
    out = torch.empty((M, N), dtype=torch.float32 if real_out else torch.complex64, device=dev)
    ks = _ksplit(M, N, K, dev)
    part = torch.empty((ks,) + tuple(out.shape), dtype=out.dtype, device=dev) if ks > 1 else None
    _call("cgemm", "f32", Aq, Bq, M, N, K, ks, 1 if im is None else 0, 1 if real_out else 0,
          sa, bscale, out, N, part)
    return out


def _cgemm_f64(X, Y, y_rows_are_k, conj_y, real_out):
    """out (M, N) = X (M, K) . Y'^T, Y' = Y^T (y_rows_are_k: Y is (K, N)) or Y (N, K)."""
    dev = X.device
    X = X.contiguous()
    xr, xi, wx = _re_im(X)
    yr, yi, wy = _re_im(Y)
    M, K = X.shape
    if y_rows_are_k:
        N = Y.shape[1]
        syn, syk = wy, wy * N
    else:
        N = Y.shape[0]
        syn, syk = wy * K, wy
    out = torch.empty((M, N), dtype=torch.float64 if real_out else torch.complex128, device=dev)
    _call("cgemm", "f64", xr, xi, wx * K, wx, yr, yi, syn, syk, M, N, K, 0, 1 if conj_y else 0,
          1 if real_out else 0, out, N)
    return out


class _AlmForward(torch.autograd.Function):
    """out = p Ylm (real part when real_out); backward dL/dp = G conj(Ylm)^T."""

    @staticmethod
    def forward(ctx, p, plan, real_out):
        _need_cuda(p)
        ctx.plan, ctx.p_complex, ctx.real_out = plan, p.is_complex(), real_out
        M, K = p.shape
        assert K == plan.ncoeff, "params (%d modes) do not match Ylm (%d)" % (K, plan.ncoeff)
        if plan.f64:
            return _cgemm_f64(p.to(torch.complex128 if p.is_complex() else torch.float64),
                              plan.Ylm, True, False, real_out)
        p = p.to(torch.complex64 if p.is_complex() else torch.float32)
        return _cgemm_f32(p, plan.Bq_fwd, plan.scale, M, plan.npix, K, False, real_out)

    @staticmethod
    def backward(ctx, G):
        plan = ctx.plan
        G = G.contiguous()
        if ctx.real_out and G.is_complex():
            G = G.real.contiguous()
        M = G.shape[0]
        real_grad = not ctx.p_complex
        if plan.f64:
            g = _cgemm_f64(G, plan.Ylm, False, True, real_grad)
        else:
            g = _cgemm_f32(G, plan.Bq_adj, plan.scale, M, plan.ncoeff, plan.npix, False, real_grad)
        return g, None, None


def alm_forward(p, plan, real_out=False):
    """(..., Ncoeff) coefficients -> (..., Npix) map through the Ylm of `plan`
    (AlmModel.forward_alm's einsum "...i,ij->...j", sph_harm.py:1366)."""
    lead = p.shape[:-1]
    out = _AlmForward.apply(p.reshape(-1, p.shape[-1]), plan, bool(real_out))
    return out.reshape(lead + (plan.npix,))
