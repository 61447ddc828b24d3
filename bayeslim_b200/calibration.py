"""
Gain application, mirroring the reference's ``calibration.apply_cal`` / ``_apply_cal``
(bayeslim/calibration.py:2348-2487): V_out = g_1 V g_2^H per baseline, the step that follows
the RIME in a BayesLIM ``Sequential``.

The product and its adjoints (to the visibilities and to the gains) run in
``csrc/cal_kernels.cu``; the reference builds g_1, g_2 and G = g_1 conj(g_2) as three
visibility-sized temporaries with ``index_select`` (calibration.py:2462-2468).  Inverting the
gains for ``undo`` stays in torch: it acts on the small (Nants, Ntimes, Nfreqs) table.

``JonesModel`` / ``JonesResponse`` mirror the module that calls it (calibration.py:416-875):
parameters -> complex gains (O(Nants Nt Nf), torch), reference-antenna phase, time selection
for minibatches, antenna index lookup, then the CUDA product on the visibilities of a VisData.
"""
import copy
import math
import numpy as np
import torch

from . import ops, utils


class _ApplyCal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vis, gains, g1_idx, g2_idx, full, cov):
        ops._need_cuda(vis, gains)
        sfx = ops._sfx(vis.dtype)
        vis_c = vis.contiguous()
        gains_c = gains.to(vis.dtype).contiguous()
        npol, nbl, nt, nf = vis.shape[0], vis.shape[2], vis.shape[3], vis.shape[4]
        nant, ntg, nfg = gains.shape[2], gains.shape[3], gains.shape[4]
        out = torch.empty_like(vis_c)
        cov_out = None
        if cov is not None:
            cov = cov.to(ops._real(vis.dtype)).contiguous()
            cov_out = torch.empty_like(cov)
        g1 = g1_idx.to(device=vis.device, dtype=torch.int32).contiguous()
        g2 = g2_idx.to(device=vis.device, dtype=torch.int32).contiguous()
        ops._call("apply_cal", sfx, torch.view_as_real(vis_c), torch.view_as_real(gains_c), g1, g2,
                  npol, int(full), nbl, nt, nf, nant, ntg, nfg, cov, torch.view_as_real(out),
                  cov_out)
        ctx.save_for_backward(vis_c, gains_c, g1, g2)
        ctx.full = int(full)
        ctx.gshape, ctx.gdtype = gains.shape, gains.dtype
        if cov_out is None:
            cov_out = torch.empty(0, device=vis.device)
        ctx.mark_non_differentiable(cov_out)
        return out, cov_out

    @staticmethod
    def backward(ctx, gout, _gcov):
        vis, gains, g1, g2 = ctx.saved_tensors
        sfx = ops._sfx(vis.dtype)
        npol, nbl, nt, nf = vis.shape[0], vis.shape[2], vis.shape[3], vis.shape[4]
        nant, ntg, nfg = gains.shape[2], gains.shape[3], gains.shape[4]
        gout = gout.contiguous()
        dvis = dgains = None
        if ctx.needs_input_grad[0]:
            # adjoint to vis = the same product with conjugate-transposed gains
            gH = gains.conj().resolve_conj().transpose(0, 1).contiguous()
            dvis = torch.empty_like(vis)
            ops._call("apply_cal", sfx, torch.view_as_real(gout), torch.view_as_real(gH), g1, g2,
                      npol, ctx.full, nbl, nt, nf, nant, ntg, nfg, None, torch.view_as_real(dvis),
                      None)
        if ctx.needs_input_grad[1]:
            # fixed-order lists of the baselines of every antenna
            g1c, g2c = g1.cpu().numpy(), g2.cpu().numpy()
            dev = vis.device

            def csr(idx):
                order = np.argsort(idx, kind='stable').astype(np.int32)
                ptr = np.zeros(nant + 1, dtype=np.int32)
                np.add.at(ptr, idx + 1, 1)
                return (torch.as_tensor(np.cumsum(ptr).astype(np.int32), device=dev),
                        torch.as_tensor(order, device=dev))
            p1, b1 = csr(g1c)
            p2, b2 = csr(g2c)
            dg = torch.empty(npol, npol, nant, nt, nf, dtype=vis.dtype, device=dev)
            ops._call("apply_cal_bwd_gains", sfx, torch.view_as_real(vis),
                      torch.view_as_real(gains), torch.view_as_real(gout), g1, g2, p1, b1, p2, b2,
                      npol, ctx.full, nbl, nt, nf, nant, ntg, nfg, torch.view_as_real(dg))
            if ntg == 1 and nt > 1:
                dg = dg.sum(dim=3, keepdim=True)
            if nfg == 1 and nf > 1:
                dg = dg.sum(dim=4, keepdim=True)
            dgains = dg.reshape(ctx.gshape).to(ctx.gdtype)
        return dvis, dgains, None, None, None, None


def _apply_cal(vis, gains, g1_idx, g2_idx, cal_2pol=False, cov=None, vis_type='com', undo=False,
               inplace=False):
    """See apply_cal; g1_idx / g2_idx index the Nants axis of gains for the two antennas of every
    baseline (calibration.py:2412-2487).  Returns (new_vis, new_cov)."""
    assert vis.shape[:2] == gains.shape[:2], "vis and gains must have same Npols"
    polmode = '1pol' if tuple(vis.shape[:2]) == (1, 1) else '4pol'
    if cal_2pol and polmode == '4pol':
        polmode = '2pol'
    if undo:
        if polmode in ('1pol', '2pol'):
            if vis_type == 'com':
                inv = torch.zeros_like(gains)
                for p in range(gains.shape[0]):
                    inv[p, p] = 1 / gains[p, p]                     # linalg.diag_inv
                gains = inv
            else:
                gains = -gains
        else:
            assert vis_type == 'com', 'must have complex vis_type for 4pol mode'
            gains = torch.linalg.pinv(gains.permute(2, 3, 4, 0, 1)).permute(3, 4, 0, 1, 2)
    if vis_type == 'dly':
        # float delays: vis + g1 - g2 (calibration.py:2478-2479); three adds, left to torch
        assert polmode in ('1pol', '2pol')
        g1 = gains.index_select(2, g1_idx.to(gains.device))
        g2 = gains.index_select(2, g2_idx.to(gains.device))
        vout = vis + g1 - g2
        if inplace:
            vis.copy_(vout)
            vout = vis
        return vout, cov
    assert vis.is_complex(), "vis_type 'com' needs complex visibilities"
    if polmode == '4pol' and cov is not None:
        raise NotImplementedError("covariance update exists for 1pol / 2pol only, as in the reference")
    vout, cov_out = _ApplyCal.apply(vis, gains, g1_idx, g2_idx, polmode == '4pol', cov)
    if cov is None:
        cov_out = None
    if inplace:
        vis.data.copy_(vout.detach())
        if cov is not None:
            cov.data.copy_(cov_out)
    return vout, cov_out


def apply_cal(vis, bls, gains, ants, cal_2pol=False, cov=None, vis_type='com', undo=False,
              inplace=False):
    """V_out = g_1 V g_2^H (undo: the inverse) for vis (Npol, Npol, Nbls, Ntimes, Nfreqs), gains
    (Npol, Npol, Nants, Ntimes|1, Nfreqs|1), bls antenna pairs (or blnums) along Nbls and ants
    along Nants (calibration.py:2348-2410)."""
    bls = utils.blnum2ants(bls)
    if isinstance(bls, tuple):
        bls = [bls]
    ants = [int(a) for a in ants]
    g1_idx = torch.as_tensor([ants.index(int(bl[0])) for bl in bls], device=gains.device)
    g2_idx = torch.as_tensor([ants.index(int(bl[1])) for bl in bls], device=gains.device)
    return _apply_cal(vis, gains, g1_idx, g2_idx, cal_2pol=cal_2pol, cov=cov, vis_type=vis_type,
                      undo=undo, inplace=inplace)


# ----------------------------------------------------------------------------- JonesModel
GAIN_TYPES = ('com', 'real', 'amp', 'phs', 'amp_phs', 'dly', 'dly_slope', 'phs_slope')


def params2complex(params, param_type):
    """Gain parameters -> complex gains for the representations shared by gains and
    visibilities (calibration.py:215-251); 'amp_phs' keeps (amp, phs) on a trailing axis."""
    if param_type == 'real':
        return params + 0j
    if param_type == 'amp':
        return torch.exp(params) + 0j
    if param_type == 'phs':
        return torch.exp(1j * params)
    if param_type == 'amp_phs':
        return torch.exp(params[..., 0] + 1j * params[..., 1])
    return params


class JonesResponse:
    """Parameter -> gain response of JonesModel (calibration.py:745-875 over BaseResponse :11-148):
    optional LinearModel on the raw parameters, complex view of 2-real 'com' parameters, linear
    frequency / time bases, perturbation about base0, then the gain representation:
    'com' | 'real' | 'amp' (exp) | 'phs' (exp i) | 'amp_phs' | 'dly' (ns; exp(2 pi i nu tau)) |
    'dly_slope' / 'phs_slope' (EW, NS gradients over antpos).  The redcal-degeneracy projection
    block of the reference (setup_projection) is not mirrored."""

    def __init__(self, freq_mode='channel', time_mode='channel', param_type='com', vis_type='com',
                 antpos=None, device=None, freq_LM=None, time_LM=None, freqs=None, times=None,
                 LM=None, base0=None):
        assert param_type in GAIN_TYPES
        assert freq_mode in ('channel', 'linear') and time_mode in ('channel', 'linear')
        self.freq_mode, self.time_mode, self.param_type = freq_mode, time_mode, param_type
        self.vis_type, self.antpos, self.device = vis_type, antpos, device
        self.freq_LM, self.time_LM, self.LM, self.base0 = freq_LM, time_LM, LM, base0
        self.freqs, self.times = freqs, times
        if param_type.endswith('_slope'):
            assert antpos is not None, 'need antpos for dly_slope or phs_slope'
            vec = torch.stack([torch.as_tensor(antpos[a]) for a in antpos]).to(device)
            self.antpos_EW = vec[:, 0][None, None, :, None, None]
            self.antpos_NS = vec[:, 1][None, None, :, None, None]
        if 'dly' in param_type:
            assert freqs is not None, 'need frequencies for delay gain type'
        self._args = dict(freq_mode=freq_mode, time_mode=time_mode, param_type=param_type)

    def _ghz(self, like):
        f = torch.as_tensor(self.freqs).to(like.device)
        return (f / 1e9).to(like.real.dtype if like.is_complex() else like.dtype)

    def params2complex(self, jones):
        jones = params2complex(jones, self.param_type)
        if self.param_type == 'dly' and self.vis_type == 'com':
            jones = torch.exp(2j * math.pi * jones * self._ghz(jones))
        elif self.param_type in ('dly_slope', 'phs_slope'):
            tot = jones[:, :, :1] * self.antpos_EW + jones[:, :, 1:] * self.antpos_NS
            if self.param_type == 'phs_slope':
                jones = torch.exp(1j * tot)
            elif self.vis_type == 'com':
                jones = torch.exp(2j * math.pi * tot * self._ghz(tot))
            else:
                jones = tot
        return jones

    def forward(self, params, **kwargs):
        if not utils.check_devices(params.device, self.device):
            params = params.to(self.device)
        if self.LM is not None:
            params = self.LM(params)
        if self.param_type == 'com' and not torch.is_complex(params):
            params = utils.viewcomp(params)
        if self.freq_mode == 'linear':
            params = self.freq_LM(params)
        if self.time_mode == 'linear':
            params = self.time_LM(params)
        if self.base0 is not None:
            params = params + self.base0
        params = self.params2complex(params)
        if isinstance(params, torch.nn.Parameter):
            params = params.view(params.shape)
        return params

    __call__ = forward

    def push(self, device):
        if not isinstance(device, torch.dtype):
            self.device = device
        for name in ('freq_LM', 'time_LM', 'LM'):
            if getattr(self, name) is not None:
                getattr(self, name).push(device)
        for name in ('base0', 'antpos_EW', 'antpos_NS'):
            if getattr(self, name, None) is not None:
                setattr(self, name, utils.push(getattr(self, name), device))


def rephase_to_refant(params, param_type, refant_idx, p0=None, mode='rephase', inplace=False):
    """Zero the phase of the reference antenna (calibration.py:2490-2608): 'rephase' divides every
    antenna by the refant phasor ('com') or subtracts its delay / phase; 'zero' only clears the
    refant's own imaginary part / phase.  Acts on params and p0 together."""
    if refant_idx is None:
        return None
    if p0 is None:
        p0 = torch.zeros_like(params)
    if not inplace:
        params, p0 = copy.deepcopy(params), copy.deepcopy(p0)
    ref = slice(refant_idx, refant_idx + 1)
    two_real = param_type == 'com' and not torch.is_complex(params)
    if mode == 'rephase':
        if param_type == 'com':
            a = utils.viewcomp(params) if two_real else params
            b = utils.viewcomp(p0) if two_real else p0
            phasor = torch.exp(1j * torch.angle((a + b)[:, :, ref]).detach().clone())
            a, b = a / phasor, b / phasor
            params[:] = utils.viewreal(a) if two_real else a
            p0[:] = utils.viewreal(b) if two_real else b
        elif param_type in ('dly', 'phs'):
            params -= params[:, :, ref].clone()
            p0 -= p0[:, :, ref].clone()
        elif param_type == 'amp_phs':
            params[..., 1] -= params[:, :, ref, ..., 1].clone()
            p0[..., 1] -= p0[:, :, ref, ..., 1].clone()
    elif mode == 'zero':
        for x in (params, p0):
            if param_type == 'com':
                tgt = x[:, :, ref, ..., 1] if two_real else x.imag[:, :, ref]
                tgt.copy_(torch.zeros_like(tgt))
            elif param_type in ('dly', 'phs'):
                x[:, :, ref] = torch.zeros_like(x[:, :, ref])
            elif param_type == 'amp_phs':
                x[:, :, ref, ..., 1] = torch.zeros_like(x[:, :, ref, ..., 1])
    if not inplace:
        return params, p0
    return None


class JonesModel(utils.Module):
    """Antenna-based direction-independent Jones term V^d_pq = J_p V^m_pq J_q^H
    (calibration.py:416-664).  params (Npol, Npol, Nant, Ntimes|Ncoeff, Nfreqs|Ncoeff); polmode
    '1pol' | '2pol' (diagonal) | '4pol'.  forward(vd) returns a new VisData whose data went through
    the CUDA gain product (apply_cal kernels); gradients reach params through JonesResponse by
    autograd and the model visibilities through the product's adjoint."""

    def __init__(self, params, ants, p0=None, refant=None, R=None, parameter=True, polmode='1pol',
                 single_ant=False, name=None, vis_type='com', atol=1e-5):
        super().__init__(name=name)
        self.params = torch.nn.Parameter(params) if parameter else params
        self.device = params.device
        self.p0 = p0
        self.ants = [int(a) for a in ants]
        self.Nants = len(self.ants)
        self.R = R if R is not None else JonesResponse()
        self._times = getattr(self.R, 'times', None)
        self._atol = atol
        self.polmode, self.single_ant, self.vis_type = polmode, single_ant, vis_type
        self.clear_cache()
        self.set_refant(refant)
        self._args = dict(refant=refant, polmode=polmode)
        self._args[self.R.__class__.__name__] = getattr(self.R, '_args', None)

    # ---- caches (IndexCache of the reference, calibration.py:291-413)
    def clear_cache(self):
        self.cache_tidx, self.cache_bidx, self.cache_aidx = {}, {}, {}

    clear_time_cache = clear_bl_cache = clear_ant_cache = clear_cache

    def get_time_idx(self, times):
        """Rows of the gain time axis for a (minibatch of) visibility times, matched to atol."""
        if times is None or self._times is None:
            return None
        key = utils.arr_hash(times)
        if key not in self.cache_tidx:
            ref = torch.as_tensor(self._times, dtype=torch.float64).cpu()
            idx = [int(torch.where(torch.isclose(ref, torch.as_tensor(float(t), dtype=torch.float64),
                                                 atol=self._atol, rtol=1e-15))[0][0])
                   for t in torch.as_tensor(times).cpu()]
            self.cache_tidx[key] = idx
        return self.cache_tidx[key]

    def index_params(self, jones, times=None):
        idx = self.get_time_idx(times)
        if idx is None or (len(idx) == jones.shape[-2] and idx == list(range(len(idx)))):
            return jones
        return jones[..., torch.as_tensor(idx, device=jones.device), :]

    def get_ant_idx(self, bls):
        """(g1_idx, g2_idx): row of the Nants axis for the first / second antenna of every
        baseline (blnums or antenna pairs); cached per baseline set (calibration.py:529-563)."""
        key = utils.arr_hash(bls)
        if key not in self.cache_aidx:
            pairs = utils.blnum2ants(bls.cpu().numpy() if isinstance(bls, torch.Tensor) else bls)
            if self.single_ant:
                i1 = i2 = [0] * len(pairs)
            else:
                row = {a: k for k, a in enumerate(self.ants)}
                i1 = [row[int(b[0])] for b in pairs]
                i2 = [row[int(b[1])] for b in pairs]
            self.cache_aidx[key] = (torch.as_tensor(i1, device=self.device),
                                    torch.as_tensor(i2, device=self.device))
        return self.cache_aidx[key]

    # ---- reference antenna (calibration.py:565-597)
    def set_refant(self, refant):
        self.refant, self.refant_idx, self.rephase_mode = refant, None, None
        if refant is not None:
            assert refant in self.ants, "need a valid refant"
            self.refant_idx = self.ants.index(refant)
            chan = self.R.time_mode == 'channel' and self.R.freq_mode == 'channel'
            self.rephase_mode = 'rephase' if chan else 'zero'
            self.fix_refant_phs()

    def fix_refant_phs(self):
        with torch.no_grad():
            rephase_to_refant(self.params, self.R.param_type, self.refant_idx, p0=self.p0,
                              mode=self.rephase_mode, inplace=True)

    def forward(self, vd, undo=False, prior_cache=None, jones=None):
        """vd: VisData (Npol, Npol, Nbl, Ntimes, Nfreqs) of model visibilities -> VisData of
        predicted data (undo: calibrated data from raw)."""
        if self.refant_idx is not None:
            self.fix_refant_phs()
        params = self.params if self.p0 is None else self.params + self.p0
        if jones is None:
            jones = self.R(params)
        hooks = getattr(self, '_hook_registry', None)
        if hooks is not None and jones.requires_grad:
            for h in hooks:
                jones.register_hook(h)
        self.eval_prior(prior_cache, inp_params=self.params, out_params=jones)
        jones = self.index_params(jones, times=vd.times)
        g1_idx, g2_idx = self.get_ant_idx(vd._blnums)
        data, _ = _apply_cal(vd.data, jones.to(vd.data.device), g1_idx, g2_idx,
                             cal_2pol=self.polmode == '2pol', vis_type=self.vis_type, undo=undo)
        vout = dataset_shell(vd)
        vout.data = data
        return vout

    def push(self, device):
        if not isinstance(device, torch.dtype):
            self.clear_cache()
            self.device = device
        self.params = utils.push(self.params, device)
        self.R.push(device)
        if self.p0 is not None:
            self.p0 = utils.push(self.p0, device)


def dataset_shell(vd):
    """New VisData sharing vd's metadata, flags and covariance (not its data)."""
    out = copy.copy(vd)
    out.data = None
    return out
