"""
Gain application, mirroring the reference's ``calibration.apply_cal`` / ``_apply_cal``
(bayeslim/calibration.py:2348-2487): V_out = g_1 V g_2^H per baseline, the step that follows
the RIME in a BayesLIM ``Sequential``.

The product and its adjoints (to the visibilities and to the gains) run in
``csrc/cal_kernels.cu``; the reference builds g_1, g_2 and G = g_1 conj(g_2) as three
visibility-sized temporaries with ``index_select`` (calibration.py:2462-2468).  Inverting the
gains for ``undo`` stays in torch: it acts on the small (Nants, Ntimes, Nfreqs) table.
``JonesModel`` (parameter -> gain response functions, priors, refant handling) is not mirrored.
"""
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
