"""
TEST INFRASTRUCTURE (like everything under oracle/): fp64 oracle visibilities of a BASELINE
workload model (workloads.py) on a subset of baselines / channels / times, evaluated with the
restated reference formulation of oracle/rime_oracle.py from the model's own parameter values.

Used by tests/test_gpu_parity.py, __graft_entry__.smoke() and the post-timing `parity_check`
of bench.py -- as the checker, never as the thing measured.
"""
import torch

from . import rime_oracle as orc


def oracle_vis_subset(rime, zenaz, bl_sel, f_idx=None, kind=None):
    """V (Npol, Npol, len(bl_sel), Nt, len(f_idx)) complex128 of `rime` (a bayeslim_b200.RIME
    built by workloads.py) for the current time group.

    zenaz : list over the times of (zen, az) [deg] tensors of all sky pixels / sources
    bl_sel: indices into rime.sim_bls;  f_idx: channel indices (None = all)
    kind  : 'airy' | 'interp' | 'interp_pol' (None: from the beam object)"""
    zenaz = [(torch.as_tensor(z).detach().cpu().double(), torch.as_tensor(a).detach().cpu().double())
             for z, a in zenaz]
    allf = rime.array.freqs.detach().cpu().double()
    f_idx = torch.arange(len(allf)) if f_idx is None else torch.as_tensor(f_idx)
    freqs = allf[f_idx]
    bls = [rime.sim_bls[i] for i in bl_sel]
    blvecs = rime.sim_blvecs.detach().cpu().double()[list(bl_sel)]
    beam = rime.beam
    if kind is None:
        if beam.R.__class__.__name__ == 'AiryResponse':
            kind = 'airy'
        else:
            kind = 'interp' if beam.powerbeam else 'interp_pol'
    with torch.no_grad():
        sky = rime.sky.forward().data.detach().cpu().double()[:, :, f_idx]
        if kind == 'airy':
            p = beam.params.detach().cpu().double()

            def beam_fn(z, a):
                return orc.airy_response(p, z, a, freqs, powerbeam=True)
            powerbeam = True
        else:
            bmap = beam.params.detach().cpu().double()[:, :, :, f_idx]
            if kind == 'interp':
                bmap = bmap.abs()
            tg, pg = beam.R.theta_grid.cpu().double(), beam.R.phi_grid.cpu().double()

            def beam_fn(z, a):
                inds, wgts = orc.rect_interp_weights(tg, pg, z, a, 'linear')
                return orc.interp_map(bmap, inds, wgts)
            powerbeam = kind == 'interp'
        return orc.rime_forward(sky, zenaz, beam_fn, bls, blvecs, freqs, fov=beam.fov,
                                powerbeam=powerbeam, bl_chunk=4)
