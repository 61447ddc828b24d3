"""
B200-native drop-in for the reference's ``rime_model.RIME`` (bayeslim/rime_model.py:13-482).

Same constructor, attributes, minibatch API and ``VisData`` output; the arithmetic of
``forward`` -- beam evaluation at the source directions, FOV cut, beam x sky product, fringe
generation and the sum over sources (rime_model.py:326-440) -- runs in the sm_100a CUDA
kernels of libb200rime.so through ``ops``:

    reference (per time, torch)                         here (per time group, CUDA)
    ------------------------------------------------    --------------------------------------
    beam.gen_beam -> R(params, zen, az)   :354          ops.build_airy / ops.build_interp
    cut_sky_fov(sky, cut)                 :360            (fused: beam, cut gather, beam*sky,
    beam.apply_beam(beam, bls, cut_sky)   :423             write tiled layout)   or, for other
                                                           responses, torch + ops.pack_planes
    array.gen_fringe(blvecs, zen, az)     :426          ops.fringe_sum (fringe generated on
    torch.sum(fringe * psky, -1)          :429             the fly, never materialised)
    torch.stack(skyvis, dim=3)            :368          written in place into (.., Nbl, Nt, Nf)

Gradients reach sky.params, beam.params and array.antvecs through torch.autograd Functions
whose backward passes are CUDA kernels as well.  There is no CPU path: tensors must be on a
CUDA device, otherwise forward() raises.
"""
from datetime import datetime
import os

import numpy as np
import torch

from . import utils, dataset, beam_model, ops
from .dataset import VisData


# ---- beam accessors written against the REFERENCE PixelBeam attribute set (params, p0, fov,
# ant2beam, skycut_cache/query_cache/set_skycut_cache), so that reference beam objects can be
# handed to this RIME unchanged
def _beam_params(beam):
    return beam.params if getattr(beam, 'p0', None) is None else beam.params + beam.p0


def _beam_sky_cut(beam, zen):
    """FOV cut of beam_model.py:221-224 (strict zen < fov/2; everything for fov >= 360)."""
    cached = beam.query_cache(zen) if getattr(beam, 'skycut_cache', False) else None
    if cached is not None:
        return cached
    cut = torch.where(zen < beam.fov / 2)[0] if beam.fov < 360 else slice(None)
    if getattr(beam, 'skycut_cache', False):
        beam.set_skycut_cache(zen, cut, device=getattr(beam, 'skycut_device', None))
    return cut


def _beam_model_pairs(beam, bls):
    """Sorted unique (model_i, model_j) pairs and the pair index of every baseline
    (beam_model.py:303-305, 366-367)."""
    pairs = [(beam.ant2beam[b[0]], beam.ant2beam[b[1]]) for b in bls]
    uniq = sorted(set(pairs))
    lookup = {mp: i for i, mp in enumerate(uniq)}
    return uniq, [lookup[p] for p in pairs]


class _GeometryRecord:
    """Per (sky component, time group) device tables, built once and reused by every forward."""

    def __init__(self):
        self.src_ids = None
        self.geom = None
        self.cuts = None      # per-time int64 index tensors (or None when the FOV keeps all)
        self.zen = None       # per-time cut zen/az [deg], float64, device
        self.az = None
        self.airy = {}        # dtype -> ops.AiryTable
        self.interp = {}      # (id(R), interp_mode, dtype) -> ops.InterpTable


def _is_interp_response(R):
    """Responses whose beam is a pixel map (beam_cache, set once per forward) interpolated at the
    source directions: PixelResponse, and YlmResponse in 'interpolate' mode (its map comes from
    the CUDA spherical-harmonic product)."""
    name = R.__class__.__name__
    if getattr(R, 'Rchi', None) is not None:
        return False
    return name == 'PixelResponse' or (name == 'YlmResponse' and getattr(R, 'mode', None) == 'interpolate')


class RIME(utils.Module):
    """Radio interferometric measurement equation,
    V_pq(t, nu) = sum_s A_p(s, nu) I(s, nu) A_q(s, nu)^H exp(2 pi i b_pq . s nu / c)."""

    def __init__(self, sky, telescope, beam, array, sim_bls, times, freqs, data_bls=None,
                 device=None, cache_eq2top=True, name=None, verbose=False):
        super().__init__(name=name)
        self.sky = sky
        self.telescope = telescope
        self.beam = beam
        self.array = array
        self.device = device
        self.cache_eq2top = cache_eq2top
        self.verbose = verbose
        self._geom_cache = {}
        self._freq_cache = {}
        self.setup_freqs(freqs)
        self.setup_sim_bls(sim_bls, data_bls)
        self.setup_sim_times(times=times)

    # ------------------------------------------------------------------ bookkeeping
    def push(self, device):
        dtype = isinstance(device, torch.dtype)
        self.sim_blvec_groups = {k: v.to(device) for k, v in self.sim_blvec_groups.items()}
        if not dtype:
            self.device = device
            for k, v in self._sim2data.items():
                if v is not None:
                    self._sim2data[k] = utils.push(v, device)
            self.clear_geometry_cache()

    def clear_geometry_cache(self):
        self._geom_cache = {}
        self._freq_cache = {}

    @property
    def Ntimes_all(self):
        return len(self.all_sim_times)

    @property
    def Nbls_all(self):
        return len(self.all_sim_bls)

    def setup_freqs(self, freqs):
        self.freqs = freqs
        self.Nfreqs = len(freqs)

    def setup_sim_bls(self, sim_bls, data_bls=None):
        """Group the simulated baselines and build the sim->data redundancy index
        (rime_model.py:148-226)."""
        self.bl_group_id = 0
        if not isinstance(sim_bls, dict):
            ints = (int, np.integer)
            assert isinstance(sim_bls[0][0], ints) or isinstance(sim_bls[0][0][0], ints), \
                "sim_bls must be list of 2-tuples or list of list of 2-tuples"
            if isinstance(sim_bls[0], tuple) or isinstance(sim_bls[0][0], ints):
                sim_bl_groups = {0: list(sim_bls)}
            else:
                sim_bl_groups = {i: list(g) for i, g in enumerate(sim_bls)}
        else:
            sim_bl_groups = {k: list(v) for k, v in sim_bls.items()}
        for k in sim_bl_groups:
            sim_bl_groups[k] = [tuple(int(a) for a in bl) for bl in sim_bl_groups[k]]
        if data_bls is not None:
            data_bls = [tuple(int(a) for a in bl) for bl in data_bls]
        self.sim_bl_groups = sim_bl_groups
        self.all_sim_bls = [bl for g in sim_bl_groups.values() for bl in g]
        self.Nbl_groups = len(sim_bl_groups)
        # detached: a parameter antvecs is followed through _baseline_meta at every forward; a
        # cached tensor with a graph would pin antvecs' gradient accumulator to the stream of
        # this call (which breaks CUDA-graph capture of the backward)
        self.sim_blvec_groups = {k: self.array.get_blvecs(v).detach()
                                 for k, v in sim_bl_groups.items()}
        self._bl_meta = {}
        self._ant_tilings = {}
        self._tc_tilings = {}
        self._blnum_cache = {}
        if data_bls is None:
            self.data_bl_groups = self.sim_bl_groups
            self._sim2data = {k: None for k in sim_bl_groups}
        else:
            self._sim2data, self.data_bl_groups = {}, {}
            bl2red = self.array.bl2red
            for k, grp in sim_bl_groups.items():
                sim_red = [bl2red[bl] for bl in grp]
                keep = set(sim_red)
                dbls = [bl for bl in data_bls if bl2red[bl] in keep]
                data_red = [bl2red[bl] for bl in dbls]
                assert set(sim_red) == set(data_red), \
                    "non-overlapping bl type(s) in data_bls and sim_bls"
                assert len(np.where(np.diff(data_red) != 0)[0]) == len(grp) - 1
                lookup = {r: i for i, r in enumerate(sim_red)}
                self.data_bl_groups[k] = dbls
                self._sim2data[k] = torch.as_tensor([lookup[r] for r in data_red],
                                                    device=self.device)
        self._set_group()

    def setup_sim_times(self, times):
        self.time_group_id = 0
        if not isinstance(times, dict):
            if isinstance(times, list) or (isinstance(times, (np.ndarray, torch.Tensor))
                                           and times.ndim > 1):
                times = {k: np.asarray(t) for k, t in enumerate(times)}
            else:
                times = {0: np.asarray(times)}
        self.sim_time_groups = times
        self.all_sim_times = np.asarray([t for g in times.values() for t in np.atleast_1d(g)])
        self.Ntime_groups = len(times)
        self._set_group()

    @property
    def Nbatch(self):
        if hasattr(self, 'sim_bl_groups') and hasattr(self, 'sim_time_groups'):
            return len(self.sim_bl_groups) * len(self.sim_time_groups)
        return None

    @property
    def batch_idx(self):
        if hasattr(self, 'bl_group_id') and hasattr(self, 'time_group_id'):
            return self.time_group_id * len(self.sim_bl_groups) + self.bl_group_id
        return None

    @batch_idx.setter
    def batch_idx(self, val):
        assert 0 <= val < self.Nbatch
        self.time_group_id = int(val // len(self.sim_bl_groups))
        self.bl_group_id = int(val % len(self.sim_bl_groups))
        self._set_group()

    def _set_group(self):
        if hasattr(self, 'sim_bl_groups'):
            keys = list(self.sim_bl_groups.keys())
            k = keys[self.bl_group_id] if self.bl_group_id not in self.sim_bl_groups \
                else self.bl_group_id
            self._bl_key = k
            self.sim_bls = self.sim_bl_groups[k]
            self.sim_blvecs = self.sim_blvec_groups[k]
            self.Nsim_bls = len(self.sim_bls)
            self.data_bls = self.data_bl_groups[k]
            self.Ndata_bls = len(self.data_bls)
        if hasattr(self, 'sim_time_groups'):
            keys = list(self.sim_time_groups.keys())
            k = keys[self.time_group_id] if self.time_group_id not in self.sim_time_groups \
                else self.time_group_id
            self.sim_times = self.sim_time_groups[k]
            self.Ntimes = len(self.sim_times)

    # ------------------------------------------------------------------ device state
    def _compute_device(self, sky_data):
        dev = self.device if self.device is not None else sky_data.device
        dev = torch.device(dev)
        if dev.type != 'cuda':
            raise RuntimeError(
                "bayeslim_b200.RIME runs on CUDA devices only (got %s): the B200 kernels are the "
                "implementation, there is no CPU path. Push the model to 'cuda'." % dev)
        return dev if dev.index is not None else torch.device('cuda', torch.cuda.current_device())

    def _freqs64(self, obj, dev):
        """float64 device copy of obj.freqs (with its optional _freq_idx), cached by identity."""
        f = obj.freqs
        idx = getattr(obj, '_freq_idx', None)
        key = (id(f), str(idx), dev)
        if key not in self._freq_cache:
            ff = torch.as_tensor(f)
            if idx is not None:
                ff = ff[idx]
            self._freq_cache[key] = (f, ff.detach().to(device=dev, dtype=torch.float64).contiguous())
        return self._freq_cache[key][1]

    def _baseline_meta(self, dev, dtype):
        """(blvecs on the compute device, uniform-frequency flag) for the current group."""
        if getattr(self.array.antvecs, 'requires_grad', False):
            # follow the live parameter (fresh graph every forward); the antenna-row tables of
            # the group are uploaded once (no host-to-device copy per step: capturable)
            av = self.array.antvecs
            ikey = (self._bl_key, str(av.device), 'rows')
            if ikey not in self._bl_meta:
                i, j = self._antenna_rows()
                self._bl_meta[ikey] = (torch.as_tensor(i, dtype=torch.long, device=av.device),
                                       torch.as_tensor(j, dtype=torch.long, device=av.device))
            i0, i1 = self._bl_meta[ikey]
            blvecs = av.index_select(0, i1) - av.index_select(0, i0)
        else:
            blvecs = self.sim_blvecs
        blvecs = blvecs.to(dev)
        key = (self._bl_key, dev, dtype)
        if key not in self._bl_meta:
            blmax = float(blvecs.detach().double().norm(dim=1).max()) if len(blvecs) else 0.0
            uniform = ops.freqs_uniform(self._freqs64(self.array, dev), blmax, dtype)
            self._bl_meta[key] = uniform
        return blvecs, self._bl_meta[key]

    def _ant_tiling(self, dev):
        """ops.AntTiling of the current baseline group when the antenna-factorised float32
        kernels pay off (the group covers most pairs of its antennas), else None.
        B200RIME_ANT=0 forces the baseline-owned kernels."""
        if os.environ.get("B200RIME_ANT", "1") == "0":
            return None
        key = (self._bl_key, dev)
        if key not in self._ant_tilings:
            til = None
            try:
                i, j = self._antenna_rows()
                til = ops.AntTiling(i, j, len(self.array.antvecs), dev)
                if not til.usable:
                    til = None
            except (KeyError, TypeError, AttributeError, IndexError):
                til = None
            self._ant_tilings[key] = til
        return self._ant_tilings[key]

    def _antenna_rows(self):
        """Antenna row (into array.antvecs) of the first / second antenna of every baseline."""
        rows = getattr(self.array, '_ant_idx', None)
        if rows is None:
            rows = {int(a): k for k, a in enumerate(self.array.ants)}
        bls = self.sim_bls
        return [rows[int(b[0])] for b in bls], [rows[int(b[1])] for b in bls]

    def _tc_tiling(self, dev):
        """ops.TcTiling of the current baseline group when the tensor-core kernels pay off (the
        group fills its 128 x 128 antenna-pair items well enough), else None.  B200RIME_TC=0
        keeps the FP32-pipe kernels."""
        if os.environ.get("B200RIME_TC", "1") == "0" or os.environ.get("B200RIME_ANT", "1") == "0":
            return None
        key = (self._bl_key, dev)
        if key not in self._tc_tilings:
            tc = None
            try:
                i, j = self._antenna_rows()
                tc = ops.TcTiling(i, j, len(self.array.antvecs), dev)
                if not (tc.usable and tc.bwd_usable):
                    tc = None
            except (KeyError, TypeError, AttributeError, IndexError):
                tc = None
            self._tc_tilings[key] = tc
        return self._tc_tilings[key]

    def _fringe(self, A, blvecs, rec, f64, nfreq, uniform, dev):
        """Fringe sum of tiled planes A over all baselines of the current group: tensor-core
        kernels when the group is dense in antenna pairs, else the FP32 antenna-factorised
        kernels, else the baseline-owned kernels (always for float64)."""
        if A.dtype == torch.float32:
            tc = self._tc_tiling(dev)
            til = self._ant_tiling(dev)
            if tc is not None or til is not None:
                return ops.fringe_sum_ant(A, self.array.antvecs.to(dev), til, rec.geom, f64, nfreq,
                                          conj=False, tc=tc)
        return ops.fringe_sum(A, blvecs, rec.geom, f64, nfreq, conj=False, uniform=uniform)

    def _geometry(self, sky_comp, dev):
        """Per-time zen/az -> FOV cut -> packed source axis; cached per (component, time group)."""
        ra, dec = sky_comp.angs
        key = (sky_comp.name, len(ra), tuple(float(t) for t in self.sim_times), float(self.beam.fov),
               str(dev))
        zenaz, ids = [], []
        for time in self.sim_times:
            tkey = (sky_comp.name, len(ra), time)
            za = self.telescope.eq2top(time, ra, dec, store=self.cache_eq2top, key=tkey)
            zenaz.append(za)
            ids.append(id(za))
        rec = self._geom_cache.get(key)
        if rec is not None and self.cache_eq2top and rec.src_ids == tuple(ids):
            return rec
        rec = _GeometryRecord()
        rec.src_ids = tuple(ids)
        rec.cuts, rec.zen, rec.az = [], [], []
        for time, za in zip(self.sim_times, zenaz):
            zen = torch.as_tensor(za[0]).to(dev, torch.float64)
            az = torch.as_tensor(za[1]).to(dev, torch.float64)
            tkey = (sky_comp.name, len(ra), time)
            zen._arr_hash = tkey
            cut = _beam_sky_cut(self.beam, zen)
            if isinstance(cut, slice):
                cut = torch.arange(len(zen), device=dev)
            cut = cut.to(dev)
            zc, ac = zen[cut], az[cut]
            zc._arr_hash = tkey
            rec.cuts.append(cut)
            rec.zen.append(zc)
            rec.az.append(ac)
        rec.geom = ops.Geometry(rec.zen, rec.az, dev)
        self._geom_cache[key] = rec
        return rec

    # ------------------------------------------------------------------ perceived sky
    def _fused_mode(self, sky):
        """'airy' | 'interp' when the beam/sky combination has a fused CUDA builder, else None."""
        b, R = self.beam, self.beam.R
        if not (b.powerbeam and b.Nvec == 1 and b.Nmodel == 1 and sky.shape[:2] == (1, 1)):
            return None
        if sky.is_complex():
            return None
        offset = getattr(b, 'theta_x', 0) > 0 or getattr(b, 'theta_y', 0) > 0
        rname = R.__class__.__name__
        if rname == 'AiryResponse' and not getattr(R, 'brute_force', False) \
                and getattr(R, 'taper_kwargs', None) is None and not offset:
            return 'airy'
        if _is_interp_response(R):
            return 'interp'
        return None

    def _interp_pol_mode(self, sky):
        """True when the beam is an interpolated pixel beam in a polarised / voltage / multi-model
        mode: Jones and coherency planes are then built by the CUDA interpolation / gather
        kernels directly in the tiled layout and combined element-wise."""
        b, R = self.beam, self.beam.R
        return _is_interp_response(R)

    def _build_airy(self, sky, rec, dev):
        b, R = self.beam, self.beam.R
        dtype = sky.dtype
        if dtype not in rec.airy:
            rec.airy[dtype] = ops.AiryTable(rec.geom, rec.cuts, rec.zen, rec.az, sky.shape[-1], dtype)
        p = _beam_params(b).to(dev)
        f64 = self._freqs64(b, dev)
        planes = [ops.build_airy(sky[0, 0], p[ipol, 0, 0, 0], rec.geom, rec.airy[dtype], f64,
                                 freq_ratio=R.freq_ratio, square=True,
                                 full_grad=getattr(R, 'full_grad', False))
                  for ipol in range(b.Npol)]
        return planes[0] if len(planes) == 1 else torch.cat(planes, dim=0)

    def _build_interp(self, sky, rec, dev):
        b, R = self.beam, self.beam.R
        dtype = sky.dtype
        if R.beam_cache is None:
            R.set_beam_cache(_beam_params(b))
        self._register_beam_hooks()
        bmap = R.beam_cache.to(dev)
        tab = self._interp_table(sky, rec, bmap, dtype)
        planes = [ops.build_interp(sky[0, 0], bmap[ipol, 0, 0], rec.geom, tab)
                  for ipol in range(b.Npol)]
        return planes[0] if len(planes) == 1 else torch.cat(planes, dim=0)

    def _generic_planes(self, sky, rec, dev):
        """Any response function: beam and beam*sky*beam^H in torch on the device (cheap: not
        multiplied by Nbl), then one real plane per (pol product, model pair, re|im)."""
        b = self.beam
        p = _beam_params(b)
        modelpairs, mp_idx = _beam_model_pairs(b, self.sim_bls)
        per_time = []
        for cut, zen, az in zip(rec.cuts, rec.zen, rec.az):
            # the response is evaluated at the pointing-offset directions, the fringe keeps the
            # true ones (beam_model.py:244-259)
            bzen, baz = beam_model.offset_zen_az(b, zen, az)
            beam = b.R(p, bzen, baz, b.freqs)
            self._register_beam_hooks()
            beam = torch.as_tensor(beam).to(dev)
            cut_sky = sky.index_select(-1, cut)
            per_time.append(beam_model.perceived_sky(beam, cut_sky, modelpairs, b.Npol, b.Nvec,
                                                     b.powerbeam))
        return per_time, modelpairs, mp_idx

    def _register_beam_hooks(self):
        """Gradient hooks of the beam on R.beam_cache (beam_model.py:262-266)."""
        hooks = getattr(self.beam, '_hook_registry', None)
        bc = getattr(self.beam.R, 'beam_cache', None)
        if hooks is not None and bc is not None and bc.requires_grad:
            for hook in hooks:
                bc.register_hook(hook)

    def _interp_table(self, sky, rec, bmap, dtype):
        R, b = self.beam.R, self.beam
        key = (id(R), R.interp_mode, dtype, float(getattr(b, 'theta_x', 0)),
               float(getattr(b, 'theta_y', 0)))
        if key not in rec.interp:
            # interpolation weights at the pointing-offset directions (beam_model.py:244-259)
            tabs = [R.get_interp(*beam_model.offset_zen_az(b, z, a))
                    for z, a in zip(rec.zen, rec.az)]
            rec.interp[key] = ops.InterpTable(rec.geom, rec.cuts, [t[0] for t in tabs],
                                              [t[1] for t in tabs], sky.shape[-1], bmap.shape[-1],
                                              dtype)
        return rec.interp[key]

    def _tiled_pol_planes(self, sky, rec, dev):
        """Perceived sky of the polarised modes (beam_model.py:343-363) in the tiled layout:
        every Jones element J[a, b, model] is interpolated and every coherency element C[b, c]
        gathered through the FOV cut by the CUDA builder (one launch each, all times), then
        P = J C J^H is one fused CUDA pass over the planes for real 4-pol operands
        (ops.jones_sandwich) and an element-wise torch expression otherwise.  Nothing of
        shape (.., Nf, Ns, Nneighbours) is ever formed (the reference's gather temp is 25 GB per
        time at nside 256)."""
        b, R = self.beam, self.beam.R
        rdtype = ops._real(sky.dtype)
        if R.beam_cache is None:
            R.set_beam_cache(_beam_params(b))
        self._register_beam_hooks()
        bmap = R.beam_cache.to(dev)
        tab = self._interp_table(sky, rec, bmap, rdtype)

        def tiled(x2d, is_beam):
            if x2d.is_complex():
                return torch.complex(tiled(x2d.real, is_beam), tiled(x2d.imag, is_beam))
            x2d = x2d.to(rdtype)
            A = ops.build_interp(None, x2d, rec.geom, tab) if is_beam \
                else ops.build_interp(x2d, None, rec.geom, tab)
            return A[0]

        modelpairs, mp_idx = _beam_model_pairs(b, self.sim_bls)
        Jp = [[[tiled(bmap[a, v, m], True) for v in range(bmap.shape[1])]
               for a in range(bmap.shape[0])] for m in range(bmap.shape[2])]     # [model][pol][vec]
        Cp = [[tiled(sky[v, w], False) for w in range(sky.shape[1])] for v in range(sky.shape[0])]
        real = not (bmap.is_complex() or sky.is_complex())
        if real and not b.powerbeam and tuple(bmap.shape[:2]) == (2, 2) and tuple(sky.shape[:2]) == (2, 2):
            # 4-pol real Jones x real coherency: P = J1 C J2^T in one fused CUDA pass per model
            # pair (the reference's einsum, beam_model.py:363) -- no stacked operands
            per_pair = [ops.jones_sandwich(Jp[m1], Jp[m1] if m2 == m1 else Jp[m2], Cp)
                        for (m1, m2) in modelpairs]
            psky = per_pair[0].unsqueeze(2) if len(per_pair) == 1 else torch.stack(per_pair, dim=2)
            return psky, modelpairs, mp_idx
        J = torch.stack([torch.stack([torch.stack([Jp[m][a][v] for m in range(bmap.shape[2])])
                                      for v in range(bmap.shape[1])])
                         for a in range(bmap.shape[0])])            # (Npol, Nvec, Nmodel, *tile)
        C = torch.stack([torch.stack(Cp[v]) for v in range(sky.shape[0])])       # (Nvec, Nvec, *tile)
        psky = beam_model.perceived_sky(J, C, modelpairs, b.Npol, b.Nvec, b.powerbeam)
        return psky, modelpairs, mp_idx

    # ------------------------------------------------------------------ forward
    def forward(self, *args, prior_cache=None, **kwargs):
        """Simulate the visibilities of the current (time group, baseline group) batch.
        Returns VisData with data of shape (Npol, Npol, Ndata_bls, Ntimes, Nfreqs)."""
        self._set_group()
        sky_components = self.sky.forward(prior_cache=prior_cache)
        if not isinstance(sky_components, list):
            sky_components = [sky_components]
        Npol = self.beam.Npol
        pol = "{}{}".format(self.beam.pol, self.beam.pol) if Npol == 1 else None
        if hasattr(self.beam.R, 'clear_beam_cache'):
            self.beam.R.clear_beam_cache()
        self.beam.skycut_device = getattr(self.sky, 'device', None)

        start = datetime.now().timestamp()
        vis = None
        for i, sky_comp in enumerate(sky_components):
            sky = sky_comp.data
            dev = self._compute_device(sky)
            sky = sky.to(dev)
            rdtype = ops._real(sky.dtype)
            rec = self._geometry(sky_comp, dev)
            blvecs, uniform = self._baseline_meta(dev, rdtype)
            f64 = self._freqs64(self.array, dev)
            nfreq = len(f64)
            mode = self._fused_mode(sky)
            if mode is not None:
                A = self._build_airy(sky, rec, dev) if mode == 'airy' \
                    else self._build_interp(sky, rec, dev)
                V = self._fringe(A, blvecs, rec, f64, nfreq, uniform, dev)
                skyvis = V[:, None]                     # (Npol, 1, Nbl, Nt, Nf)
            elif self._interp_pol_mode(sky):
                psky, modelpairs, mp_idx = self._tiled_pol_planes(sky, rec, dev)
                skyvis = self._sum_model_pairs(psky, None, modelpairs, mp_idx, sky, rec, dev,
                                               blvecs, f64, nfreq, uniform)
            else:
                per_time, modelpairs, mp_idx = self._generic_planes(sky, rec, dev)
                skyvis = self._sum_model_pairs(None, per_time, modelpairs, mp_idx, sky, rec, dev,
                                               blvecs, f64, nfreq, uniform)
            if self.verbose:
                log("sky model {}/{} | {} elapsed".format(i + 1, len(sky_components),
                                                          elapsed_time(start)), verbose=True)
            vis = skyvis if vis is None else vis + skyvis
        self.beam.eval_prior(prior_cache)

        sim2data = self._sim2data[self._bl_key]
        if sim2data is not None:
            vis = torch.index_select(vis, 2, sim2data.to(vis.device))

        vd = VisData()
        tkw = dict(tloc=getattr(self.telescope, 'tloc', None), device=self.telescope.device)
        if hasattr(self.telescope, 'eq2top_fn'):      # this package's TelescopeModel
            tkw.update(dtype=getattr(self.telescope, 'dtype', None),
                       eq2top_fn=self.telescope.eq2top_fn)
        telescope = self.telescope.__class__(self.telescope.location, **tkw)
        vd.setup_meta(telescope, self.array.to_antpos())
        bkey = (self._bl_key, str(vis.device))
        if bkey not in self._blnum_cache:
            nums = np.asarray(utils.ants2blnum(list(self.data_bls)))
            self._blnum_cache[bkey] = (torch.as_tensor(nums, device=vis.device), nums)
        vd.setup_data(self.data_bls, self.sim_times, self.freqs, pol=pol, data=vis, flags=None,
                      cov=None, history=self._history(), _blnums=self._blnum_cache[bkey])
        return vd

    def _sum_model_pairs(self, tiled, per_time, modelpairs, mp_idx, sky, rec, dev, blvecs, f64,
                         nfreq, uniform):
        """Fringe-sum the perceived sky of every model pair over its own baselines.  The
        perceived sky comes either already tiled, (P, Q, Nmp, nchunk, S, KC), or as per-time
        row-major tensors (P, Q, Nmp, Nf, Ns_t) that are packed first.  Complex perceived skies
        are split into real planes (V = V_re + i V_im)."""
        ref = tiled if tiled is not None else per_time[0]
        P, Q = ref.shape[0], ref.shape[1]
        cplx = ref.is_complex()
        # kernel precision follows the sky tensor (float64 frequency/angle inputs of a response
        # function must not silently promote a float32 session)
        rdtype = ops._real(sky.dtype)
        mp_idx_t = torch.as_tensor(mp_idx, device=dev)
        out = None
        for m in range(len(modelpairs)):
            sel = torch.where(mp_idx_t == m)[0]
            if tiled is not None:
                Xm = tiled[:, :, m].reshape((P * Q,) + tuple(tiled.shape[3:]))
                A = (torch.cat([Xm.real, Xm.imag], dim=0) if cplx else Xm).to(rdtype).contiguous()
            else:
                planes = []
                for X in per_time:
                    Xm = X[:, :, m].reshape(P * Q, X.shape[-2], X.shape[-1])
                    planes.append(torch.cat([Xm.real, Xm.imag], dim=0) if cplx else Xm)
                A = ops.pack_planes(rec.geom, [pl.to(rdtype).contiguous() for pl in planes])
            if len(modelpairs) == 1:
                V = self._fringe(A, blvecs, rec, f64, nfreq, uniform, dev)
            else:
                V = ops.fringe_sum(A, blvecs.index_select(0, sel), rec.geom, f64, nfreq,
                                   conj=False, uniform=uniform)
            if cplx:
                V = V[:P * Q] + 1j * V[P * Q:]
            V = V.reshape(P, Q, len(sel), V.shape[-2], V.shape[-1])
            if len(modelpairs) == 1:
                return V
            if out is None:
                out = torch.zeros(P, Q, len(mp_idx), V.shape[-2], V.shape[-1], dtype=V.dtype,
                                  device=dev)
            out = out.index_copy(2, sel, V)
        return out

    # ------------------------------------------------------------------ fused likelihood
    def _single_plane_set(self, sky_comp):
        """(A, (P, Q), rec, blvecs, f64, nfreq, uniform, dev) when the perceived sky of this
        component is one set of real tiled planes summed over all baselines of the group (fused
        builders, or polarised interpolated beams with a single beam model), else None."""
        sky = sky_comp.data
        dev = self._compute_device(sky)
        sky = sky.to(dev)
        rdtype = ops._real(sky.dtype)
        rec = self._geometry(sky_comp, dev)
        blvecs, uniform = self._baseline_meta(dev, rdtype)
        f64 = self._freqs64(self.array, dev)
        mode = self._fused_mode(sky)
        if mode is not None:
            A = self._build_airy(sky, rec, dev) if mode == 'airy' else self._build_interp(sky, rec, dev)
            return A, (A.shape[0], 1), rec, blvecs, f64, len(f64), uniform, dev
        if self._interp_pol_mode(sky):
            psky, modelpairs, _ = self._tiled_pol_planes(sky, rec, dev)
            if len(modelpairs) == 1 and not psky.is_complex():
                P, Q = psky.shape[0], psky.shape[1]
                A = psky[:, :, 0].reshape((P * Q,) + tuple(psky.shape[3:])).to(rdtype).contiguous()
                return A, (P, Q), rec, blvecs, f64, len(f64), uniform, dev
        return None

    def forward_chisq(self, data, icov=None, prior_cache=None):
        """chisq = sum icov |V - data|^2 of the current batch against data (Npol, Npol|1, Nbl,
        Ntimes, Nfreqs) -- what optim.LogProb.forward_chisq computes from forward()
        (optim.py:1012-1024, apply_icov :1836 with cov_axis None), fused into the unit reduction
        of the fringe-sum kernels so the visibilities never reach HBM and the backward pass starts
        from the cotangent the epilogue wrote.  Falls back to forward() + torch for the cases the
        epilogue does not cover (several sky components, several beam models, complex perceived
        sky, data_bls inflation).  Returns (chisq, residual-free handle): the float64 scalar and
        the cotangent tensor G = 2 icov (V - data) (None on the fallback route)."""
        self._set_group()
        comps = self.sky.forward(prior_cache=prior_cache)
        comps = comps if isinstance(comps, list) else [comps]
        plan = None
        if len(comps) == 1 and self._sim2data[self._bl_key] is None:
            if hasattr(self.beam.R, 'clear_beam_cache'):
                self.beam.R.clear_beam_cache()
            self.beam.skycut_device = getattr(self.sky, 'device', None)
            plan = self._single_plane_set(comps[0])
        if plan is None:
            vd = self.forward(prior_cache=prior_cache)
            res = vd.data - data.to(vd.data.device)
            w = 1.0 if icov is None else icov.to(res.device)
            return (res.real ** 2 + res.imag ** 2).mul(w).sum().double(), None
        A, (P, Q), rec, blvecs, f64, nfreq, uniform, dev = plan
        self.beam.eval_prior(prior_cache)
        nbl, nt = len(self.sim_bls), self.Ntimes
        D = data.to(dev).reshape(P * Q, nbl, nt, nfreq)
        W = icov.to(dev).reshape(P * Q, nbl, nt, nfreq) if icov is not None else None
        tc = til = None
        if A.dtype == torch.float32:
            tc, til = self._tc_tiling(dev), self._ant_tiling(dev)
        if tc is not None or til is not None:
            chi, G = ops.fringe_chisq(A, self.array.antvecs.to(dev), rec.geom, f64, nfreq, D, W,
                                      tiling=til, tc=tc)
        else:
            chi, G = ops.fringe_chisq(A, blvecs, rec.geom, f64, nfreq, D, W, uniform=uniform)
        return chi, G.reshape(P, Q, nbl, nt, nfreq)

    def _history(self):
        return "bayeslim_b200.RIME | sky={} beam={}({}) array={} Nbls={} Ntimes={} Nfreqs={}".format(
            self.sky.__class__.__name__, self.beam.__class__.__name__,
            self.beam.R.__class__.__name__, self.array.__class__.__name__, len(self.sim_bls),
            self.Ntimes, self.Nfreqs)

    def run_batches(self, concat=True):
        """forward() for every minibatch, concatenated over baselines then times
        (rime_model.py:442-482)."""
        vis_times, vis_bls = [], []
        for i in range(self.Nbatch):
            self.batch_idx = i
            vis = self.forward()
            vis_bls.append(vis)
            if self.Nbatch == 1:
                vis_times.append(vis)
            elif self.bl_group_id == self.Nbl_groups - 1:
                if concat:
                    vis_times.append(dataset.concat_VisData(vis_bls, 'bl'))
                else:
                    vis_times.extend(vis_bls)
                vis_bls = []
        out = dataset.concat_VisData(vis_times, 'time') if concat else vis_times
        self.batch_idx = 0
        return out


def log(message, verbose=False, style=1):
    if verbose:
        if style == 1:
            print("{}".format(message))
        elif style == 2:
            print("{}\n{}".format(message, '-' * 30))
        else:
            print("\n{}\n{}\n{}".format('-' * 30, message, '-' * 30))


def elapsed_time(start):
    t = datetime.now().timestamp() - start
    unit = 'sec'
    if t > 60000:
        t, unit = t / 3600, 'hrs'
    elif t > 1000:
        t, unit = t / 60, 'min'
    return "{:.3f} {}".format(t, unit)


class GraphedStep:
    """One optimisation step -- loss_fn(rime()) and its backward -- captured ONCE into a CUDA
    graph and replayed: for small problems (HERA-37 x 1k sources: 0.3 ms of kernels behind
    1.8 ms of Python, autograd bookkeeping and ~80 launches) the step becomes one graph launch.

    Usage (whole-step capture, static shapes):

        step = GraphedStep(rime, lambda vd: ((vd.data.real ** 2 + vd.data.imag ** 2).sum()), params)
        loss = step()            # replays; params[i].grad hold the new gradients
        opt.step()               # in-place parameter updates are seen by the next replay

    Requirements: parameters are updated in place (their storage is baked into the graph), the
    geometry / interpolation caches are warm (the warm-up steps run here do that), one minibatch
    (rime.Nbatch == 1 or a fixed batch_idx), no host synchronisation inside the step (the
    library's launches and its index tables are capture-safe; a user loss_fn must be too).
    The reference has no counterpart (its forward is eager torch)."""

    def __init__(self, rime, loss_fn, params, warmup=3):
        self.params = [p for p in params if p is not None]
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError("GraphedStep needs CUDA tensors (the hot path has no CPU implementation)")
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                for p in self.params:
                    p.grad = None
                loss_fn(rime()).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = rime()
            self.loss = loss_fn(self.out)
            self.loss.backward()

    def __call__(self):
        self.graph.replay()
        return self.loss
