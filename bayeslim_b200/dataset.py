"""
Output / input containers of the RIME path, mirroring the reference's dataset module
(bayeslim/dataset.py): ``VisData`` (:289, visibilities of shape
(Npol, Npol, Nbls, Ntimes, Nfreqs)) and ``MapData`` (:1867, sky maps of shape
(Npol, 1, Nfreqs, Npix)).  Only what ``RIME.forward`` / ``run_batches`` need is here
(metadata setup, push, concatenation); selection, averaging and HDF5 IO are out of scope.
"""
import numpy as np
import torch

from . import utils


class TensorData:
    """Bare tensor holder with optional flags / (inverse) covariance (dataset.py:15-288)."""

    def __init__(self):
        self.setup_data()

    def setup_data(self, data=None, flags=None, cov=None, cov_axis=None, icov=None, history=''):
        self.data = data
        self.flags = flags
        self.set_cov(cov, cov_axis, icov=icov)
        self.history = history

    def set_cov(self, cov, cov_axis, icov=None):
        logdet = None
        if isinstance(cov, torch.Tensor):
            if cov_axis is None:
                logdet = torch.sum(torch.log(cov))
            elif cov_axis == 'full':
                logdet = torch.slogdet(cov).logabsdet
            else:
                logdet = torch.slogdet(cov).logabsdet.sum()
            if torch.is_complex(logdet):
                logdet = logdet.real
            cov = cov.clone()
        elif isinstance(icov, torch.Tensor) and cov_axis is None:
            logdet = torch.sum(-torch.log(icov))
        if isinstance(icov, torch.Tensor):
            icov = icov.clone()
        self.cov, self.icov, self.cov_axis = cov, icov, cov_axis
        self.cov_ndim = sum(self.data.shape) if self.data is not None else None
        self.cov_logdet = logdet if logdet is not None else torch.tensor(0.0)

    @property
    def device(self):
        return self.data.device if self.data is not None else None

    @property
    def dtype(self):
        return self.data.dtype if self.data is not None else None

    def push(self, device, return_obj=False):
        dtype = isinstance(device, torch.dtype)
        self.data = utils.push(self.data, device)
        if not dtype:
            self.flags = utils.push(self.flags, device)
        self.cov = utils.push(self.cov, device)
        self.icov = utils.push(self.icov, device)
        if return_obj:
            return self


class VisData(TensorData):
    """Visibility data of shape (Npol, Npol, Nbl, Ntimes, Nfreqs)."""

    def __init__(self):
        self.data = None
        self.atol = 1e-10
        self._file = None
        self.setup_meta()

    def setup_meta(self, telescope=None, antpos=None):
        self.telescope = telescope
        if antpos is not None and not (hasattr(antpos, 'ants') and hasattr(antpos, 'antvecs')):
            antpos = utils.AntposDict(list(antpos.keys()), list(antpos.values()))
        self.antpos = antpos
        self.ants = antpos.ants if antpos is not None else None

    def setup_data(self, bls, times, freqs, pol=None, data=None, flags=None, cov=None,
                   cov_axis=None, icov=None, history='', file=None, _blnums=None):
        """_blnums (extension): (device tensor, numpy array) of the baseline numbers prepared by
        the caller (RIME caches them per baseline group, so that a forward neither copies them
        to the device nor synchronises to read them back)."""
        self.data = data
        if _blnums is not None:
            self._blnums, self.blnums = _blnums
            self.Nbls = len(self.blnums)
        else:
            self._set_bls(bls)
        self.times = torch.as_tensor(times)
        self.Ntimes = len(times)
        self.freqs = torch.as_tensor(freqs)
        self.Nfreqs = len(freqs)
        self.pol = pol
        if isinstance(pol, str):
            assert pol.lower() in ['ee', 'nn'], "pol must be 'ee' or 'nn' for 1pol mode"
        self.Npol = 2 if self.pol is None else 1
        self.flags = flags
        self.set_cov(cov, cov_axis, icov=icov)
        self.history = history
        self._file = file

    def _set_bls(self, bls):
        if isinstance(bls, torch.Tensor):
            self._blnums = bls.clone()
        elif isinstance(bls, np.ndarray):
            self._blnums = torch.as_tensor(bls.copy())
        else:
            self._blnums = torch.as_tensor(np.asarray(utils.ants2blnum(list(bls))))
        if self.data is not None and not utils.check_devices(self._blnums.device, self.data.device):
            self._blnums = self._blnums.to(self.data.device)
        self.blnums = self._blnums.cpu().numpy()
        self.Nbls = len(self.blnums)

    @property
    def bls(self):
        return utils.blnum2ants(self.blnums)

    def push(self, device, return_obj=False):
        dtype = isinstance(device, torch.dtype)
        super().push(device)
        if self.antpos:
            self.antpos.push(device)
        if self.telescope:
            self.telescope.push(device)
        self.freqs = utils.push(self.freqs, device)
        if not dtype:
            self._blnums = utils.push(self._blnums, device)
        if return_obj:
            return self

    def copy(self, detach=True):
        vd = VisData()
        vd.setup_meta(self.telescope, self.antpos)
        data = self.data.detach().clone() if detach else self.data.clone()
        vd.setup_data(self.bls, self.times, self.freqs, pol=self.pol, data=data, flags=self.flags,
                      cov=self.cov, cov_axis=self.cov_axis, icov=self.icov, history=self.history)
        return vd


def concat_VisData(vds, axis, run_check=True, interleave=False, lazy=False, device=None,
                   non_blocking=True):
    """Concatenate VisData along 'bl', 'time' or 'freq' (dataset.py:3739-3865)."""
    if isinstance(vds, VisData):
        return vds
    assert len(vds) > 0
    if len(vds) == 1:
        return vds[0]
    if interleave or lazy:
        raise NotImplementedError("interleave / lazy concatenation is outside the RIME path")
    vd = vds[0]
    bls, times, freqs = vd.bls, vd.times, vd.freqs
    if axis == 'bl':
        dim, bls = 2, [bl for o in vds for bl in o.bls]
    elif axis == 'time':
        dim, times = 3, torch.cat([torch.as_tensor(o.times) for o in vds])
    elif axis == 'freq':
        dim, freqs = 4, torch.cat([torch.as_tensor(o.freqs) for o in vds])
    else:
        raise ValueError(axis)
    data = torch.cat([o.data for o in vds], dim=dim)
    flags = None
    if vd.flags is not None:
        flags = torch.cat([o.flags for o in vds], dim=dim)
    out = VisData()
    out.setup_meta(vd.telescope, vd.antpos)
    out.setup_data(bls, times, freqs, pol=vd.pol, data=data, flags=flags, history=vd.history)
    if device is not None:
        out.push(device)
    return out


class MapData(TensorData):
    """Sky map of shape (Npol, 1, Nfreqs, Npix) with pixel centres `angs` = (RA, Dec) [deg]."""

    def __init__(self):
        self.data = None
        self.atol = 1e-10

    def setup_meta(self, name=None):
        self.name = name

    def setup_data(self, freqs, df=None, pols=None, data=None, angs=None, flags=None, cov=None,
                   cov_axis=None, icov=None, norm=None, history=''):
        self.freqs = freqs
        self.df = df
        self.angs = angs
        self.Nfreqs = len(freqs)
        self.pols = pols
        self.data = data
        self.flags = flags
        self.norm = norm
        self.set_cov(cov, cov_axis, icov=icov)
        self.history = history

    def push(self, device, return_obj=False):
        super().push(device)
        if not isinstance(device, torch.dtype):
            self.angs = utils.push(self.angs, device)
        self.freqs = utils.push(self.freqs, device)
        if return_obj:
            return self
