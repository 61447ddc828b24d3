"""
The likelihood step that consumes the RIME's visibilities, mirroring the part of the reference's
``optim.LogProb`` that sits on the hot path (bayeslim/optim.py:385-1190): minibatch selection,
``forward_chisq`` (residual against the target, ``apply_icov`` with a diagonal inverse covariance,
optim.py:959-1030 and :1836), the Gaussian log-likelihood and the log-prior from ``prior_cache``.

When the model is a bare ``bayeslim_b200.RIME`` the chi-square is fused into the visibility
kernels' unit reduction (``RIME.forward_chisq``): the visibilities are
never written to HBM and the backward pass starts from the cotangent the epilogue left.  Samplers,
optimisers, main-parameter packing and the dense covariance axes of ``apply_icov`` are out of scope
(SURVEY section 2).
"""
import numpy as np
import torch

from . import utils
from .dataset import TensorData


def apply_icov(data, icov, cov_axis, mode='vis'):
    """data^H Sigma^-1 data for a diagonal (cov_axis None) or full inverse covariance
    (optim.py:1836-1920); the per-axis dense forms are not mirrored."""
    if cov_axis is None:
        out = data.conj() * data
        return out if icov is None else out * icov
    if cov_axis == 'full':
        return data.ravel().conj() @ icov @ data.ravel()
    raise NotImplementedError("cov_axis %r: only None and 'full' exist here" % (cov_axis,))


class LogProb(utils.Module):
    """Negative log-posterior of a forward model against target data.

    model   : utils.Module returning a VisData / tensor (a RIME, or a Sequential starting with one)
    target  : VisData (or list of VisData, one per minibatch of the model) with .data and
              optionally .icov / .cov_axis
    """

    def __init__(self, model, target, start_inp=None, complex_circular=False, negate=True,
                 compute='post', fuse=True, name=None):
        super().__init__(name=name)
        self.model = model
        self.target = target if isinstance(target, (list, tuple)) else [target]
        self.start_inp = start_inp
        self.complex_circular, self.negate, self.compute, self.fuse = complex_circular, negate, compute, fuse
        self.prior_cache = {}
        self.batch_idx = 0

    @property
    def Nbatch(self):
        return getattr(self.model, 'Nbatch', 1) or 1

    @property
    def batch_idx(self):
        return self._batch_idx

    @batch_idx.setter
    def batch_idx(self, val):
        self._batch_idx = int(val)
        if hasattr(self.model, 'batch_idx') and (self.Nbatch > 1 or val == 0):
            try:
                self.model.batch_idx = int(val)
            except AttributeError:
                pass

    def clear_prior_cache(self):
        self.prior_cache = {}

    def get_batch_data(self, idx=None):
        if idx is not None:
            self.batch_idx = idx
        target = self.target[self.batch_idx if len(self.target) > 1 else 0]
        inp = None if self.start_inp is None else self.start_inp[self.batch_idx]
        return target, inp

    def _fusable(self, cov_axis, sum_chisq):
        return (self.fuse and sum_chisq and cov_axis is None
                and hasattr(self.model, 'forward_chisq'))

    def forward_chisq(self, idx=None, sum_chisq=True, **kwargs):
        """(chisq, res) of minibatch idx (optim.py:959-1030).  On the fused route (bare RIME,
        diagonal weights, summed chisq) `res` is None: the residual is not materialised; the
        cotangent 2 icov (V - data) is kept on self.cotangent instead."""
        target, inp = self.get_batch_data(idx)
        data = target.data
        icov = getattr(target, 'icov', None)
        cov_axis = getattr(target, 'cov_axis', None)
        if self.batch_idx == 0:
            self.clear_prior_cache()
        if self._fusable(cov_axis, sum_chisq):
            chisq, self.cotangent = self.model.forward_chisq(data, icov, prior_cache=self.prior_cache)
            if self.cotangent is not None:
                return chisq, None
            return chisq, None
        prediction = self.model(inp, prior_cache=self.prior_cache)
        if isinstance(prediction, TensorData):
            prediction = prediction.data
        res = prediction - data.to(prediction.device)
        chisq = apply_icov(res, None if icov is None else icov.to(res.device), cov_axis)
        if sum_chisq:
            chisq = torch.sum(chisq)
        if torch.is_complex(chisq):
            chisq = chisq.real
        return chisq, res

    def forward_like(self, idx=None, **kwargs):
        """Gaussian log-likelihood (negated when self.negate) of minibatch idx
        (optim.py:1032-1074)."""
        chisq, _ = self.forward_chisq(idx)
        target, _ = self.get_batch_data()
        norm = 0.0
        if getattr(target, 'icov', None) is not None and hasattr(target, 'cov_logdet'):
            if self.complex_circular:
                norm = target.cov_ndim * np.log(np.pi) + target.cov_logdet
            else:
                norm = 0.5 * (target.cov_ndim * np.log(2 * np.pi) + target.cov_logdet)
        loglike = (-chisq if self.complex_circular else -0.5 * chisq) - norm
        return -loglike if self.negate else loglike

    def forward_prior(self, idx=None, **kwargs):
        """Sum of the log-priors the modules left in prior_cache; counted on minibatch 0 only
        (optim.py:1076-1131)."""
        if idx is not None:
            self.batch_idx = idx
        total = torch.zeros(1)
        if self.batch_idx == 0:
            for v in self.prior_cache.values():
                total = total.to(v.device) + v
        return -total if self.negate else total

    def forward(self, idx=None, **kwargs):
        prob = self.forward_like(idx)
        if self.compute == 'post':
            prob = prob + self.forward_prior().to(prob.device).sum()
        return prob

    def closure(self, **kwargs):
        """Zero the gradients, evaluate every minibatch and backpropagate (optim.py:1191-1226)."""
        for p in self.parameters():
            p.grad = None
        loss = 0.0
        for i in range(self.Nbatch):
            out = self.forward(i)
            if out.requires_grad:
                out.backward()
            loss = loss + out.detach()
        return loss
