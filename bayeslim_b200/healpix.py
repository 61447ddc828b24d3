"""
Closed-form HEALPix (RING ordering) helpers standing in for the healpy calls on the RIME
path: ``healpy.get_interp_weights`` (reference utils.py:765-769) and the
``pix2ang / nside2npix / nside2pixarea`` calls used to lay out pixel skies
(reference tests/test_sky.py:15-19).  Vectorised numpy; healpy itself is not required.
"""
import math

import numpy as np
import torch


def nside2npix(nside):
    return 12 * nside * nside


def nside2pixarea(nside):
    return 4 * math.pi / nside2npix(nside)


def _rings(nside):
    """start pixel, pixel count, z = cos(theta) and half-pixel phase shift of every ring."""
    r = np.arange(1, 4 * nside)
    ncap = 2 * nside * (nside - 1)
    npix = nside2npix(nside)
    north, south = r < nside, r > 3 * nside
    belt = ~(north | south)
    rs = 4 * nside - r
    npr = np.where(north, 4 * r, np.where(south, 4 * rs, 4 * nside))
    start = np.where(north, 2 * r * (r - 1),
                     np.where(south, npix - 2 * rs * (rs + 1), ncap + (r - nside) * 4 * nside))
    z = np.where(north, 1.0 - r * r / (3.0 * nside * nside),
                 np.where(south, -1.0 + rs * rs / (3.0 * nside * nside),
                          (2 * nside - r) * 2.0 / (3.0 * nside)))
    shift = np.where(belt, ((r - nside) % 2 == 0) * 0.5, 0.5)
    return start.astype(np.int64), npr.astype(np.int64), z.astype(np.float64), shift


def pix2ang(nside, ipix=None):
    """(theta, phi) [rad] of RING pixel centres."""
    start, npr, z, shift = _rings(nside)
    ring = np.repeat(np.arange(len(start)), npr)
    j = np.arange(nside2npix(nside)) - start[ring]
    theta = np.arccos(z[ring])
    phi = (j + shift[ring]) * 2 * np.pi / npr[ring]
    if ipix is not None:
        return theta[ipix], phi[ipix]
    return theta, phi


def get_interp_weights(nside, theta, phi):
    """Bilinear interpolation neighbours: (inds (N,4) int64, wgts (N,4) float64) tensors.

    Linear in phi between the two bracketing pixels of the ring above and of the ring below,
    linear in theta between the rings; in the polar caps the missing ring is replaced by the
    mean of the four polar pixels.  theta, phi in radians (tensors or arrays)."""
    dev = theta.device if isinstance(theta, torch.Tensor) else None
    th = np.asarray(theta.detach().cpu() if isinstance(theta, torch.Tensor) else theta, dtype=np.float64)
    ph = np.mod(np.asarray(phi.detach().cpu() if isinstance(phi, torch.Tensor) else phi,
                           dtype=np.float64), 2 * np.pi)
    start, npr, z, shift = _rings(nside)
    rth = np.arccos(z)
    nring, npix, n = len(start), nside2npix(nside), len(th)
    ir = np.searchsorted(rth, th, side='right') - 1

    def pair(r):
        nr = npr[r]
        t = ph / (2 * np.pi / nr) - shift[r]
        i1 = np.floor(t).astype(np.int64)
        w = t - i1
        return start[r] + np.mod(i1, nr), start[r] + np.mod(i1 + 1, nr), 1.0 - w, w

    ra = np.clip(ir, 0, nring - 1)
    rb = np.clip(ir + 1, 0, nring - 1)
    a1, a2, wa1, wa2 = pair(ra)
    b1, b2, wb1, wb2 = pair(rb)
    denom = np.where(rb > ra, rth[rb] - rth[ra], 1.0)
    wt = np.clip((th - rth[ra]) / denom, 0.0, 1.0)
    inds = np.stack([a1, a2, b1, b2], axis=1)
    wgts = np.stack([(1 - wt) * wa1, (1 - wt) * wa2, wt * wb1, wt * wb2], axis=1)
    north = ir < 0
    if north.any():
        k = np.where(north)[0]
        w_t = th[k] / rth[0]
        inds[k] = np.arange(4)[None, :]
        wk = np.repeat(((1 - w_t) * 0.25)[:, None], 4, axis=1)
        np.add.at(wk, (np.arange(len(k)), a1[k]), w_t * wa1[k])
        np.add.at(wk, (np.arange(len(k)), a2[k]), w_t * wa2[k])
        wgts[k] = wk
    south = ir >= nring - 1
    if south.any():
        k = np.where(south)[0]
        w_t = (th[k] - rth[-1]) / (np.pi - rth[-1])
        base = npix - 4
        inds[k] = base + np.arange(4)[None, :]
        wk = np.repeat((w_t * 0.25)[:, None], 4, axis=1)
        p1, p2, w1, w2 = pair(np.full(n, nring - 1))
        np.add.at(wk, (np.arange(len(k)), p1[k] - base), (1 - w_t) * w1[k])
        np.add.at(wk, (np.arange(len(k)), p2[k] - base), (1 - w_t) * w2[k])
        wgts[k] = wk
    return torch.as_tensor(inds, device=dev), torch.as_tensor(wgts, device=dev)
