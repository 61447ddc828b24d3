"""
Host-side mirror of the pieces of the reference's ``utils`` module that the RIME path
touches (reference: bayeslim/utils.py).  Same names, argument meaning and behaviour, so
code written against ``bayeslim.utils`` keeps working; implementations are new.

    Module / Sequential     utils.py:1123-1411   parameter access by dotted path, priors
    PixInterp               utils.py:684-878     neighbour indices + weights for beam maps
    AntposDict              utils.py:2280-2348   antenna number -> ENU vector
    arr_hash, push, ...     small helpers
"""
import math
from contextlib import nullcontext

import numpy as np
import torch

__version__ = "0.1.0"

D2R = math.pi / 180.0
viewreal = torch.view_as_real
viewcomp = torch.view_as_complex


def _float(numpy=False):
    """Real dtype of the session: follows torch.set_default_dtype (utils.py:52-63)."""
    ft = torch.get_default_dtype()
    if not numpy:
        return ft
    return {torch.float16: np.float16, torch.float32: np.float32, torch.float64: np.float64}[ft]


def _cfloat(float_type=None, numpy=False):
    """Complex dtype matching _float() (utils.py:66-82)."""
    ft = float_type if float_type is not None else torch.get_default_dtype()
    if not numpy:
        return {torch.float64: torch.complex128, torch.float32: torch.complex64,
                torch.float16: torch.complex32}[ft]
    return {torch.float64: np.complex128, torch.float32: np.complex64}[ft]


def colat2lat(theta, deg=True):
    """Co-latitude <-> latitude (its own inverse)."""
    return (90.0 if deg else math.pi / 2) - theta


# ----------------------------------------------------------------------------- devices
def parse_device(d):
    if d is None or isinstance(d, torch.device) and d.type == 'cpu' or d == 'cpu':
        return 'cpu'
    d = torch.device(d)
    return d if d.index is not None else torch.device(d.type, 0)


def check_devices(d1, d2):
    """True if two device specifiers name the same device (None == 'cpu')."""
    return parse_device(d1) == parse_device(d2)


def push(tensor, device, parameter=False):
    """Move a tensor (or an object with .push) to a device, or cast it to a dtype
    (complex tensors stay complex).  Parameters stay Parameters.  (utils.py:1683-1735)"""
    if tensor is None or device is None:
        return tensor
    if hasattr(tensor, 'push') and not isinstance(tensor, torch.Tensor):
        tensor.push(device)
        return tensor
    if not isinstance(tensor, torch.Tensor):
        return tensor
    if isinstance(device, torch.dtype):
        if not (tensor.is_floating_point() or tensor.is_complex()):
            return tensor
        if tensor.is_complex() and not device.is_complex:
            device = {torch.float16: torch.complex32, torch.float32: torch.complex64,
                      torch.float64: torch.complex128}[device]
    if parameter or isinstance(tensor, torch.nn.Parameter):
        return torch.nn.Parameter(tensor.to(device))
    return tensor.to(device)


def tensor2numpy(tensor, clone=True):
    if isinstance(tensor, torch.Tensor):
        tensor = tensor.detach().cpu()
        if clone:
            tensor = tensor.clone()
        return tensor.numpy()
    return tensor


def arr_hash(arr, pntr=False):
    """Cheap identity of an angle array: the injected ``_arr_hash`` attribute if present
    (RIME attaches (sky name, Nsources, time), rime_model.py:345-357), else a hash of
    (first, last, length).  (utils.py:1643-1680)"""
    if pntr:
        return id(arr)
    if hasattr(arr, '_arr_hash'):
        return arr._arr_hash
    if isinstance(arr, torch.Tensor):
        h = hash((arr[0].cpu().item(), arr[-1].cpu().item(), len(arr)))
        arr._arr_hash = h
        return h
    return hash((arr[0], arr[-1], len(arr)))


def clear_cache_depth(cache, depth):
    """FIFO-trim an insertion-ordered dict to `depth` entries (utils.py:881-899)."""
    if depth is None:
        return
    extra = len(cache) - depth
    if extra > 0:
        for k in list(cache.keys())[:extra]:
            del cache[k]


def flatten(arr, Nelem=None):
    """Flatten a list of lists (or an iterable of arrays) by one level."""
    out = []
    for sub in arr:
        if isinstance(sub, (list, tuple, np.ndarray, torch.Tensor)) and not (
                isinstance(sub, tuple) and len(sub) > 0 and isinstance(sub[0], (int, np.integer))):
            out.extend(list(sub))
        else:
            out.append(sub)
    return out


def split_into_groups(arr, Nelem=None, Ngroup=None, interleave=False):
    """Split a sequence into consecutive groups of Nelem elements (or into Ngroup groups)."""
    N = len(arr)
    if Nelem is None:
        Nelem = int(math.ceil(N / Ngroup))
    if interleave:
        Ng = int(math.ceil(N / Nelem))
        return [arr[i::Ng] for i in range(Ng)]
    return [arr[i:i + Nelem] for i in range(0, N, Nelem)]


def _make_hex(N, D=15):
    """Hexagonally packed array with N antennas per side (3N^2-3N+1 in total), spacing D [m].
    Returns (ants, antvecs[Nant, 3]) centred on the array mean, z = 0.  (utils.py:1943-1962)"""
    xs, ys = [], []
    for row in range(2 * N - 1):
        extra = min(row, 2 * N - 2 - row)
        for j in range(N + extra):
            xs.append(j - 0.5 * extra)
            ys.append(row * math.sin(math.pi / 3))
    xs = np.asarray(xs) - np.mean(xs)
    ys = np.asarray(ys) - np.mean(ys)
    return list(range(len(xs))), np.vstack([xs, ys, np.zeros_like(xs)]).T * D


class SimpleIndex:
    """Mapping that answers every key with the same value (default ant2beam)."""

    def __init__(self, value=0):
        self.value = value

    def __getitem__(self, k):
        return self.value


# ----------------------------------------------------------------------------- baselines
def ants2blnum(antnums, separate=False, tensor=False):
    """(ant1, ant2) -> baseline integer ant1*1000 + ant2 with a +100 offset on each antenna
    (utils.py:2416-2468)."""
    if isinstance(antnums, tuple) and isinstance(antnums[0], (int, np.integer)):
        return int((antnums[0] + 100) * 1000 + (antnums[1] + 100))
    a = np.asarray(antnums)
    nums = (a[:, 0] + 100) * 1000 + (a[:, 1] + 100)
    return torch.as_tensor(nums) if tensor else nums


def blnum2ants(blnum, separate=False):
    """Inverse of ants2blnum; tuples and lists of tuples pass through (utils.py:2352-2413)."""
    if isinstance(blnum, tuple):
        return blnum
    if isinstance(blnum, list):
        if len(blnum) == 0 or isinstance(blnum[0], tuple):
            return blnum
        if isinstance(blnum[0], list):
            return [tuple(b) for b in blnum]
    if isinstance(blnum, (int, np.integer)):
        return (int(blnum // 1000) - 100, int(blnum % 1000) - 100)
    b = np.asarray(tensor2numpy(blnum))
    a1, a2 = b // 1000 - 100, b % 1000 - 100
    if separate:
        return a1, a2
    return [(int(i), int(j)) for i, j in zip(a1, a2)]


def conjbl(bl):
    return (bl[1], bl[0])


class AntposDict:
    """Dictionary of antenna positions held as one contiguous (Nants, 3) tensor
    (utils.py:2280-2348)."""

    def __init__(self, ants, antvecs):
        self.ants = [int(a) for a in ants]
        self._ant_idx = {a: i for i, a in enumerate(self.ants)}
        if isinstance(antvecs, torch.Tensor):
            self.antvecs = antvecs
        else:
            antvecs = list(antvecs)
            if len(antvecs) and isinstance(antvecs[0], torch.Tensor):
                self.antvecs = torch.vstack([a.reshape(1, -1) for a in antvecs])
            else:
                self.antvecs = torch.as_tensor(np.asarray(antvecs))

    def keys(self):
        return iter(self.ants)

    def values(self):
        return iter(self.antvecs)

    def items(self):
        return zip(self.ants, self.antvecs)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self.antvecs[self._ant_idx[int(key)]]
        if isinstance(key, torch.Tensor):
            key = key.tolist()
        return self.antvecs[[self._ant_idx[int(k)] for k in key]]

    def __setitem__(self, key, value):
        self.antvecs[self._ant_idx[key]] = value

    def __len__(self):
        return len(self.ants)

    def __contains__(self, key):
        return key in self._ant_idx

    def __iter__(self):
        return self.keys()

    def __repr__(self):
        return "Antpos{{{}}}".format(self.ants)

    def push(self, device):
        self.antvecs = push(self.antvecs, device)

    def select(self, new_ants):
        return AntposDict(new_ants, self.antvecs[[self._ant_idx[a] for a in new_ants]])


# ----------------------------------------------------------------------------- attribute paths
def has_model_attr(model, name):
    parts = name.split('.') if isinstance(name, str) else list(name)
    obj = model
    for p in parts:
        if not hasattr(obj, p):
            return False
        obj = getattr(obj, p)
    return True


def get_model_attr(model, name, pop=0):
    """model.a.b.c for name 'a.b.c'; `pop` drops that many trailing components."""
    parts = name.split('.') if isinstance(name, str) else list(name)
    if pop > 0:
        parts = parts[:-pop]
    obj = model
    for p in parts:
        obj = getattr(obj, p)
    return obj


def set_model_attr(model, name, value, clobber_param=False, no_grad=True, idx=None, add=False,
                   fill=None):
    """Assign `value` at the dotted path `name` (utils.py:1453-1545).

    If the target is a Parameter it stays one (its data is replaced) unless clobber_param;
    idx/add/fill select in-place insertion, accumulation and pre-fill."""
    parts = name.split('.') if isinstance(name, str) else list(name)
    owner = get_model_attr(model, parts[:-1]) if len(parts) > 1 else model
    leaf = parts[-1]
    with (torch.no_grad() if no_grad else nullcontext()):
        cur = getattr(owner, leaf, None)
        if cur is None:
            setattr(owner, leaf, value)
            return
        was_param = isinstance(cur, torch.nn.Parameter)
        if clobber_param or was_param:
            data = cur.data
            delattr(owner, leaf)
            setattr(owner, leaf, data)
            cur = data
        if isinstance(value, torch.Tensor) and not check_devices(cur.device, value.device):
            value = value.to(cur.device)
        if fill is not None:
            cur.data[:] = fill.to(cur.dtype) if isinstance(fill, torch.Tensor) else fill
        if add:
            if idx is None:
                cur += value
            else:
                cur[idx] += value
        elif idx is None:
            setattr(owner, leaf, value)
        else:
            cur[idx] = value
        if was_param and not clobber_param:
            setattr(owner, leaf, torch.nn.Parameter(getattr(owner, leaf)))


def del_model_attr(model, name):
    parts = name.split('.') if isinstance(name, str) else list(name)
    owner = get_model_attr(model, parts[:-1]) if len(parts) > 1 else model
    delattr(owner, parts[-1])


class Module(torch.nn.Module):
    """torch.nn.Module plus dotted-path access, ParamDict updates and log-priors on the
    input / response-mapped parameters (utils.py:1123-1320)."""

    def __init__(self, name=None):
        super().__init__()
        self.__version__ = __version__
        self.set_priors()
        self._name = name

    @property
    def name(self):
        return self._name if self._name is not None else self.__class__.__name__

    @property
    def named_params(self):
        return [k for k, _ in self.named_parameters()]

    def forward(self, inp=None, prior_cache=None, **kwargs):
        raise NotImplementedError

    def __getitem__(self, name):
        return get_model_attr(self, name)

    def __setitem__(self, name, value):
        with torch.no_grad():
            set_model_attr(self, name, value)

    def __delitem__(self, name):
        del_model_attr(self, name)

    def update(self, pdict, clobber_param=False):
        for key, val in pdict.items():
            set_model_attr(self, key, val, clobber_param=clobber_param)

    def unset_param(self, name):
        if isinstance(name, list):
            for n in name:
                self.unset_param(n)
            return
        param = self[name].detach()
        del self[name]
        self[name] = param

    def set_param(self, name):
        if isinstance(name, list):
            for n in name:
                self.set_param(n)
            return
        param = self[name]
        if not isinstance(param, torch.nn.Parameter):
            self[name] = torch.nn.Parameter(param)

    def set_priors(self, priors_inp_params=None, priors_out_params=None):
        if priors_inp_params is not None and not isinstance(priors_inp_params, (list, tuple)):
            priors_inp_params = [priors_inp_params]
        if priors_out_params is not None and not isinstance(priors_out_params, (list, tuple)):
            priors_out_params = [priors_out_params]
        self.priors_inp_params = priors_inp_params
        self.priors_out_params = priors_out_params

    def eval_prior(self, prior_cache, inp_params=None, out_params=None):
        """Sum the log-priors of this module into prior_cache[self.name] (once per key)."""
        if prior_cache is None or self.name in prior_cache:
            return
        total = torch.as_tensor(0.0)
        if inp_params is None and hasattr(self, 'params'):
            inp_params = self.params
        if self.priors_inp_params is not None and inp_params is not None:
            for prior in self.priors_inp_params:
                if prior is not None:
                    total = total + prior(inp_params)
        if self.priors_out_params is not None:
            if out_params is None and hasattr(self, 'params') and hasattr(self, 'R'):
                p = self.params
                if getattr(self, 'p0', None) is not None:
                    p = p + self.p0
                out_params = self.R(p)
            if out_params is not None:
                for prior in self.priors_out_params:
                    if prior is not None:
                        total = total + prior(out_params)
        prior_cache[self.name] = total

    def register_response_hooks(self, registry=None):
        if registry is not None and not isinstance(registry, (list, tuple)):
            registry = [registry]
        self._hook_registry = registry

    def clear_graph_tensors(self):
        pass


class Sequential(Module):
    """Evaluate sub-modules in order, optionally updating parameters from a ParamDict first;
    forwards the minibatch API of its first block (utils.py:1323-1411)."""

    def __init__(self, models):
        super().__init__()
        self._models = list(models)
        for name, model in models.items():
            self.add_module(name, model)

    def forward(self, inp=None, pdict=None, prior_cache=None, **kwargs):
        if pdict is not None:
            self.update(pdict)
        for name in self._models:
            inp = self.get_submodule(name)(inp, prior_cache=prior_cache, **kwargs)
        return inp

    @property
    def Nbatch(self):
        first = self.get_submodule(self._models[0])
        return first.Nbatch if hasattr(first, 'Nbatch') else 1

    @property
    def batch_idx(self):
        first = self.get_submodule(self._models[0])
        return first.batch_idx if hasattr(first, 'batch_idx') else 0

    @batch_idx.setter
    def batch_idx(self, val):
        first = self.get_submodule(self._models[0])
        if hasattr(first, 'batch_idx'):
            first.batch_idx = val


# ----------------------------------------------------------------------------- interpolation
_DEGREE = {'nearest': 0, 'linear': 1, 'quadratic': 2, 'cubic': 3}


def _uniform_nodes(grid, x, n, wrap):
    """First index of the n grid nodes nearest to x on a uniform grid and the offset of x
    from that node in grid steps.  Closed form of the reference's argsort-based search
    (utils.py:1003-1010): n even -> floor, n odd -> round, clamped (or wrapped) at the ends."""
    g0 = float(grid[0])
    dx = float(grid[1] - grid[0])
    N = len(grid)
    t = (x - g0) / dx
    if n % 2 == 0:
        start = torch.floor(t) - (n // 2 - 1)
    else:
        start = torch.floor(t + 0.5) - (n - 1) // 2
    if not wrap:
        start = torch.clamp(start, 0, N - n)
    rel = t - start
    idx = start.long()[:, None] + torch.arange(n, device=x.device)[None, :]
    if wrap:
        idx = idx % N
    return idx, rel


def _lagrange_weights(rel, n):
    cols = []
    for i in range(n):
        w = torch.ones_like(rel)
        for j in range(n):
            if j != i:
                w = w * (rel - j) / (i - j)
        cols.append(w)
    return torch.stack(cols, dim=-1)


class PixInterp:
    """Neighbour indices and weights for interpolating a pixelised map at (zen, az).

    pixtype 'rect': bi-polynomial interpolation on a uniform (phi, theta) grid, az wrapping
    (utils.py:772-798); weights are the tensor-product Lagrange basis, which is what the
    reference's least-squares solve evaluates to.  pixtype 'healpix': RING bilinear
    interpolation (the reference calls healpy.get_interp_weights, utils.py:765-769).
    Results are cached by arr_hash(zen) with an optional FIFO depth.
    """

    def __init__(self, pixtype, nside=None, interp_mode='nearest', theta_grid=None, phi_grid=None,
                 device=None, interp_cache_depth=None):
        self.pixtype = pixtype
        self.nside = nside
        self.interp_cache = {}
        self.interp_mode = interp_mode
        self.theta_grid = theta_grid
        self.phi_grid = phi_grid
        self.device = device
        self.interp_cache_depth = interp_cache_depth

    def clear_cache(self, depth=None):
        if depth is None:
            self.interp_cache = {}
        else:
            clear_cache_depth(self.interp_cache, depth)

    def _rect_weights(self, zen, az):
        mode = self.interp_mode
        deg = [_DEGREE[s.strip()] for s in mode.split(',')] if ',' in mode else [_DEGREE[mode]] * 2
        nx, ny = deg[0] + 1, deg[1] + 1
        zen = torch.as_tensor(zen, dtype=torch.float64)
        az = torch.as_tensor(az, dtype=torch.float64).to(zen.device)
        xi, xrel = _uniform_nodes(self.phi_grid, az, nx, wrap=True)
        yi, yrel = _uniform_nodes(self.theta_grid, zen, ny, wrap=False)
        nphi = len(self.phi_grid)
        inds = (xi[:, None, :] + nphi * yi[:, :, None]).reshape(len(zen), nx * ny)
        wx, wy = _lagrange_weights(xrel, nx), _lagrange_weights(yrel, ny)
        wgts = (wy[:, :, None] * wx[:, None, :]).reshape(len(zen), nx * ny)
        return inds, wgts            # float64; cast to the map dtype at the point of use

    def _healpix_weights(self, zen, az):
        from .healpix import get_interp_weights
        theta = torch.as_tensor(zen, dtype=torch.float64) * D2R
        phi = torch.as_tensor(az, dtype=torch.float64) * D2R
        return get_interp_weights(self.nside, theta, phi)

    def get_interp(self, zen, az):
        h = arr_hash(zen)
        if h in self.interp_cache:
            return self.interp_cache[h]
        if self.pixtype == 'healpix':
            inds, wgts = self._healpix_weights(zen, az)
        elif self.pixtype == 'rect':
            inds, wgts = self._rect_weights(zen, az)
        else:
            raise ValueError("pixtype must be 'healpix' or 'rect'")
        if self.interp_cache_depth is None or self.interp_cache_depth > 0:
            if not check_devices(inds.device, self.device):
                inds, wgts = inds.to(self.device), wgts.to(self.device)
            self.interp_cache[h] = (inds, wgts)
            if self.interp_cache_depth is not None:
                self.clear_cache(depth=self.interp_cache_depth)
        return inds, wgts

    def interp(self, m, zen, az):
        """out[..., s] = sum_i m[..., inds[s, i]] * wgts[s, i]   (torch; the RIME hot path uses
        the fused CUDA builder instead, see ops.build_interp)."""
        inds, wgts = self.get_interp(zen, az)
        inds = inds.to(m.device)
        nearest = m.index_select(-1, inds.reshape(-1)).view(m.shape[:-1] + inds.shape)
        return torch.einsum('...i,...i->...', nearest, wgts.to(device=m.device, dtype=nearest.dtype))

    def push(self, device):
        dtype = isinstance(device, torch.dtype)
        if not dtype:
            self.device = device
        self.theta_grid = push(self.theta_grid, device)
        self.phi_grid = push(self.phi_grid, device)
        for k, (inds, wgts) in list(self.interp_cache.items()):
            self.interp_cache[k] = (inds if dtype else push(inds, device), push(wgts, device))
