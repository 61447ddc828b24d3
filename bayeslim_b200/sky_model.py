"""
Sky components feeding the RIME, mirroring the reference's sky_model
(bayeslim/sky_model.py): they are *producers* of the (Nstokes|2,2 ; Nfreqs ; Nsources) sky
tensor and its (RA, Dec); their arithmetic is O(Nf Nsrc) -- not multiplied by the number of
baselines -- and stays in torch.  ``RIME`` consumes ``forward()``'s
MapData exactly as the reference does (rime_model.py:306).
"""
import torch

from . import utils, dataset


class SkyBase(utils.Module):
    """params -> R(params + p0) -> MapData (sky_model.py:13-140)."""

    def __init__(self, params, R=None, name=None, parameter=True, p0=None):
        super().__init__(name=name)
        self.params = params
        self.device = self.params.device
        self.p0 = p0
        if parameter:
            self.params = torch.nn.Parameter(self.params)
        self.R = R if R is not None else DefaultResponse()
        self._args = dict(name=name)
        self._args[self.R.__class__.__name__] = getattr(self.R, '_args', None)

    def _push(self, device, attrs=[]):
        dtype = isinstance(device, torch.dtype)
        if not dtype:
            self.device = device
        self.params = utils.push(self.params, device)
        for attr in attrs:
            if hasattr(self, attr):
                setattr(self, attr, getattr(self, attr).to(device))
        self.R.push(device)
        if self.p0 is not None:
            self.p0 = utils.push(self.p0, device)
        if isinstance(self.angs, torch.Tensor):
            if not dtype or self.angs.is_floating_point():
                self.angs = utils.push(self.angs, device)
        else:
            self.angs = tuple(utils.push(a, device) for a in self.angs)
        for priors in (self.priors_inp_params, self.priors_out_params):
            if priors is not None:
                for pr in priors:
                    if pr is not None:
                        pr.push(device)

    def push(self, device, **kwargs):
        self._push(device, **kwargs)

    def _response(self, params=None):
        p = self.params if params is None else params
        if self.p0 is not None:
            p = p + self.p0
        sky = self.R(p)
        if getattr(self, '_hook_registry', None) is not None and sky.requires_grad:
            for hook in self._hook_registry:
                sky.register_hook(hook)
        return sky

    def _mapdata(self, sky):
        comp = dataset.MapData()
        comp.setup_meta(name=getattr(self, 'name', None))
        angs = torch.vstack(list(self.angs)) if isinstance(self.angs, (tuple, list)) \
            else torch.as_tensor(self.angs)
        freqs = self.R.freqs
        idx = getattr(self.R, '_freq_idx', None)
        if idx is not None and freqs is not None:
            freqs = freqs[idx]
        comp.setup_data(freqs=freqs, data=sky, angs=angs)
        return comp


class DefaultResponse:
    def __init__(self, freqs=None):
        self.freqs = freqs
        self.freq_mode = 'channel'

    def set_freq_index(self, idx=None):
        pass

    def _setup(self):
        pass

    def __call__(self, params):
        return params

    def push(self, device):
        pass


class PointSky(SkyBase):
    """Point sources at fixed (RA, Dec) with parameterised flux density
    (sky_model.py:154-286).  params: (Nstokes, 1, Ncoeff, Nsources); angs: (2, Nsources) deg."""

    def __init__(self, params, angs, R=None, name=None, parameter=True, p0=None):
        super().__init__(params, R=R, name=name, parameter=parameter, p0=p0)
        self.angs = angs

    def forward(self, params=None, prior_cache=None, **kwargs):
        sky = self._response(params)
        self.eval_prior(prior_cache, inp_params=self.params, out_params=sky)
        return self._mapdata(sky)


class PointSkyResponse:
    """Frequency parameterisation of point sources: 'channel', 'linear' (via freq_LM) or
    'powerlaw' (amp * (nu/f0)^alpha) (sky_model.py:289-386)."""

    def __init__(self, freqs, freq_mode='linear', log=False, device=None, LM=None, freq_LM=None,
                 f0=None):
        self.log = log
        self.freqs = freqs
        self.freq_mode = freq_mode
        self.device = device
        self.LM = LM
        self.freq_LM = freq_LM
        self.f0 = f0
        self._args = dict(freq_mode=self.freq_mode)

    def __call__(self, params):
        if not utils.check_devices(params.device, self.device):
            params = params.to(self.device)
        if self.LM is not None:
            params = self.LM(params)
        if self.freq_mode == 'linear':
            params = self.freq_LM(params)
        elif self.freq_mode == 'powerlaw':
            amp = params[..., 0:1, :]
            if self.log:
                amp = torch.exp(amp)
            params = amp * (self.freqs[:, None] / self.f0) ** params[..., 1:2, :]
        if self.log and self.freq_mode in ['channel', 'linear']:
            params = torch.exp(params)
        if getattr(self, '_freq_idx', None) is not None:
            params = params[..., self._freq_idx, :]
        return params

    def set_freq_index(self, idx=None):
        self._freq_idx = idx

    def _setup(self):
        pass

    def push(self, device):
        if not isinstance(device, torch.dtype):
            self.device = device
        self.freqs = self.freqs.to(device)
        if self.LM is not None:
            self.LM.push(device)
        if self.freq_LM is not None:
            self.freq_LM.push(device)


class PixelSky(SkyBase):
    """Pixelised specific-intensity sky; forward() multiplies by the pixel solid angle to give
    flux density (sky_model.py:389-507)."""

    def __init__(self, params, angs, px_area, R=None, name=None, parameter=True, p0=None):
        super().__init__(params, R=R, name=name, parameter=parameter, p0=p0)
        self.angs = angs
        # scalars stay Python floats (exact in any session dtype); arrays become tensors
        self.px_area = px_area if isinstance(px_area, (int, float)) else torch.as_tensor(px_area)

    def forward(self, params=None, prior_cache=None, **kwargs):
        sky = self._response(params)
        self.eval_prior(prior_cache, inp_params=self.params, out_params=sky)
        return self._mapdata(sky * self.px_area)

    def push(self, device, **kwargs):
        self._push(device, **kwargs)
        self.px_area = utils.push(self.px_area, device)


class PixelSkyResponse:
    """Spatial ('pixel' | 'linear' via spat_LM | 'alm': spat_LM is a sph_harm.AlmModel, the a_lm ->
    pixel product runs on the CUDA spherical-harmonic GEMM) and frequency ('channel' | 'linear' |
    'powerlaw') parameterisation of a PixelSky (sky_model.py:510-732).  The 'bessel' frequency
    mode needs the spherical Fourier-Bessel stack and the cosmology module, which are outside
    the RIME path."""

    def __init__(self, freqs, comp_params=False, spatial_mode='pixel', freq_mode='channel',
                 device=None, transform_order=0, cosmo=None, spat_LM=None, freq_LM=None, f0=None,
                 gln=None, kbins=None, log=False, real_output=True, abs_output=False, LM=None,
                 sky0=None):
        if freq_mode == 'bessel':
            raise NotImplementedError("the spherical Fourier-Bessel sky parameterisation is out of scope")
        if spatial_mode in ('alm', 'linear'):
            assert spat_LM is not None, "spatial_mode '%s' needs spat_LM" % spatial_mode
        self.freqs = freqs
        self.comp_params = comp_params
        self.Nfreqs = len(freqs)
        self.spatial_mode = spatial_mode
        self.freq_mode = freq_mode
        self.device = device
        self.transform_order = transform_order
        self.log = log
        self.LM = LM
        self.real_output = real_output
        self.abs_output = abs_output
        self.sky0 = sky0
        self.freq_LM = freq_LM
        self.spat_LM = spat_LM
        self.f0 = f0
        self._args = dict(freq_mode=self.freq_mode, spatial_mode=self.spatial_mode)

    def spatial_transform(self, params):
        if self.comp_params and not torch.is_complex(params):
            params = utils.viewcomp(params)
        if self.spatial_mode == 'pixel':
            return params
        return self.spat_LM(params)

    def freq_transform(self, params):
        if self.comp_params and not torch.is_complex(params):
            params = utils.viewcomp(params)
        if self.freq_mode == 'channel':
            return params
        if self.freq_mode == 'linear':
            return self.freq_LM(params)
        if self.freq_mode == 'powerlaw':
            return params[..., 0:1, :] * (self.freqs[:, None] / self.f0) ** params[..., 1:2, :]
        raise ValueError(self.freq_mode)

    def __call__(self, params):
        if not utils.check_devices(params.device, self.device):
            params = params.to(self.device)
        if self.LM is not None:
            params = self.LM(params)
        if self.transform_order == 0:
            params = self.freq_transform(self.spatial_transform(params))
        else:
            params = self.spatial_transform(self.freq_transform(params))
        if self.real_output:
            params = params.real
        if self.log:
            params = torch.exp(params)
        if getattr(self, '_freq_idx', None) is not None:
            params = params[..., self._freq_idx, :]
        if self.sky0 is not None:
            params = params + self.sky0
        if self.abs_output:
            params = params.abs()
        return params

    def set_freq_index(self, idx=None):
        self._freq_idx = idx

    def _setup(self):
        pass

    def push(self, device):
        if self.spat_LM is not None:
            self.spat_LM.push(device)
        if self.freq_LM is not None:
            self.freq_LM.push(device)
        self.freqs = self.freqs.to(device)
        if self.LM is not None:
            self.LM.push(device)
        if not isinstance(device, torch.dtype):
            self.device = device
        if self.sky0 is not None:
            self.sky0 = utils.push(self.sky0, device)


class CompositeModel(utils.Module):
    """Several sky models evaluated together; returns the list of their MapData
    (sky_model.py:778-935, sum_output=False branch).  RIME sums the visibilities of the
    components (the reference's own multi-component sum raises a TypeError,
    rime_model.py:377 -- SURVEY section 9 item 5 -- so the list form has no reference oracle)."""

    def __init__(self, models, name=None):
        super().__init__(name=name)
        self.models = list(models)
        for k, m in models.items():
            self.add_module(k, m)
        first = self.get_submodule(self.models[0])
        self.device = first.device

    def forward(self, *args, prior_cache=None, **kwargs):
        return [self.get_submodule(k).forward(prior_cache=prior_cache) for k in self.models]

    def push(self, device):
        for k in self.models:
            self.get_submodule(k).push(device)
        if not isinstance(device, torch.dtype):
            self.device = device


class Stokes2Coherency(utils.Module):
    """Stokes [I, fQ, fU, fV] (Q = I fQ, ...) -> coherency [[I+Q, U-iV], [U+iV, I-Q]]
    (sky_model.py:1160-1353).  Input (Nstokes, 1, ...) tensor or MapData."""

    def __init__(self, params=None, parameter=False):
        super().__init__()
        self.params = params
        if parameter and isinstance(self.params, torch.Tensor):
            self.params = torch.nn.Parameter(self.params)

    def forward(self, sky_comp, prior_cache=None):
        if isinstance(sky_comp, dataset.MapData):
            sky_comp.data = self.forward(sky_comp.data, prior_cache=prior_cache)
            return sky_comp
        S = sky_comp
        I = S[0, 0]
        zero = 0
        if len(S) == 1:
            if self.params is None:
                return S
            frac = self.params if isinstance(self.params, torch.Tensor) else self.params().data
            fQ = frac[0, 0]
            fU = frac[1, 0] if len(frac) > 1 else zero
            fV = frac[2, 0] if len(frac) > 2 else None
        elif S.shape[:2] == (2, 2):
            fQ, fU, fV = S[0, 1], S[1, 0], S[1, 1]
        else:
            fQ = S[1, 0]
            fU = S[2, 0] if len(S) > 2 else zero
            fV = S[3, 0] if len(S) > 3 else None
        Q, U = I * fQ, I * fU
        rdt = I.dtype
        cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
        B = torch.zeros(2, 2, *S.shape[2:], dtype=rdt if fV is None else cdt, device=S.device)
        B[0, 0] = I + Q
        B[1, 1] = I - Q
        if fV is None:
            B[0, 1] = U
            B[1, 0] = U
        else:
            V = I * fV
            B[0, 1] = torch.complex(U + 0 * I, -V)
            B[1, 0] = torch.complex(U + 0 * I, V)
        frac_pol = fQ ** 2 + fU ** 2 + (fV ** 2 if fV is not None else 0)
        self.eval_prior(prior_cache, inp_params=frac_pol)
        return B
