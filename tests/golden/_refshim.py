"""
Import the UNMODIFIED reference (``/root/reference/bayeslim``) in the build container.

The reference hard-imports astropy / h5py (and soft-imports healpy), none of which are
installed and none of which the RIME hot path calls once (zen, az) are injected through
``telescope.conv_cache``.  This module registers inert stub modules under those names
so that ``import bayeslim`` succeeds; nothing in the reference is patched.

Only ``tests/golden/make_golden.py`` (fixture generation, run by hand in the build
container) and the optional cross-check tests use this.  ``/root/reference`` does not
exist on the GPU box, so nothing on the ``-m gpu`` path may import it.
"""
import importlib.machinery
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("BAYESLIM_REFERENCE", "/root/reference")


class _Inert(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        mod = _Inert(self.__name__ + '.' + key)
        setattr(self, key, mod)
        return mod

    def __call__(self, *args, **kwargs):
        return self

    def __mul__(self, other):
        return self
    __rmul__ = __truediv__ = __rtruediv__ = __mul__


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bayeslim"))


def load():
    """Return the reference package (``import bayeslim``), stubbing missing third-party deps."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ['astropy', 'astropy.units', 'astropy.constants', 'astropy.coordinates',
                 'astropy.time', 'astropy.cosmology', 'h5py', 'healpy']:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = _Inert(name)
        mod.__path__ = []
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = mod
    cosmo = sys.modules['astropy.cosmology']
    if isinstance(cosmo, _Inert):
        cosmo.FlatLambdaCDM = type('FlatLambdaCDM', (), {'__init__': lambda s, *a, **k: None})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import bayeslim
    return bayeslim
