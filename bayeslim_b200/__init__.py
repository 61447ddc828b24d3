"""
bayeslim_b200 -- B200-native (sm_100a CUDA) implementation of BayesLIM's differentiable RIME
hot path, behind the reference's own Python API.

    import bayeslim_b200 as ba
    rime = ba.rime_model.RIME(sky, telescope, beam, array, sim_bls, times, freqs)
    vis = rime()           # VisData, data (Npol, Npol, Nbls, Ntimes, Nfreqs) on the GPU

Module names mirror the reference package (``bayeslim.rime_model`` etc.) for the subset that
lies on the path; ``ops`` and ``_lib`` are the binding of the C-ABI CUDA library
(include/b200rime.h), which must be built (``make -C bayeslim_b200/csrc``) -- importing this
package without it raises.
"""
from . import _lib                      # noqa: F401  (fails loudly if the library is missing)
from . import utils, paramdict, dataset, healpix, sph_harm, telescope_model, sky_model, beam_model
from . import ops, rime_model, imaging, calibration, parallel, optim
from .utils import D2R, _float, _cfloat
from .paramdict import ParamDict
from .dataset import VisData, MapData
from .rime_model import RIME

__version__ = utils.__version__
