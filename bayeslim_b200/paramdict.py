"""
ParamDict: a dictionary of named parameter tensors with element-wise arithmetic, as consumed
by ``utils.Module.update`` (reference bayeslim/paramdict.py:8-305).  Boundary type only -- the
RIME never looks inside it.
"""
import torch

from . import utils


class ParamDict:
    def __init__(self, params):
        self.params = params
        self._setup()

    def _setup(self):
        self.devices = {k: self.params[k].device for k in self.keys()}

    def keys(self):
        return list(self.params.keys())

    def values(self):
        return list(self.params.values())

    def items(self):
        return list(self.params.items())

    def __iter__(self):
        return iter(self.params)

    def __len__(self):
        return len(self.params)

    def __contains__(self, key):
        return key in self.params

    def __getitem__(self, key):
        return self.params[key]

    def __setitem__(self, key, val):
        self.params[key] = val

    def push(self, device, inplace=True, copy=True):
        obj = self if inplace else (self.copy() if copy else self.clone())
        for k in obj.params:
            d = device if not isinstance(device, dict) else device[k]
            obj.params[k] = utils.push(obj.params[k], d)
        obj._setup()
        if not inplace:
            return obj

    def clone(self, **kwargs):
        return ParamDict({k: v.clone(**kwargs) for k, v in self.params.items()})

    def copy(self):
        out = {}
        for k, p in self.params.items():
            q = p.detach().clone()
            out[k] = torch.nn.Parameter(q) if p.requires_grad else q
        return ParamDict(out)

    def detach(self):
        return ParamDict({k: v.detach() for k, v in self.params.items()})

    def update(self, other):
        for key in other:
            self[key] = other[key]
        self._setup()

    def _binary(self, other, op):
        if isinstance(other, ParamDict):
            return ParamDict({k: op(self.params[k], other.params[k]) for k in self.keys()})
        return ParamDict({k: op(self.params[k], other) for k in self.keys()})

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self._binary(other, lambda a, b: a / b)

    def __neg__(self):
        return ParamDict({k: -v for k, v in self.params.items()})


def model2pdict(model, parameters=True):
    """Collect a model's named parameters into a ParamDict (paramdict.py:308-350)."""
    return ParamDict({k: (v if parameters else v.detach()) for k, v in model.named_parameters()})
