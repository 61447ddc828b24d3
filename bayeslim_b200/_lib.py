"""
ctypes binding of libb200rime.so (the C-ABI CUDA library declared in include/b200rime.h).

There is no fallback: if the library has not been built, importing this module raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C bayeslim_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RIME_LIB", os.path.join(_HERE, "csrc", "libb200rime.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "bayeslim_b200: %s not found. The CUDA library is the product; there is no CPU or "
        "PyTorch fallback. Build it with `make -C %s`." % (LIB_PATH, os.path.dirname(LIB_PATH)))

lib = ctypes.CDLL(LIB_PATH)

_P, _I, _L, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double

lib.b200rime_version.restype = ctypes.c_char_p
lib.b200rime_last_error.restype = ctypes.c_char_p
lib.b200rime_device_info.argtypes = [_I] + [ctypes.POINTER(_I)] * 4
lib.b200rime_kc.argtypes = [_I]
lib.b200rime_microbench.argtypes = [_I, _I, ctypes.POINTER(_D), ctypes.POINTER(_D)]
lib.b200rime_airy_bwd_blocks.argtypes = [_I, _I]
lib.b200rime_chisq_blocks.argtypes = [_I, _I]
lib.b200rime_eq2top_f64.argtypes = [_P, _P, _L, ctypes.POINTER(_D), ctypes.POINTER(_D), _P, _P, _P]
lib.b200rime_eq2top_f64.restype = _I

_SIGS = {
    "fringe_sum_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _I, _P, _P],
    "reduce_units": [_P, _P, _I, _I, _I, _P, _L, _L, _L, _D, _D, _I, _P],
    "reduce_units_chisq": [_P, _P, _I, _I, _I, _P, _P, _P, _L, _L, _L, _I, _P, _P],
    "fringe_sum_bwd_sky": [_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _I, _P, _P],
    "fringe_sum_bwd_bl": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _I, _I, _P, _P],
    "pack": [_P, _L, _I, _I, _I, _L, _L, _P, _P],
    "unpack": [_P, _L, _I, _I, _L, _L, _P, _P],
    "build_interp": [_P, _L, _P, _P, _I, _P, _L, _P, _I, _I, _I, _L, _L, _P, _P],
    "build_interp_t": [_P, _L, _P, _P, _I, _P, _L, _P, _I, _I, _I, _L, _L, _P, _P],
    "build_interp_bwd": [_P, _P, _L, _P, _P, _I, _P, _L, _P, _I, _I, _L, _L, _P, _P, _L, _P, _P],
    "gather_times": [_P, _L, _P, _I, _I, _I, _P, _L, _P],
    "interp_transpose": [_P, _L, _P, _P, _P, _I, _I, _P, _L, _P],
    "apply_cal": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "apply_cal_bwd_gains": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P,
                            _P],
    "jones_sandwich": [_P, _P, _P, _L, _P, _P],
    "jones_sandwich_bwd": [_P, _P, _P, _P, _L, _I, _P, _P, _P, _P],
    "build_airy": [_D, _D, _P, _D, _I, _P, _P, _P, _P, _L, _P, _I, _I, _I, _L, _L, _P, _P, _L, _P],
    "build_airy_bwd": [_P, _D, _D, _P, _D, _I, _I, _P, _P, _P, _P, _L, _P, _I, _I, _L, _L, _P, _P, _P,
                       _L, _P],
}
for _name, _sig in _SIGS.items():
    for _suffix in ("f32", "f64"):
        _fn = getattr(lib, "b200rime_%s_%s" % (_name, _suffix))
        _fn.argtypes = _sig
        _fn.restype = _I


_ANT_SIGS = {
    "antfringe_fwd": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _L, _I, _P, _P],
    "antfringe_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _I, _P, _P, _P],
    "tcfringe_fwd": [_P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _L, _I, _P, _P],
    "tcfringe_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _L, _I, _P, _P, _P],
}
_ANT_SIGS.update({
    "build_interp_bwd_t": [_P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _L, _L, _P, _L, _P, _P],
    "tc_pack_cotangent": [_P, _L, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "cgemm_pack_a": [_P, _P, _L, _L, _I, _I, _P, _I, _P, _P],
    "cgemm_pack_b": [_P, _P, _L, _L, _I, _I, _P, _I, _P, _P],
    "cgemm": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _L, _P, _P],
})
lib.b200rime_cgemm_f64.argtypes = [_P, _P, _L, _L, _P, _P, _L, _L, _I, _I, _I, _I, _I, _I, _P, _L, _P]
lib.b200rime_cgemm_f64.restype = _I
lib.b200rime_cgemm_a_bytes.argtypes = [_I, _I]
lib.b200rime_cgemm_a_bytes.restype = _L
lib.b200rime_cgemm_b_bytes.argtypes = [_I, _I]
lib.b200rime_cgemm_b_bytes.restype = _L
for _name, _sig in _ANT_SIGS.items():        # float32 only
    _fn = getattr(lib, "b200rime_%s_f32" % _name)
    _fn.argtypes = _sig
    _fn.restype = _I


class B200RimeError(RuntimeError):
    pass


def call(name, suffix, *args):
    """Invoke b200rime_<name>_<suffix>; raise B200RimeError with the library's message on failure."""
    rc = getattr(lib, "b200rime_%s_%s" % (name, suffix))(*args)
    if rc != 0:
        raise B200RimeError("b200rime_%s_%s failed (%d): %s" % (
            name, suffix, rc, lib.b200rime_last_error().decode()))


def version():
    return lib.b200rime_version().decode()


SRC_PAD = lib.b200rime_src_pad()
SRC_TILE = lib.b200rime_src_tile()
KC = {"f32": lib.b200rime_kc(0), "f64": lib.b200rime_kc(1)}
ANT_TILE = lib.b200rime_ant_tile()      # antennas per tile side of the antenna-factorised kernels
ANT_STAGE = lib.b200rime_ant_stage()    # reduction indices per shared-memory stage
TC_ROWS = lib.b200rime_tc_rows()        # first antennas per item of the tensor-core kernels
TC_COLS_MAX = lib.b200rime_tc_cols_max()  # second antennas per item, at most


def device_info(device=0):
    vals = [_I() for _ in range(4)]
    rc = lib.b200rime_device_info(device, *[ctypes.byref(v) for v in vals])
    if rc != 0:
        raise B200RimeError(lib.b200rime_last_error().decode())
    return dict(sm_count=vals[0].value, clock_khz=vals[1].value,
                cc=(vals[2].value, vals[3].value))


def microbench(kind, iters=4096):
    """kind: 'fp32' | 'fp64' | 'mufu' | 'fp32x2' -> (Gop/s, ms). FMA counted as 2 flop
    (a packed FFMA2 as 4)."""
    k = {"fp32": 0, "fp64": 1, "mufu": 2, "fp32x2": 3, "rf3_fp32": 4, "rf3_fp32x2": 5,
         "mix_rot_mac": 6, "mix_mac": 7, "mix_rot": 8, "rot_4ch": 9, "rot_8ch": 10,
         "rotmac_4ch": 11, "rotmac_8ch": 12, "rot_1ch": 13}[kind]
    g, ms = _D(), _D()
    rc = lib.b200rime_microbench(k, iters, ctypes.byref(g), ctypes.byref(ms))
    if rc != 0:
        raise B200RimeError(lib.b200rime_last_error().decode())
    return g.value, ms.value
