"""
Spherical-harmonic forward model (reference sph_harm.py:1244-1745
``AlmModel``, beam_model.py:1019-1267 ``YlmResponse``): package against golden vectors from the
unmodified reference (tests/golden/alm_forward.npz, rime_ylm.npz) and against the fp64 oracle.

CPU tests run the host logic over the byte-exact test double of the pack / GEMM entry points
(tests/cpu_double.py); ``-m gpu`` tests are the parity tests proper, through the C ABI.
"""
import ctypes

import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from bayeslim_b200 import ops, _lib
from oracle import rime_oracle as orc
from tests import oracle_cases as oc
from tests import model_cases as mc
from tests.cpu_double import emulated_kernels

TOL = {torch.float32: 1e-5, torch.float64: 1e-10}       # north-star tolerances


def relmax(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    return float((a.to(b.dtype) - b).abs().max() / b.abs().max())


def _alm_cases(g, dev, dtype):
    """(tag, AlmModel with its Ylm set, reference output, cotangent, reference gradient)"""
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    am = torch.as_tensor(g["alm_mult"]).to(dev, dtype)
    out = []
    for tag, real in (("complex", False), ("real", True)):
        A = ba.sph_harm.AlmModel(g["l"], g["m"], real_output=real)
        A.setup_Ylm(g["theta"], g["phi"], Ylm=torch.as_tensor(g["Ylm"]).to(dev, cd), alm_mult=am)
        out.append((tag, A))
    A = ba.sph_harm.AlmModel(g["l"], g["m"], real_output=True)
    A.setup_Ylm(g["theta_grid"], g["phi_grid"], separable=True, alm_mult=am,
                Ylm=(torch.as_tensor(g["Theta"]).to(dev, cd), torch.as_tensor(g["Phi"]).to(dev, cd)))
    out.append(("sep", A))
    return out


def _run_alm(g, dev, dtype, tol):
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    for tag, A in _alm_cases(g, dev, dtype):
        p = torch.as_tensor(g["params"]).to(dev, cd).requires_grad_()
        y = A(p)
        assert tuple(y.shape) == g["out_" + tag].shape and y.is_complex() == (tag == "complex")
        assert relmax(y, g["out_" + tag]) < tol, tag
        G = torch.as_tensor(g["G_" + tag]).to(dev, y.dtype)
        (torch.sum(G.real * y.real + G.imag * y.imag) if y.is_complex() else torch.sum(G * y)).backward()
        assert relmax(p.grad, g["grad_" + tag]) < 5 * tol, tag
    # multigrid: two pixel sets concatenated and re-indexed (sph_harm.py:1316-1335)
    A = ba.sph_harm.AlmModel(g["l"], g["m"], real_output=True)
    Y = torch.as_tensor(g["Ylm"]).to(dev, cd)
    am = torch.as_tensor(g["alm_mult"]).to(dev, dtype)
    idx = torch.arange(Y.shape[1] - 1, -1, -1, device=dev)
    A.setup_multigrid_forward([g["theta"][:60], g["theta"][60:]], [g["phi"][:60], g["phi"][60:]],
                              [Y[:, :60].contiguous(), Y[:, 60:].contiguous()], [am, am], idx=idx)
    y = A(torch.as_tensor(g["params"]).to(dev, cd))
    assert relmax(y, g["out_real"][..., ::-1].copy()) < tol


# ------------------------------------------------------------------ CPU: host logic
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_alm_model_host_logic(dtype):
    g = oc.load("alm_forward")
    with emulated_kernels() as calls:
        _run_alm(g, 'cpu', dtype, 2e-6 if dtype == torch.float32 else 1e-12)
    assert "cgemm" in calls
    assert ("cgemm_pack_b" in calls) == (dtype == torch.float32)


def test_rime_ylm_host_logic_float32_and_generate_mode():
    g = oc.load("rime_ylm")
    with emulated_kernels() as calls:
        rime, leaves = mc.build_ylm(g, 'cpu', torch.float32)
        V = rime().data
        assert relmax(V, g["vis"]) < 1e-5
        assert relmax(rime.beam.R.beam_cache, g["beam_cache"]) < 2e-6
        assert "cgemm" in calls and "build_interp_t" in calls or "build_interp" in calls
        # 'generate' mode: the harmonics are evaluated at the grid the response is called with
        rime2, _ = mc.build_ylm(g, 'cpu', torch.float64, mode='generate')
        R = rime2.beam.R
        th, ph = R.theta, R.phi
        b = R(rime2.beam.params, th, ph, rime2.beam.freqs)
        assert relmax(b, g["beam_cache"]) < 1e-12


def test_generated_ylm_matches_oracle_recurrence_and_scope():
    l, m = ba.sph_harm.gen_lm(15, real_field=False)
    rng = np.random.default_rng(3)
    th, ph = np.arccos(rng.uniform(-1, 1, 40)), rng.uniform(0, 2 * np.pi, 40)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        Y, norm, am = ba.sph_harm.gen_sph2pix(th, ph, l, m)
    finally:
        torch.set_default_dtype(old)
    assert np.abs(Y.numpy() - orc.sph_harm_matrix(l, m, th, ph)).max() < 1e-13
    assert float(am.max()) == 1.0           # negative orders present: no doubling
    with pytest.raises(NotImplementedError):
        ba.sph_harm.gen_sph2pix(th, ph, l + 0.5, m)
    with pytest.raises(NotImplementedError):
        ba.sph_harm.gen_sph2pix(th, ph, l, m, method='stripe', theta_crit=1.0)


def test_alm_has_no_cpu_path():
    A = ba.sph_harm.AlmModel(*ba.sph_harm.gen_lm(2))
    A.setup_Ylm(np.zeros(4), np.zeros(4), Ylm=torch.zeros(6, 4, dtype=torch.complex64))
    with pytest.raises(RuntimeError, match="CUDA"):
        A(torch.zeros(6, dtype=torch.complex64))


# ------------------------------------------------------------------ GPU: parity through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_alm_model_vs_reference(dtype):
    _run_alm(oc.load("alm_forward"), torch.device('cuda'), dtype, TOL[dtype])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(5, 37, 300), (130, 20, 129), (64, 528, 3000), (257, 100, 1000)])
def test_cgemm_ragged_shapes_vs_complex128(shape, dtype):
    """Edge cases of the blocking: rows / columns / modes that are not multiples of 128 / 128 / 16,
    a single tile with split k, real and complex operands on either side."""
    dev = torch.device('cuda')
    M, C, P = shape
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    gen = torch.Generator(device='cpu').manual_seed(M + C + P)
    rnd = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    for pc, yc, ro in [(True, True, True), (True, True, False), (False, True, True), (True, False, False)]:
        Y64 = torch.complex(rnd(C, P), rnd(C, P)) if yc else rnd(C, P)
        p64 = torch.complex(rnd(M, C), rnd(M, C)) if pc else rnd(M, C)
        p = p64.to(dev, cd if pc else dtype).requires_grad_()
        plan = ops.AlmPlan(Y64.to(dev, cd if yc else dtype))
        out = ops.alm_forward(p, plan, real_out=ro)
        pr = p.detach().cpu().to(p64.dtype).requires_grad_()
        Yr = Y64.to(cd if yc else dtype).to(Y64.dtype)       # the rounded operand
        ct = torch.complex128 if (pc or yc) else torch.float64
        ref = pr.to(ct) @ Yr.to(ct)
        ref = ref.real if ro else ref
        assert relmax(out, ref) < TOL[dtype], (shape, pc, yc, ro)
        G = torch.randn_like(out)
        out.backward(G)
        ref.backward(G.cpu().to(ref.dtype))
        assert relmax(p.grad, pr.grad) < TOL[dtype], (shape, pc, yc, ro)
        assert p.grad.dtype == p.dtype


@pytest.mark.gpu
def test_cgemm_operand_far_from_unit_scale_and_zero():
    """The power-of-two operand scaling keeps float16 hi/lo splits in range for any magnitude."""
    dev = torch.device('cuda')
    torch.manual_seed(1)
    for scale in (1e-12, 1e9):
        Y = torch.randn(90, 700, dtype=torch.complex64, device=dev) * scale
        p = torch.randn(40, 90, dtype=torch.complex64, device=dev) / scale
        out = ops.alm_forward(p, ops.AlmPlan(Y), real_out=False)
        ref = p.cpu().to(torch.complex128) @ Y.cpu().to(torch.complex128)
        assert relmax(out, ref) < 1e-5
    out = ops.alm_forward(torch.zeros(3, 90, dtype=torch.complex64, device=dev), ops.AlmPlan(Y))
    assert float(out.abs().max()) == 0.0


@pytest.mark.gpu
def test_alm_beam_map_size_nside64_against_fp64_on_a_subset_and_linearity():
    """The size the product is built for: 512 channel rows x lmax-60 modes (1891, m >= 0) x the
    49152 pixels of an nside-64 map.  The fp64 oracle checks a subset of rows / pixels; linearity
    and the adjoint identity <G, A p> = <A^H G, p> hold over the whole output."""
    dev = torch.device('cuda')
    l, m = ba.sph_harm.gen_lm(60, real_field=True)
    rng = np.random.default_rng(7)
    P = 49152
    theta = np.arccos(rng.uniform(0, 1, P))
    phi = rng.uniform(0, 2 * np.pi, P)
    rows_p = np.r_[0:64, P - 64:P]
    Ysub = torch.as_tensor(orc.sph_harm_matrix(l, m, theta[rows_p], phi[rows_p]))
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        Y, _, am = ba.sph_harm.gen_sph2pix(theta, phi, l, m, device=dev)
    finally:
        torch.set_default_dtype(old)
    A = ba.sph_harm.AlmModel(l, m, real_output=True)
    A.setup_Ylm(np.degrees(theta), np.degrees(phi), Ylm=Y, alm_mult=am.to(dev))
    g = torch.Generator().manual_seed(2)
    p = (torch.complex(torch.randn(512, len(l), generator=g), torch.randn(512, len(l), generator=g))
         / torch.as_tensor(1.0 + l)).to(dev).requires_grad_()
    y = A(p)
    rows_m = [0, 1, 255, 511]
    ref = orc.alm_forward(p.detach().cpu().to(torch.complex128)[rows_m], Ysub,
                          am.to(torch.float64), real_output=True)
    assert relmax(y[rows_m][:, rows_p], ref) < 1e-5
    G = torch.randn(y.shape, generator=g).to(dev)
    (G * y).sum().backward()
    lhs = float((G.double() * y.detach().double()).sum())
    rhs = float((p.grad.conj().to(torch.complex128) * p.detach().to(torch.complex128)).sum().real)
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    with torch.no_grad():
        y2 = A(2.5 * p.detach()) - 2.5 * y
    assert float(y2.abs().max() / y.detach().abs().max()) < 1e-5


@pytest.mark.gpu
def test_cgemm_c_abi_direct():
    """The C ABI as a maintainer's binding would call it: raw device pointers, no torch types."""
    dev = torch.device('cuda')
    M, K, N = 70, 50, 200
    torch.manual_seed(4)
    X = torch.randn(M, K, dtype=torch.complex64, device=dev)
    Y = torch.randn(N, K, dtype=torch.complex64, device=dev)           # rows = output columns
    one = torch.ones(1, dtype=torch.float32, device=dev) * 4096.0
    lib = _lib.lib
    Aq = torch.empty(lib.b200rime_cgemm_a_bytes(M, K), dtype=torch.uint8, device=dev)
    Bq = torch.empty(lib.b200rime_cgemm_b_bytes(N, K), dtype=torch.uint8, device=dev)
    out = torch.empty(M, N, dtype=torch.complex64, device=dev)
    part = torch.empty(2, M, N, dtype=torch.complex64, device=dev)
    P = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + off)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.b200rime_cgemm_pack_a_f32(P(X), P(X, 4), 2 * K, 2, M, K, P(one), 1, P(Aq), st) == 0
    assert lib.b200rime_cgemm_pack_b_f32(P(Y), P(Y, 4), 2 * K, 2, N, K, P(one), 0, P(Bq), st) == 0
    assert lib.b200rime_cgemm_f32(P(Aq), P(Bq), M, N, K, 2, 0, 0, P(one), P(one), P(out), N, P(part), st) == 0
    torch.cuda.synchronize()
    ref = X.cpu().to(torch.complex128) @ Y.cpu().to(torch.complex128).T
    assert relmax(out, ref) < 1e-5
    # errors come back as a status and a message, not as an exception or a crash
    assert lib.b200rime_cgemm_f32(P(Aq), P(Bq), M, N, K, 99, 0, 0, P(one), P(one), P(out), N, P(part), st) != 0
    assert b"ksplit" in lib.b200rime_last_error()
