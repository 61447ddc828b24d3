"""
Pin the CPU oracle (oracle/rime_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) and against the known-answer
checks the reference's own tests hold for this path.
Runs on CPU in float64, like every reference test (tests/test_rime.py:5).
"""
import numpy as np
import pytest
import torch

from oracle import rime_oracle as orc
from tests import oracle_cases as oc

torch.set_default_dtype(torch.float64)


def relmax(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float((a.detach() - b).abs().max() / b.abs().max())


# ------------------------------------------------------------------ pieces
def test_make_hex_matches_reference_positions():
    g = oc.load("fringe")
    ants, vecs = orc.make_hex(3, D=15)
    assert len(ants) == 19
    assert np.abs(vecs - g["antvecs"]).max() < 1e-12
    # tests/test_telescope.py:50: blvec (0 -> 1) = [15, 0, 0]
    bl = orc.get_blvecs(vecs, ants, [(0, 1)])
    assert np.allclose(bl.numpy(), [[15.0, 0, 0]], atol=1e-12)


def test_fringe_golden_and_known_answers():
    g = oc.load("fringe")
    blvecs = torch.as_tensor(g["blvecs"])
    f = orc.gen_fringe(blvecs, g["zen"], g["az"], g["freqs"])
    fc = orc.gen_fringe(blvecs, g["zen"], g["az"], g["freqs"], conj=True)
    assert f.shape == g["fringe"].shape and f.dtype == torch.complex128
    assert relmax(f, g["fringe"]) < 1e-13
    assert relmax(fc, g["fringe_conj"]) < 1e-13
    # tests/test_telescope.py:73,75 : conj flag; fringe == 1 at zenith; |fringe| <= 1
    assert (fc - f.conj()).abs().max() < 1e-14
    assert (f[:, :, 0] - 1.0).abs().max() < 1e-12
    assert (f.abs() - 1).abs().max() < 1e-12
    # autocorrelation baseline (0,0) has b = 0 -> fringe 1
    assert (f[1] - 1.0).abs().max() == 0


def test_airy_golden():
    g = oc.load("airy")
    for nm, pb in [("sym_power", True), ("asym_power", True), ("asym_volt", False),
                   ("twopol_power", True)]:
        out = orc.airy_response(torch.as_tensor(g["params_" + nm]), g["zen"], g["az"],
                                torch.as_tensor(g["freqs"]), powerbeam=pb)
        assert out.shape == g["beam_" + nm].shape
        assert relmax(out, g["beam_" + nm]) < 1e-13
    out = orc.airy_response(torch.as_tensor(g["params_sym_power"]), g["zen"], g["az"],
                            torch.as_tensor(g["freqs"]), freq_ratio=1.1)
    assert relmax(out, g["beam_ratio"]) < 1e-13


@pytest.mark.parametrize("mode", ['nearest', 'linear', 'quadratic', 'cubic', 'linear,quadratic'])
def test_rect_interp_golden(mode):
    g = oc.load("rect_interp")
    key = mode.replace(',', '_')
    inds, wgts = orc.rect_interp_weights(g["theta_grid"], g["phi_grid"], g["zen"], g["az"], mode)
    assert inds.shape == g["inds_" + key].shape
    out = orc.interp_map(torch.as_tensor(g["m"]), inds, wgts)
    # operator equality (node choice may differ at exact ties, the interpolant may not).
    # The reference obtains the weights from pinv(A^T A) A^T of a monomial design matrix
    # (utils.py:1084-1116); cond(A^T A) reaches ~1e21 for the 16 cubic monomials, so the
    # reference's own weights carry ~2e-10 (quadratic) .. 1e-5 (cubic) noise.  The oracle's
    # Lagrange form is the exact interpolant, hence the graded tolerance.
    tol = {'nearest': 1e-13, 'linear': 1e-12, 'cubic': 1e-4}.get(mode, 1e-8)
    assert (out - torch.as_tensor(g["interp_" + key])).abs().max() < tol
    same = (inds.numpy() == g["inds_" + key]).all(axis=1)
    assert same.mean() > 0.9
    assert np.abs(wgts.numpy()[same] - g["wgts_" + key][same]).max() < tol


def test_rect_interp_of_airy_close_to_analytic():
    # tests/test_beam.py:46-63: interpolated rect-grid Airy vs analytic Airy, std < 1e-3
    freqs = torch.linspace(120e6, 130e6, 10)
    theta = torch.arange(0, 90.1, 1.0)
    phi = torch.arange(0, 360, 1.0)
    b_phi, b_theta = torch.meshgrid(phi, theta, indexing='xy')
    m = orc.airy_disk(b_theta.ravel() * orc.D2R, b_phi.ravel() * orc.D2R, 10.0, freqs)
    az, zen = torch.meshgrid(torch.arange(0, 360, 10.0), torch.arange(0, 90, 2.5), indexing='ij')
    az, zen = az.ravel(), zen.ravel()
    inds, wgts = orc.rect_interp_weights(theta, phi, zen, az, 'linear')
    out1 = orc.interp_map(m, inds, wgts)
    out2 = orc.airy_disk(zen * orc.D2R, az * orc.D2R, 10.0, freqs)
    assert (out1 - out2).std() < 1e-3


def test_point_sky_powerlaw():
    # tests/test_sky.py:42-48
    freqs = torch.linspace(120e6, 130e6, 10)
    params = torch.ones(1, 1, 2, 10)
    params[..., 1, :] = -2.2
    data = orc.point_sky_response(params, freqs, 'powerlaw', f0=freqs[0])
    assert data.shape == (1, 1, 10, 10)
    assert data.isclose((freqs[:, None] / freqs[0]) ** -2.2).all()


def test_healpix_restatement_self_consistency():
    # parity unpinned (healpy absent): closed-form checks only
    for nside in (1, 2, 4, 8):
        theta, phi = orc.healpix_pix2ang(nside)
        npix = orc.healpix_npix(nside)
        assert len(theta) == npix
        # pixel centres integrate z to ~0 and sum of areas = 4 pi
        assert abs(np.cos(theta).sum()) < 1e-9
        assert abs(orc.healpix_pixarea(nside) * npix - 4 * np.pi) < 1e-12
        inds, wgts = orc.healpix_interp_weights(nside, theta, phi)
        assert (wgts >= -1e-12).all() and np.allclose(wgts.sum(1).numpy(), 1.0)
        # exact at pixel centres
        m = torch.randn(npix)
        assert (orc.interp_map(m, inds, wgts) - m).abs().max() < 1e-9
    rng = np.random.default_rng(0)
    th = np.arccos(rng.uniform(-1, 1, 500))
    ph = rng.uniform(0, 2 * np.pi, 500)
    inds, wgts = orc.healpix_interp_weights(8, th, ph)
    assert (wgts >= -1e-12).all() and np.allclose(wgts.sum(1).numpy(), 1.0)
    # a smooth function is reproduced to O(pixel^2)
    t8, p8 = orc.healpix_pix2ang(8)
    f = lambda t, p: np.cos(t) + 0.3 * np.sin(t) * np.cos(p)
    m = torch.as_tensor(f(t8, p8))
    assert np.abs(orc.interp_map(m, inds, wgts).numpy() - f(th, ph)).max() < 2e-2


# ------------------------------------------------------------------ full RIME vs reference
def _check(V, leaves, g, grads):
    assert tuple(V.shape) == g["vis"].shape
    assert relmax(V, g["vis"]) < 1e-12
    oc.real_loss(V, g["G"]).backward()
    for k, gk in grads.items():
        assert relmax(leaves[k].grad, g[gk]) < 1e-10, (k, gk)


def test_rime_point_airy_vs_reference():
    g = oc.load("rime_point_airy")
    V, leaves = oc.oracle_point_airy(g)
    _check(V, leaves, g, dict(sky="grad_sky", beam="grad_beam_truncated", antvecs="grad_antvecs"))


def test_rime_airy_brute_force_vs_reference():
    """AiryResponse(brute_force=True): trapezoid Bessel integral (special.py:498-533); the
    gradient to the dish diameter is the full one (no truncation at bessel_j1)."""
    g = oc.load("rime_airy_brute")
    V, leaves = oc.oracle_point_airy(g, brute_force=True)
    _check(V, leaves, g, dict(sky="grad_sky", beam="grad_beam"))
    assert np.abs(g["grad_beam"]).max() > 0


def test_alm_forward_vs_reference():
    """AlmModel.forward_alm (sph_harm.py:1289-1373): full / separable, complex / real output."""
    g = oc.load("alm_forward")
    Y, am = torch.as_tensor(g["Ylm"]), torch.as_tensor(g["alm_mult"])
    # the oracle's own Ylm recurrence reproduces the reference's gen_sph2pix (m >= 0)
    Yo = orc.sph_harm_matrix(g["l"], g["m"], np.radians(g["theta"]), np.radians(g["phi"]))
    assert np.abs(Yo - g["Ylm"]).max() < 1e-13
    for tag, real in (("complex", False), ("real", True)):
        p = oc.tt(g["params"], torch.complex128, grad=True)
        y = orc.alm_forward(p, Y, am, real_output=real)
        assert relmax(y, g["out_" + tag]) < 1e-13
        G = torch.as_tensor(g["G_" + tag])
        (oc.real_loss(y, G) if not real else torch.sum(G * y)).backward()
        assert relmax(p.grad, g["grad_" + tag]) < 1e-12
    p = oc.tt(g["params"], torch.complex128, grad=True)
    y = orc.alm_forward_separable(p, torch.as_tensor(g["Theta"]), torch.as_tensor(g["Phi"]), am,
                                  real_output=True)
    assert relmax(y, g["out_sep"]) < 1e-13
    torch.sum(torch.as_tensor(g["G_sep"]) * y).backward()
    assert relmax(p.grad, g["grad_sep"]) < 1e-12


def test_rime_ylm_vs_reference():
    """YlmResponse in interpolate mode through the RIME (beam_model.py:1019-1267)."""
    g = oc.load("rime_ylm")
    assert np.abs(oc.ylm_grid_matrix(g).numpy()[:, ::23] - g["Ylm_sample"]).max() < 1e-13
    V, leaves, beam_cache = oc.oracle_ylm(g)
    assert relmax(beam_cache, g["beam_cache"]) < 1e-12
    _check(V, leaves, g, dict(sky="grad_sky", beam="grad_beam", antvecs="grad_antvecs"))


def test_rime_alm_sky_vs_reference():
    """PixelSkyResponse(spatial_mode='alm') through the RIME (sky_model.py:510-732)."""
    g = oc.load("rime_alm_sky")
    V, leaves, Ylm = oc.oracle_alm_sky(g)
    assert np.abs(Ylm.numpy()[:, ::17] - g["Ylm_sample"]).max() < 1e-13
    _check(V, leaves, g, dict(sky="grad_sky", antvecs="grad_antvecs"))


def test_rime_pixel_interp_vs_reference():
    g = oc.load("rime_pixel_interp")
    V, leaves = oc.oracle_pixel_interp(g)
    _check(V, leaves, g, dict(sky="grad_sky", beam="grad_beam", antvecs="grad_antvecs"))


def test_rime_pointing_offset_vs_reference():
    """Pointing offset (beam_model.py:244-256): the beam is evaluated at the rotated directions,
    the fringe at the true ones; interpolated pixel beam and Airy beam."""
    g = oc.load("rime_pointing")
    V, leaves = oc.oracle_pointing(g, 'interp')
    assert relmax(V, g["vis_interp"]) < 1e-12
    oc.real_loss(V, g["G"]).backward()
    for k in ("sky", "beam", "antvecs"):
        assert relmax(leaves[k].grad, g["grad_%s_interp" % k]) < 1e-10, k
    V, leaves = oc.oracle_pointing(g, 'airy')
    assert relmax(V, g["vis_airy"]) < 1e-12
    oc.real_loss(V, g["G"]).backward()
    assert relmax(leaves["sky"].grad, g["grad_sky_airy"]) < 1e-10
    # and the offset matters: without it the visibilities differ at the per-cent level
    g0 = dict(g)
    g0["offset"] = (0.0, 0.0)
    with torch.no_grad():
        V0, _ = oc.oracle_pointing(g0, 'interp')
    assert relmax(V0, g["vis_interp"]) > 1e-3


def test_rime_batched_vs_reference():
    g = oc.load("rime_batched")
    with torch.no_grad():
        V, _ = oc.oracle_pixel_interp(g, interp_mode='quadratic', grad=False)
    # quadratic weights: reference pinv noise ~2e-10 (see test_rect_interp_golden)
    assert relmax(V, g["vis"]) < 1e-9


def test_rime_2pol_vs_reference():
    g = oc.load("rime_2pol")
    V, leaves = oc.oracle_2pol(g)
    _check(V, leaves, g, dict(sky="grad_sky"))


def test_rime_4pol_vs_reference():
    g = oc.load("rime_4pol")
    V, leaves = oc.oracle_4pol(g)
    _check(V, leaves, g, dict(sky="grad_sky", beam="grad_beam", antvecs="grad_antvecs"))


def test_rime_multimodel_vs_reference():
    g = oc.load("rime_multimodel")
    V, leaves = oc.oracle_multimodel(g)
    _check(V, leaves, g, dict(sky="grad_sky"))


def test_rime_databls_vs_reference():
    g = oc.load("rime_databls")
    V, leaves = oc.oracle_databls(g)
    _check(V, leaves, g, dict(sky="grad_sky"))
