"""
On-device eq2top (reference telescope_model.py:469-502).  Parity with
astropy is UNPINNED (astropy is absent and the reference holds no golden values for it): the
product's rotation-matrix form is checked against the oracle's independent spherical-trigonometry
derivation (Meeus ch. 21-23), against known answers, and the CUDA kernel against the host form.
"""
import numpy as np
import pytest
import torch

import bayeslim_b200 as ba
from bayeslim_b200 import telescope_model as tm
from oracle import rime_oracle as orc

LOC = (21.42827, -30.72148, 1051.7)


def _sky(n, seed=0, dec_max=70.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 360, n), rng.uniform(-dec_max, dec_max, n)


def _angsep(z1, a1, z2, a2):
    """Great-circle distance [deg] between two (zen, az) directions."""
    d2r = np.pi / 180
    v = lambda z, a: np.stack([np.sin(z * d2r) * np.sin(a * d2r), np.sin(z * d2r) * np.cos(a * d2r),
                               np.cos(z * d2r)])
    # from the chord, not arccos of the dot product (whose rounding alone is 1e-6 deg near 0)
    chord = np.sqrt(((v(z1, a1) - v(z2, a2)) ** 2).sum(0))
    return np.degrees(2 * np.arcsin(np.clip(chord / 2, 0, 1)))


@pytest.mark.parametrize("jd", [2451545.0, 2458148.2, 2460676.75])
def test_matrix_form_matches_trigonometric_oracle(jd):
    ra, dec = _sky(2000)
    zen, az = tm.eq2top_apparent(LOC, jd, ra, dec)
    zo, ao = orc.eq2top_apparent(jd, ra, dec, LOC[0], LOC[1])
    # two derivations, second-order agreement in the 20-arcsecond corrections (tan(dec) <= 2.7)
    assert _angsep(zen, az, zo, ao).max() < 2e-6          # < 8 milliarcseconds
    m9, v3 = tm.icrs_to_enu(LOC, jd)
    assert np.abs(m9 @ m9.T - np.eye(3)).max() < 1e-14 and abs(np.linalg.det(m9) - 1) < 1e-14
    assert abs(np.linalg.norm(v3) / (20.49552 * np.pi / 180 / 3600) - 1) < 1e-12


def test_known_answers():
    # at J2000.0 precession vanishes: the apparent place differs from the catalogue place only by
    # nutation (<= 17.3 arcsec) and aberration (<= 20.5 arcsec)
    ra, dec = _sky(500, seed=1)
    z1, a1 = tm.eq2top_apparent(LOC, 2451545.0, ra, dec)
    z0, a0 = tm.eq2top_rigid(LOC, 2451545.0, ra, dec)
    sep = _angsep(z1, a1, z0, a0)
    assert sep.max() < 40.0 / 3600 and sep.max() > 5.0 / 3600
    # 18 years later general precession has moved the frame by 50.3 arcsec / yr
    z1, a1 = tm.eq2top_apparent(LOC, 2458148.2, ra, dec)
    z0, a0 = tm.eq2top_rigid(LOC, 2458148.2, ra, dec)
    yrs = (2458148.2 - 2451545.0) / 365.25
    assert _angsep(z1, a1, z0, a0).max() < (50.3 * yrs + 45) / 3600
    # the celestial pole of date sits at altitude |lat| due south for a southern site
    m9, _ = tm.icrs_to_enu(LOC, 2458148.2)
    T = (2458148.2 - 2451545.0) / 36525.0
    # pole of date in J2000 coordinates: precession angle theta from the J2000 pole
    e, n, u = m9 @ np.array([0.0, 0.0, 1.0])
    alt = np.degrees(np.arcsin(u))
    assert abs(alt - LOC[1]) < 0.15              # the J2000 pole is 0.1 deg from the pole of date


def test_telescope_model_uses_the_apparent_form_without_astropy():
    if tm._have_astropy():
        pytest.skip("astropy present: the reference's own conversion is used")
    ra, dec = _sky(50, seed=2)
    tel = ba.telescope_model.TelescopeModel(LOC)
    with pytest.warns(UserWarning) if not tm._WARNED_NO_ASTROPY else _nullcontext():
        za = tel.eq2top(2458148.2, ra, dec, store=True)
    zen, az = tm.eq2top_apparent(LOC, 2458148.2, ra, dec)
    assert np.allclose(za[0].numpy(), zen) and np.allclose(za[1].numpy(), az)
    assert tel.hash(2458148.2, ra) in tel.conv_cache


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


@pytest.mark.gpu
def test_eq2top_kernel_matches_host_form():
    ra, dec = _sky(200000, seed=3, dec_max=89.9)
    for jd in (2451545.0, 2458148.2):
        zen, az = tm.eq2top_apparent(LOC, jd, ra, dec)
        out = tm.eq2top_device(LOC, jd, ra, dec, torch.device('cuda'))
        assert out.shape == (2, len(ra)) and out.dtype == torch.float64
        sep = _angsep(out[0].cpu().numpy(), out[1].cpu().numpy(), zen, az)
        assert sep.max() < 1e-9
    tel = ba.telescope_model.TelescopeModel(LOC, device='cuda')
    if not tm._have_astropy():
        za = tel.eq2top(2458148.2, torch.as_tensor(ra[:1000]), torch.as_tensor(dec[:1000]))
        assert za.is_cuda and float((za[0].cpu() - torch.as_tensor(zen[:1000])).abs().max()) < 1e-9
