"""
Oracle stage 2 (LOS).  The reference holds no LOS arithmetic (hammurabiX is external), so
these are the analytic pins the oracle's definition rests on (SURVEY.md A.4 self-checks).
"""
import numpy as np
import pytest

from oracle import fields as F
from oracle import healpix as H
from oracle import los as L


def test_healpix_documented_values():
    """healpy docs: pix2ang(16, [1440, 427, 1520, 0]) (RING)."""
    th, ph = H.pix2ang_ring(16, [1440, 427, 1520, 0])
    assert np.allclose(th, [1.5291175943723188, 0.78550497, 1.57079633, 0.05103658], atol=5e-9)
    assert np.allclose(ph, [0., 0.78539816, 1.61988371, 0.78539816], atol=5e-9)


@pytest.mark.parametrize('nside', [1, 2, 4, 8, 32])
def test_healpix_symmetries(nside):
    v = H.pix2vec_ring(nside)
    npix = 12 * nside * nside
    assert v.shape == (npix, 3)
    assert np.abs(np.linalg.norm(v, axis=1) - 1).max() < 4e-16
    assert np.abs(v.sum(0)).max() < 1e-11
    # north/south mirror: pixel p <-> ring-mirrored pixel has z -> -z
    z = v[:, 2]
    assert np.allclose(np.sort(z), np.sort(-z), atol=1e-15)
    # rings are ordered north to south, phi increasing within a ring
    assert np.all(np.diff(z) <= 1e-15)
    # equal-area: 4*pi/npix per pixel => ring z values step by known amounts at the equator
    ze = np.unique(np.round(z, 13))
    assert len(ze) == 4 * nside - 1


def _uniform_setup(nside=4, p=3.0, By=2.0, Bx=0.0, Bz=0.0, ne=0.01):
    g = F.Grid([[-20, 20], [-20, 20], [-4, 4]], [21, 21, 9])
    dirs = H.pix2vec_ring(nside)
    cfg = L.LosConfig(alpha=p)
    B = F.constant_B(g, Bx, By, Bz)[None]
    return g, dirs, cfg, B, F.constant_ne(g, ne)[None], F.cre_const(g, 0.0217)[None]


def test_uniform_medium_closed_forms():
    g, dirs, cfg, B, ne, cre = _uniform_setup()
    out = L.integrate(g, cfg, dirs, B, ne, cre)
    _, inside, _, _ = L.sample_geometry(g, cfg, dirs)
    path_pc = inside.sum(1) * cfg.dr * 1e3
    assert np.allclose(out['dm'][0], 0.01 * path_pc, rtol=1e-13)
    bpar = -(dirs @ np.array([0., 2., 0.]))
    assert np.allclose(out['fd'][0], 0.812 * 0.01 * bpar * path_pc, rtol=1e-12, atol=1e-12)
    # sign convention: field pointing toward the observer => RM > 0
    toward = dirs[:, 1] < -0.3
    assert np.all(out['fd'][0][toward] > 0)


def test_polarisation_fraction_without_faraday_rotation():
    for p in (2.4, 3.0, 3.6):
        g, dirs, cfg, B, ne, cre = _uniform_setup(p=p)
        out = L.integrate(g, cfg, dirs, B, None, cre)      # no ne => no Faraday rotation
        pi = np.sqrt(out['Q']**2 + out['U']**2)
        ok = out['I'][0] > 0
        assert np.allclose((pi / out['I'])[0][ok], (p + 1) / (p + 7. / 3.), rtol=1e-12)


def test_polarisation_angle_iau_convention():
    # B along +z (Galactic north) on the plane: E-vector is east-west => chi = 90deg,
    # Q = -P, U = 0.  B along +y seen toward l = 0: psi_B = +90deg (east) => chi = 180 => Q = +P.
    g, dirs, cfg, B, ne, cre = _uniform_setup(Bx=0, By=0, Bz=2.0)
    d = np.array([[1.0, 0, 0]])
    out = L.integrate(g, cfg, d, B, None, cre)
    assert out['Q'][0, 0] < 0 and abs(out['U'][0, 0]) < 1e-18
    g, dirs, cfg, B, ne, cre = _uniform_setup(By=2.0)
    out = L.integrate(g, cfg, d, B, None, cre)
    assert out['Q'][0, 0] > 0 and abs(out['U'][0, 0]) < 1e-18
    # B at 45 deg from north toward east: psi = 45 => chi = 135 => Q = 0, U = -P
    g, dirs, cfg, B, ne, cre = _uniform_setup(By=1.0, Bz=1.0)
    out = L.integrate(g, cfg, d, B, None, cre)
    assert abs(out['Q'][0, 0]) < 1e-16 * abs(out['U'][0, 0]) + 1e-30 and out['U'][0, 0] < 0


def test_faraday_depolarisation_closed_form():
    """Uniform slab: Q+iU = P0 * mean_j exp(2 i phi_j lam^2), phi_j = (j+1/2) f."""
    g, dirs, cfg, B, ne, cre = _uniform_setup(Bx=1.5, By=2.0, ne=0.05)
    cfg = L.LosConfig(nu_ghz=1.4)
    d = np.array([[1.0, 0, 0]])
    out = L.integrate(g, cfg, d, B, ne, cre)
    out0 = L.integrate(g, cfg, d, B, None, cre)
    _, inside, _, _ = L.sample_geometry(g, cfg, d)
    J = int(inside.sum())
    f = 0.812 * 0.05 * (-1.5) * cfg.dr * 1e3
    ang = 2 * (np.arange(J) + 0.5) * f * L.lambda_sq_m2(1.4)
    rot = np.exp(1j * ang).mean()
    got = (out['Q'][0, 0] + 1j * out['U'][0, 0])
    want = (out0['Q'][0, 0] + 1j * out0['U'][0, 0]) * rot
    assert abs(got - want) <= 1e-12 * abs(want)
    assert out['I'][0, 0] == out0['I'][0, 0]


def test_sync_coefficient_power_law_in_frequency():
    k1, k2 = L.sync_coefficient(3.0, 1.0, 20.6), L.sync_coefficient(3.0, 10.0, 20.6)
    # T_b ~ nu^{-(p+3)/2}
    assert k2 / k1 == pytest.approx(10.0 ** (-3.0), rel=1e-12)
    assert 1e-27 < L.sync_coefficient(3.0, 23.0, 20.6) < 1e-25


def test_trilinear_reproduces_linear_fields_and_box_mask():
    g = F.Grid([[-20, 20], [-20, 20], [-4, 4]], [17, 13, 11])
    x, y, z = g.xyz()
    lin = (0.3 * x - 0.2 * y + 1.7 * z + 5.0) * np.ones(g.shape)
    dirs = H.pix2vec_ring(2)
    cfg = L.LosConfig(dr=0.37, r_max=30)
    pos, inside, i0, t = L.sample_geometry(g, cfg, dirs)
    val = L._trilinear(lin, i0, t)
    want = 0.3 * pos[..., 0] - 0.2 * pos[..., 1] + 1.7 * pos[..., 2] + 5.0
    assert np.allclose(val[inside], want[inside], rtol=1e-13, atol=1e-13)
    lo, hi = g.box[:, 0], g.box[:, 1]
    geo_in = np.all((pos >= lo) & (pos <= hi), axis=-1)
    assert np.array_equal(geo_in, inside)
    assert L.in_box_samples(g, cfg, dirs) == int(inside.sum())


def test_tutorial_hammurabi_pin():
    """
    The one hammurabiX OUTPUT recorded in the reference repository (tutorials/tutorial_hammurabi.ipynb, cell 14):
    Stokes I at 23 GHz, Nside 32, of the built-in WMAP LSA field + analytic CRE model, first and last three
    RING pixels.  The oracle's emissivity definition, evaluated with oracle/fields.py's LSA and CRE formulas
    directly at the ray samples, reproduces all six to 3e-5 once Rayleigh-Jeans brightness temperature is
    expressed as thermodynamic CMB temperature (the reference's "K-cmb", observable_dict.py:15-19) -- which pins
    the CRE normalisation from (E0, j0), the K(p) prefactor, the LSA and CRE model definitions, the observer and
    the integration limits (test_ham.xml:122-131), and settles K_RJ vs K_CMB: without the conversion every
    pixel is low by the same 1.36 %.
    """
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import hammurabi_tutorial_pin as pin
    from oracle import los as L
    npix = 12 * 32 ** 2
    pix = [k % npix for k in pin.NOTEBOOK_I_MK]
    want = np.array(list(pin.NOTEBOOK_I_MK.values()))
    ours_rj = pin.stokes_i_mk(pix)
    ratio_rj = ours_rj / want
    assert np.ptp(ratio_rj) < 2e-7                               # one constant factor for all six pixels ...
    assert abs(ratio_rj.mean() - 0.98641) < 2e-5                 # ... 1.36 % low as Rayleigh-Jeans temperature
    ours_cmb = ours_rj * L.rj_to_cmb(23.0)
    assert np.abs(ours_cmb / want - 1.0).max() < 5e-5            # and equal as K_CMB
    # the limits matter at a level the pin resolves: |z| <= 4 kpc or r_sun = 8.5 kpc would be off by 0.2 % / 2 %
    assert abs((pin.stokes_i_mk(pix, gc_z_max=4.0) * L.rj_to_cmb(23.0) / want).mean() - 1.0) > 1e-3
    assert abs((pin.stokes_i_mk(pix, r_sun=8.5) * L.rj_to_cmb(23.0) / want).mean() - 1.0) > 1e-2
