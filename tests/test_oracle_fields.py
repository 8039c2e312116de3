"""Oracle stage 1 vs the reference's tutorial goldens."""
import numpy as np
import pytest

from oracle import fields as F


def test_grid_mgrid_semantics():
    g = F.Grid([[0, 4], [1, 2], [1, 1]], [5, 2, 1])          # grid.py docstring example
    assert np.array_equal(g.axis_coords(0), [0, 1, 2, 3, 4])
    assert np.array_equal(g.axis_coords(1), [1, 2])
    assert np.array_equal(g.axis_coords(2), [1])
    box = [[-15, 15], [-15, 15], [-15, 15]]
    g = F.Grid(box, [15, 15, 15])
    ref = np.mgrid[tuple(slice(lo, hi, 15j) for lo, hi in box)]
    for a in range(3):
        assert np.array_equal(np.moveaxis(ref[a], a, 0)[:, 0, 0], g.axis_coords(a))


def test_exponential_ne_tutorial_golden():
    """tutorials/tutorial_fields.ipynb:249-256."""
    g = F.Grid([[-15, 15]] * 3, [15, 15, 15])
    ne = F.exponential_ne(g, 1.0, 3.3, 3.3)
    want = [0.00189093, 0.00361981, 0.00289678, 0.0055453, 0.00289678, 0.0055453,
            0.00473694, 0.00906791]
    assert np.allclose(ne[3:5, 3:5, 3:5].ravel(), want, rtol=0, atol=5e-9)


def test_r_spherical_tutorial_golden():
    """tutorials/tutorial_fields.ipynb:93."""
    g = F.Grid([[-15, 15]] * 3, [15, 15, 15])
    x, y, z = g.xyz()
    r = np.sqrt((x**2 + y**2) + z**2) * np.ones(g.shape)
    assert r[5, 5, 5] == pytest.approx(7.423074889580904, abs=1e-15)


def test_mock_data_golden_seed_233():
    """tutorials/tutorial_one.ipynb:91-103 / examples/basic_pipeline.py:82-92."""
    x = np.linspace(0.01, 2. * np.pi - 0.01, 10)
    np.random.seed(233)
    signal = (1 + np.cos(x)) * np.random.normal(loc=3., scale=6., size=10)
    fd = signal + np.random.normal(loc=0., scale=0.1, size=10)
    assert np.allclose(fd[:4], [16.4217790817552, 7.172468731201507, -3.2254947821460433,
                                0.27949334758966465], rtol=0, atol=1e-13)


def test_reference_rng_equals_scipy_norm_rvs():
    import scipy.stats as stats
    g = F.Grid([[0, 1]] * 3, [4, 3, 2])
    want = stats.norm(loc=1.5, scale=0.3).rvs(size=g.shape + (3,), random_state=77)
    assert np.array_equal(F.gaussian_B_reference(g, 1.5, 0.3, 77), want)


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = F.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_philox_normals_moments_and_independence():
    z = F.philox_normals(987654321, 0, 1_000_001)
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1) < 3e-3
    assert abs((z**3).mean()) < 1e-2 and abs((z**4).mean() - 3) < 3e-2
    z2 = F.philox_normals(987654321, 1, 100_000)
    assert abs(np.corrcoef(z[:100_000], z2)[0, 1]) < 1e-2
    z3 = F.philox_normals(987654322, 0, 100_000)
    assert abs(np.corrcoef(z[:100_000], z3)[0, 1]) < 1e-2


def test_sum_rule_and_realisation_axis():
    g = F.Grid([[-1, 1]] * 3, [4, 5, 6])
    descs = [(F.CONST_B, [0., 1., 0.], 0), (F.GAUSS_B, [0.5, 2.0], 1),
             (F.EXP_NE, [0.01, 3., .1], 2), (F.COS_NE, [1, 1, .1, 2, .2, 3, .3], 3)]
    out = F.evaluate_fields(g, descs, [5, 6, 7])
    assert out['magnetic_field'].shape == (3, 4, 5, 6, 3)
    assert out['thermal_electron_density'].shape == (1, 4, 5, 6)
    want = F.constant_B(g, 0, 1, 0) + F.gaussian_B_philox(g, 0.5, 2.0, 6, 1)
    assert np.array_equal(out['magnetic_field'][1], want)


def test_test_simulator_matches_scipy_nearest():
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(0)
    g = F.Grid([[0, 2 * np.pi], [0, 0], [-1, 1]], [100, 1, 7])
    B = rng.normal(size=g.shape + (3,))
    ne = rng.random(g.shape)
    pts = np.stack([rng.uniform(0, 2 * np.pi, 50), np.zeros(50), rng.uniform(-1, 1, 50)], 1)
    # exact ties and end points
    c = g.axis_coords(0)
    pts[0, 0], pts[1, 0], pts[2, 0] = 0.5 * (c[3] + c[4]), c[0], c[-1]
    interp = RegularGridInterpolator(tuple(g.axis_coords(a) for a in range(3)),
                                     B[..., 1] * ne, method='nearest')
    assert np.array_equal(F.test_simulator_sample(g, B, ne, pts, 1), interp(pts))
    with pytest.raises(ValueError):
        F.test_simulator_sample(g, B, ne, [[7.0, 0, 0]], 1)
