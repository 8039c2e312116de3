"""GPU parity, stage 1: field kernels, HEALPix directions, TestSimulator sampling vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fields as OF          # noqa: E402
from oracle import healpix as OH         # noqa: E402
from tests.helpers import b_to_reference_layout   # noqa: E402


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    from imagine_b200 import engine, _lib
    _lib.init(0)
    return engine


BOX = [[-20, 20], [-20, 20], [-4, 4]]


def _gpu_fields(eng, box, res, descs, seeds, bs):
    grid = eng.DeviceGrid(box, res)
    out = eng.field_eval(grid, descs, seeds, b_stride=bs)
    res = tuple(int(r) for r in res)
    host = {}
    for k, v in out.items():
        if k == 'magnetic_field':
            a = b_to_reference_layout(v, len(seeds), res)
            if v.dim() == 3 and bs == 4:
                assert bool((v[..., 3] == 0.0).all())
        else:
            a = v.cpu().numpy()
            a = a.reshape(a.shape[0], *res)
        host[k] = a
    return host


@pytest.mark.parametrize('bs', [3, 4])
@pytest.mark.parametrize('res', [[16, 12, 10], [7, 5, 3], [30, 1, 1], [64, 64, 64]])
def test_deterministic_fields_match_oracle(eng, res, bs):
    """north_star: deterministic fields within 1e-9 relative in FP64."""
    box = BOX if res[1] > 1 else [[0, 2 * np.pi], [0, 0], [0, 0]]
    descs = [(OF.CONST_B, [0.3, -1.0, 2.5], 0), (OF.CONST_B, [1.0, 1.0, 1.0], 0),
             (OF.CONST_NE, [0.02], 0), (OF.EXP_NE, [0.01, 3.0, 0.1], 0),
             (OF.COS_NE, [1.0, 1.0, 0.1, 0.5, np.pi / 2, 0.25, np.pi / 2], 0),
             (OF.CRE_ANALYTIC, [0.0217, 5.0, 1.0, 8.3], 0), (OF.CRE_CONST, [0.01], 0)]
    got = _gpu_fields(eng, box, res, descs, [1], bs)
    want = OF.evaluate_fields(OF.Grid(box, res), descs, [1])
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape
        assert np.allclose(got[k], want[k], rtol=1e-9, atol=0), k
    assert np.array_equal(got['magnetic_field'], want['magnetic_field'])   # constants are exact


def test_tutorial_golden_on_gpu(eng):
    """tutorials/tutorial_fields.ipynb:249-256 through the CUDA kernel."""
    got = _gpu_fields(eng, [[-15, 15]] * 3, [15, 15, 15], [(OF.EXP_NE, [1.0, 3.3, 3.3], 0)], [1], 3)
    want = [0.00189093, 0.00361981, 0.00289678, 0.0055453, 0.00289678, 0.0055453, 0.00473694, 0.00906791]
    assert np.allclose(got['thermal_electron_density'][0, 3:5, 3:5, 3:5].ravel(), want, rtol=0, atol=5e-9)


@pytest.mark.parametrize('bs', [3, 4, 24, 48, 96])
@pytest.mark.parametrize('res', [[12, 10, 8], [5, 3, 3], [9, 9, 9]])
def test_philox_gaussian_fields_match_cpu_philox(eng, res, bs):
    """Counter-based Philox: the CPU restatement reproduces the GPU stream (odd and even ncell)."""
    seeds = [12345, 2 ** 31 + 7, 1431655764, 2 ** 40 + 3, 99, 7]
    descs = [(OF.CONST_B, [0.0, 1.0, 0.0], 0), (OF.GAUSS_B, [0.5, 2.0], 7),
             (OF.RANDOM_NE, [0.01, 0.02, 0.0], 9), (OF.EXP_NE, [0.01, 3.0, 0.1], 0)]
    got = _gpu_fields(eng, BOX, res, descs, seeds, bs)
    want = OF.evaluate_fields(OF.Grid(BOX, res), descs, seeds)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert np.allclose(got[k], want[k], rtol=1e-12, atol=1e-13), k
    assert got['thermal_electron_density'].min() >= 0.0       # min_ne clamp


def test_philox_field_moments_at_scale(eng):
    """Distribution check at 64^3 x 3 x 4 realisations; realisations and components independent."""
    seeds = [11, 12, 13, 14]
    got = _gpu_fields(eng, BOX, [64, 64, 64], [(OF.GAUSS_B, [1.5, 2.0], 0)], seeds, 4)['magnetic_field']
    z = (got - 1.5) / 2.0
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15 / n)
    assert abs((z ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
    flat = z.reshape(4, -1)
    c = np.corrcoef(flat)
    assert np.abs(c - np.eye(4)).max() < 5 / np.sqrt(flat.shape[1])
    comp = np.corrcoef(z.reshape(-1, 3).T)
    assert np.abs(comp - np.eye(3)).max() < 5 / np.sqrt(z.size / 3)
    # the same seed reproduces the same realisation regardless of its position in the ensemble
    again = _gpu_fields(eng, BOX, [64, 64, 64], [(OF.GAUSS_B, [1.5, 2.0], 0)], [13], 4)['magnetic_field']
    assert np.array_equal(again[0], got[2])


@pytest.mark.parametrize('nside', [1, 2, 8, 32, 64])
def test_healpix_dirs_match_oracle(eng, nside):
    got = eng.healpix_ring_dirs(nside).cpu().numpy()
    want = OH.pix2vec_ring(nside)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2.3e-16


def test_test_simulator_sampling(eng):
    import torch
    rng = np.random.default_rng(5)
    box, res = [[0, 2 * np.pi], [0, 0], [-1, 1]], [100, 1, 7]
    og = OF.Grid(box, res)
    descs = [(OF.GAUSS_B, [3.0, 6.0], 0), (OF.COS_NE, [1, 1, 0, 0, np.pi / 2, 0, np.pi / 2], 0)]
    seeds = [3, 4, 5]
    grid = eng.DeviceGrid(box, res)
    f = eng.field_eval(grid, descs, seeds, b_stride=3)
    pts = np.stack([rng.uniform(0, 2 * np.pi, 40), np.zeros(40), rng.uniform(-1, 1, 40)], 1)
    c = og.axis_coords(0)
    pts[0, 0], pts[1, 0], pts[2, 0] = 0.5 * (c[3] + c[4]), c[0], c[-1]     # tie + end points
    out = eng.test_simulate(grid, f['magnetic_field'], f['thermal_electron_density'], 3,
                            torch.from_numpy(pts).cuda(), b_axis=1, scale=1.0).cpu().numpy()
    of = OF.evaluate_fields(og, descs, seeds)
    for r in range(3):
        Bh = f['magnetic_field'][r].cpu().numpy().reshape(*og.shape, 3)
        want = OF.test_simulator_sample(og, Bh, of['thermal_electron_density'][0], pts, 1)
        assert np.array_equal(out[r], want)
    bad = torch.tensor([[7.0, 0.0, 0.0]], dtype=torch.float64).cuda()
    with pytest.raises(ValueError):
        eng.test_simulate(grid, f['magnetic_field'], f['thermal_electron_density'], 3, bad)


def test_template_golden_end_to_end_on_gpu(eng):
    """-55.581767201053125 (tests/mocks_for_templates.py:111-112) through the CUDA path."""
    import torch
    box, res = [[0, 2 * np.pi], [0, 0], [0, 0]], [30, 1, 1]
    grid = eng.DeviceGrid(box, res)
    f = eng.field_eval(grid, [(OF.CONST_B, [10.5, 0.5, 3.0], 0), (OF.CONST_NE, [1.0], 0)], [1, 2],
                       b_stride=3)
    pts = torch.tensor([[float(i), 0.0, 0.0] for i in range(5)], dtype=torch.float64).cuda()
    sims = eng.test_simulate(grid, f['magnetic_field'], f['thermal_electron_density'], 2, pts, b_axis=1)
    meas = torch.ones(5, dtype=torch.float64).cuda()
    var = torch.full((5,), 0.1 ** 2, dtype=torch.float64).cuda()
    assert eng.loglike_simple(sims, meas, var) == pytest.approx(-55.581767201053125, abs=1e-11)


@pytest.mark.parametrize('bs', [3, 4])
def test_lsa_field_matches_oracle(eng, bs):
    """WMAP LSA regular field (gridded counterpart of hamx/breg_lsa.py), summed with a constant field."""
    p = [6.0, np.radians(27.0), np.radians(0.9), np.radians(25.0), 8.0, 1.0, 3.0, 20.0]
    descs = [(OF.LSA_B, p, 0), (OF.CONST_B, [0.1, 0.0, -0.2], 0)]
    for res in ([33, 31, 9], [16, 16, 8]):
        got = _gpu_fields(eng, BOX, res, descs, [1], bs)['magnetic_field']
        want = OF.evaluate_fields(OF.Grid(BOX, res), descs, [1])['magnetic_field']
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    og = OF.Grid(BOX, [33, 31, 9])
    B = OF.lsa_B(og, *p)
    x, y, z = (v * np.ones(og.shape) for v in og.xyz())
    R = np.sqrt(x ** 2 + y ** 2)
    inside = (R > 3.0) & (R < 20.0)
    assert np.allclose(np.linalg.norm(B, axis=-1)[inside], 6.0, rtol=1e-12)      # |B| = b0 in the annulus
    assert np.all(B[~inside] == 0.0)
    br = (B[..., 0] * x + B[..., 1] * y)[inside] / R[inside]                      # pitch angle psi(R) at z = 0
    mid = np.abs(z[inside]) < 1e-9
    psi = np.radians(27.0) + np.radians(0.9) * np.log(R[inside][mid] / 8.0)
    assert np.allclose(br[mid], 6.0 * np.sin(psi), rtol=1e-10)


def test_jf12_regular_field_matches_oracle(eng):
    """Jansson & Farrar 2012 (hamx/breg_jf12.py parameter set): CUDA vs oracle/fields.py:jf12_B, 1e-9."""
    p = [0.1, 3.0, -0.9, -0.8, -2.0, -4.2, 0.0, 0.1, 0.40, 0.27, 1.4, -1.1, 9.22, 16.7, 0.20, 5.3, 4.6, 49, 4.8, 2.9]
    for res in ([33, 31, 9], [41, 41, 17]):
        grid = eng.DeviceGrid(BOX, res)
        og = OF.Grid(BOX, res)
        got = eng.field_eval(grid, [(OF.JF12_B, p, 0)], [0], b_stride=3)['magnetic_field']
        got = got.cpu().numpy().reshape(*og.shape, 3)
        want = OF.jf12_B(og, *p)
        assert np.abs(want).max() > 1.0
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    # summed with a Gaussian component in the wide layout (generic path of the interleaved kernel)
    grid, og = eng.DeviceGrid(BOX, [20, 18, 12]), OF.Grid(BOX, [20, 18, 12])
    descs = [(OF.JF12_B, p, 0), (OF.GAUSS_B, [0.0, 1.5], 1)]
    f = eng.field_eval(grid, descs, [11, 12, 13], b_stride=24)['magnetic_field']
    from tests.helpers import b_to_reference_layout
    got = b_to_reference_layout(f, 3, og.shape)
    want = OF.evaluate_fields(og, descs, [11, 12, 13])['magnetic_field']
    assert np.allclose(got, want, rtol=1e-9, atol=1e-11)
