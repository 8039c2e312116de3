"""GPU parity, stage 2: the LOS ray-march kernel vs oracle/los.py on identical fields."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fields as OF          # noqa: E402
from oracle import healpix as OH         # noqa: E402
from oracle import los as OL             # noqa: E402
from tests.helpers import b_to_reference_layout   # noqa: E402

BOX = [[-20, 20], [-20, 20], [-4, 4]]
KEYS = ('dm', 'fd', 'I', 'Q', 'U')


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    from imagine_b200 import engine, _lib
    _lib.init(0)
    return engine


def _run(eng, res, descs, seeds, nside, bs, cfg_kw, want=KEYS, dirs=None):
    import torch
    grid = eng.DeviceGrid(BOX, res)
    N = len(seeds)
    f = eng.field_eval(grid, descs, seeds, b_stride=bs)
    d_dirs = eng.healpix_ring_dirs(nside) if dirs is None else torch.from_numpy(dirs).cuda()
    cfg = eng.LosSettings(**cfg_kw)
    maps = eng.los_integrate(grid, cfg, d_dirs, f['magnetic_field'],
                             f.get('thermal_electron_density'), f.get('cosmic_ray_electron_density'),
                             n_real=N, want=want)
    torch.cuda.synchronize()
    og = OF.Grid(BOX, res)
    # identical realisations for the oracle: the GPU-generated grids themselves (north_star)
    B = b_to_reference_layout(f['magnetic_field'], N, og.shape)
    ne = f.get('thermal_electron_density')
    cre = f.get('cosmic_ray_electron_density')
    ne = None if ne is None else ne.cpu().numpy().reshape(-1, *og.shape)
    cre = None if cre is None else cre.cpu().numpy().reshape(-1, *og.shape)
    omaps = OL.integrate(og, OL.LosConfig(**cfg_kw), d_dirs.cpu().numpy(), B, ne, cre)
    return {k: v.cpu().numpy() for k, v in maps.items()}, omaps


def _check(maps, omaps, keys=KEYS, rel=1e-9):
    for k in keys:
        got, want = maps[k], omaps[k]
        assert got.shape == want.shape, k
        scale = np.abs(want).max()
        assert scale > 0, k
        # norm-wise 1e-9 plus element-wise 1e-9 with an absolute floor far below the signal
        assert np.abs(got - want).max() <= rel * scale, (k, np.abs(got - want).max() / scale)
        assert np.allclose(got, want, rtol=rel, atol=1e-12 * scale), k


DESCS = [(OF.CONST_B, [0.0, 1.0, 0.0], 0), (OF.GAUSS_B, [0.0, 2.0], 1),
         (OF.EXP_NE, [0.01, 3.0, 0.1], 2), (OF.CRE_ANALYTIC, [0.0217, 5.0, 1.0, 8.3], 3)]


@pytest.mark.parametrize('bs', [3, 4, 24, 48, 96])
@pytest.mark.parametrize('N', [1, 2, 3, 5, 8, 17, 33])
def test_all_observables_match_oracle(eng, bs, N):
    maps, omaps = _run(eng, [20, 18, 12], DESCS, list(range(50, 50 + N)), 4, bs, dict(dr=0.2))
    _check(maps, omaps)


@pytest.mark.parametrize('bs', [4, 24, 48, 96])
def test_nside16_grid32_23ghz(eng, bs):
    maps, omaps = _run(eng, [32, 32, 16], DESCS, [7, 8, 9, 10, 11, 12], 16, bs, dict(dr=0.1))
    _check(maps, omaps)


@pytest.mark.parametrize('il', [24, 48, 96])
def test_interleaved_layout_variants(eng, il):
    """Interleaved layouts through every kernel variant: Faraday only, sync only, variable index, random n_e."""
    maps, omaps = _run(eng, [20, 18, 12], DESCS[:3], [1, 2, 3, 4, 5], 4, il, dict(dr=0.2), want=('dm', 'fd'))
    _check(maps, omaps, ('dm', 'fd'))
    d = [DESCS[0], DESCS[1], DESCS[3]]
    maps, omaps = _run(eng, [20, 18, 12], d, [1, 2, 3, 4, 5, 6, 7], 4, il, dict(dr=0.2), want=('I', 'Q', 'U'))
    _check(maps, omaps, ('I', 'Q', 'U'))
    maps, omaps = _run(eng, [20, 18, 12], DESCS, list(range(1, 10)), 4, il,
                       dict(dr=0.2, alpha=2.7, beta=0.02, theta=-0.05, nu_ghz=2.3))
    _check(maps, omaps)
    descs = DESCS[:2] + [(OF.RANDOM_NE, [0.02, 0.01, 0.0], 4), DESCS[3]]
    maps, omaps = _run(eng, [16, 16, 8], descs, [1, 2, 3, 4, 5, 6], 4, il, dict(dr=0.2))
    _check(maps, omaps)
    # the layouts agree with each other bit for bit on the fields and to rounding on the maps
    a, _ = _run(eng, [20, 18, 12], DESCS, [1, 2, 3, 4, 5], 4, 4, dict(dr=0.2))
    b, _ = _run(eng, [20, 18, 12], DESCS, [1, 2, 3, 4, 5], 4, il, dict(dr=0.2))
    for k in KEYS:
        assert np.allclose(a[k], b[k], rtol=1e-11, atol=1e-13 * np.abs(a[k]).max()), k


def test_low_frequency_strong_faraday_rotation(eng):
    descs = [(OF.CONST_B, [1.5, 2.0, 0.3], 0), (OF.GAUSS_B, [0.0, 1.0], 1),
             (OF.EXP_NE, [0.05, 5.0, 1.0], 2), (OF.CRE_CONST, [0.0217], 3)]
    maps, omaps = _run(eng, [24, 24, 12], descs, [1, 2, 3, 4], 4, 4, dict(dr=0.1, nu_ghz=1.4))
    _check(maps, omaps)


def test_variable_spectral_index(eng):
    maps, omaps = _run(eng, [20, 18, 12], DESCS, [1, 2, 3, 4], 4, 4,
                       dict(dr=0.2, alpha=3.0, beta=0.02, theta=-0.05))
    _check(maps, omaps, rel=1e-9)


def test_non_integer_spectral_index(eng):
    maps, omaps = _run(eng, [20, 18, 12], DESCS, [1, 2], 4, 3, dict(dr=0.2, alpha=2.6))
    _check(maps, omaps)


def test_faraday_only_and_sync_only(eng):
    maps, omaps = _run(eng, [20, 18, 12], DESCS[:3], [1, 2, 3, 4], 4, 4, dict(dr=0.2), want=('dm', 'fd'))
    _check(maps, omaps, ('dm', 'fd'))
    d = [DESCS[0], DESCS[1], DESCS[3]]
    maps, omaps = _run(eng, [20, 18, 12], d, [1, 2, 3, 4], 4, 4, dict(dr=0.2), want=('I', 'Q', 'U'))
    _check(maps, omaps, ('I', 'Q', 'U'))


def test_deterministic_ensemble_is_replicated(eng):
    d = [DESCS[0], DESCS[2], DESCS[3]]
    maps, omaps = _run(eng, [20, 18, 12], d, [1, 2, 3], 4, 4, dict(dr=0.2))
    for k in KEYS:
        assert np.array_equal(maps[k][0], maps[k][1]) and np.array_equal(maps[k][0], maps[k][2])
        assert np.allclose(maps[k][0], omaps[k][0], rtol=1e-9, atol=1e-12 * np.abs(omaps[k]).max())


def test_random_ne_per_realisation(eng):
    descs = [(OF.CONST_B, [1.0, 1.0, 0.5], 0), (OF.RANDOM_NE, [0.02, 0.01, 0.0], 4),
             (OF.CRE_CONST, [0.0217], 3)]
    maps, omaps = _run(eng, [16, 16, 8], descs, [1, 2, 3, 4, 5], 4, 3, dict(dr=0.2))
    _check(maps, omaps)


def test_arbitrary_directions_and_observer_outside(eng):
    rng = np.random.default_rng(1)
    v = rng.normal(size=(37, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[0] = [1.0, 0.0, 0.0]
    v[1] = [0.0, 0.0, 1.0]          # pole: sky basis degenerate -> East = +y by definition
    v[2] = [0.0, -1.0, 0.0]
    maps, omaps = _run(eng, [20, 18, 12], DESCS, [1, 2], 4, 4,
                       dict(dr=0.15, observer=(-25.0, 1.0, 0.5), r_max=60.0), dirs=v)
    _check(maps, omaps)


def test_closed_forms_on_gpu(eng):
    """Uniform medium: DM = ne * L, RM = 0.812 ne B_par L; zero rotation => PI/I = (p+1)/(p+7/3)."""
    descs = [(OF.CONST_B, [0.0, 2.0, 0.0], 0), (OF.CONST_NE, [0.01], 0), (OF.CRE_CONST, [0.0217], 0)]
    maps, _ = _run(eng, [21, 21, 9], descs, [1], 4, 4, dict(dr=0.1))
    og = OF.Grid(BOX, [21, 21, 9])
    dirs = OH.pix2vec_ring(4)
    _, inside, _, _ = OL.sample_geometry(og, OL.LosConfig(dr=0.1), dirs)
    path_pc = inside.sum(1) * 0.1 * 1e3
    assert np.allclose(maps['dm'][0], 0.01 * path_pc, rtol=1e-12)
    assert np.allclose(maps['fd'][0], 0.812 * 0.01 * (-2.0 * dirs[:, 1]) * path_pc, rtol=1e-11, atol=1e-10)
    descs = [(OF.CONST_B, [0.0, 2.0, 0.0], 0), (OF.CRE_CONST, [0.0217], 0)]
    maps, _ = _run(eng, [21, 21, 9], descs, [1], 4, 4, dict(dr=0.1, alpha=2.8), want=('I', 'Q', 'U'))
    ok = maps['I'][0] > 0
    frac = np.sqrt(maps['Q'] ** 2 + maps['U'] ** 2)[0][ok] / maps['I'][0][ok]
    assert np.allclose(frac, 3.8 / (2.8 + 7 / 3), rtol=1e-12)


def test_linearity_in_ne_and_cre(eng):
    """Size-independent property: fd, dm linear in n_e; I linear in the CRE normalisation."""
    base = [(OF.CONST_B, [0.4, 1.0, -0.2], 0), (OF.GAUSS_B, [0.0, 2.0], 1)]
    a, _ = _run(eng, [24, 24, 12], base + [(OF.EXP_NE, [0.01, 3.0, 0.5], 2), (OF.CRE_CONST, [0.01], 3)],
                [5, 6], 8, 4, dict(dr=0.2), want=('dm', 'fd', 'I'))
    b, _ = _run(eng, [24, 24, 12], base + [(OF.EXP_NE, [0.04, 3.0, 0.5], 2), (OF.CRE_CONST, [0.03], 3)],
                [5, 6], 8, 4, dict(dr=0.2), want=('dm', 'fd', 'I'))
    assert np.allclose(b['dm'], 4 * a['dm'], rtol=1e-13)
    assert np.allclose(b['fd'], 4 * a['fd'], rtol=1e-12, atol=1e-12 * np.abs(a['fd']).max())
    assert np.allclose(b['I'], 3 * a['I'], rtol=1e-13)


def test_invalid_arguments_are_rejected(eng):
    import torch
    grid = eng.DeviceGrid(BOX, [8, 8, 8])
    f = eng.field_eval(grid, DESCS[:2], [1], b_stride=4)
    dirs = eng.healpix_ring_dirs(1)
    with pytest.raises(ValueError):      # fd without n_e
        eng.los_integrate(grid, eng.LosSettings(), dirs, f['magnetic_field'], None, None, 1, want=('fd',))
    flat = eng.DeviceGrid([[0, 1], [0, 0], [0, 0]], [8, 1, 1])
    fb = eng.field_eval(flat, DESCS[:1], [1], b_stride=4)
    with pytest.raises(ValueError):      # trilinear needs >= 2 vertices per axis
        eng.los_integrate(flat, eng.LosSettings(), dirs, fb['magnetic_field'],
                          torch.zeros(1, 8, dtype=torch.float64).cuda(), None, 1, want=('fd',))


def test_edge_cases_empty_and_all_outside(eng):
    import torch
    grid = eng.DeviceGrid(BOX, [8, 8, 8])
    f = eng.field_eval(grid, DESCS, [1, 2], b_stride=24)
    none = torch.zeros((0, 3), dtype=torch.float64, device='cuda')
    out = eng.los_integrate(grid, eng.LosSettings(), none, f['magnetic_field'], f['thermal_electron_density'],
                            f['cosmic_ray_electron_density'], n_real=2)
    assert all(tuple(t.shape) == (2, 0) for t in out.values())
    # observer far outside, looking away: every sample is outside the box -> exact zeros
    away = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=torch.float64, device='cuda')
    out = eng.los_integrate(grid, eng.LosSettings(observer=(100.0, 100.0, 50.0)), away, f['magnetic_field'],
                            f['thermal_electron_density'], f['cosmic_ray_electron_density'], n_real=2)
    assert all(bool((t == 0).all()) for t in out.values())
    # zero radial steps
    out = eng.los_integrate(grid, eng.LosSettings(n_steps=0), away, f['magnetic_field'],
                            f['thermal_electron_density'], None, n_real=2, want=('dm', 'fd'))
    assert all(bool((t == 0).all()) for t in out.values())
