"""
GPU: the reference-facing drop-ins (Field / Simulator / Likelihood plugins).  The likelihood
tests are the reference's own imagine/tests/test_likelihood.py with the GPU classes swapped in;
the pipeline test follows imagine/tests/test_templates.py:146-219 (golden -55.581767201053125).
"""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fields as OF           # noqa: E402
from oracle import healpix as OH          # noqa: E402
from oracle import likelihood as OK       # noqa: E402
from oracle import los as OL              # noqa: E402


@pytest.fixture(scope='module')
def api():
    import torch
    assert torch.cuda.is_available()
    import imagine_b200.api as api
    from imagine_b200 import _lib
    _lib.init(0)
    return api


# ------------------------------------------------------------ reference test_likelihood.py, GPU classes
def test_simple_without_cov(api):
    simdict, meadict = api.Simulations(), api.Measurements()
    arr_a, arr_b = np.random.rand(1, 48), np.random.rand(3, 48)
    meadict.append(name=('test', None, 2, None), data=api.Observable(arr_a, 'measured'), otype='HEALPix')
    simdict.append(name=('test', None, 2, None), data=api.Observable(arr_b, 'simulated'), otype='HEALPix')
    result = api.B200SimpleLikelihood(meadict)(simdict)
    diff = np.mean(arr_b, axis=0) - arr_a
    assert np.allclose(result, -0.5 * float(np.vdot(diff, diff)))


def test_simple_with_cov(api):
    simdict, meadict, covdict = api.Simulations(), api.Measurements(), api.Covariances()
    arr_a, arr_b, arr_c = np.random.rand(1, 4), np.random.rand(5, 4), np.random.rand(4, 4)
    key = ('test', None, 4, None)
    meadict.append(name=key, data=api.Observable(arr_a, 'measured'), otype='plain')
    simdict.append(name=key, data=api.Observable(arr_b, 'simulated'), otype='plain')
    covdict.append(name=key, cov_data=api.Observable(arr_c, 'covariance'))
    result = api.B200SimpleLikelihood(meadict, covdict)(simdict)
    diff = np.mean(arr_b, axis=0) - arr_a
    sign, logdet = np.linalg.slogdet(arr_c * 2. * np.pi)
    baseline = -0.5 * (np.vdot(diff, np.linalg.solve(arr_c, diff.T)) + sign * logdet)
    assert np.allclose(result, baseline)


@pytest.mark.parametrize('trace', [False, True])
def test_ensemble_equals_simple_for_identical_realisations(api, trace):
    for _ in range(5):
        simdict, meadict, covdict = api.Simulations(), api.Measurements(), api.Covariances()
        key = ('test', None, 4, None)
        meadict.append(name=key, data=api.Observable(np.random.rand(1, 4), 'measured'), otype='plain')
        arr_c = np.diag(np.random.rand(4)) if trace else np.random.rand(4, 4)
        covdict.append(name=key, cov_data=api.Observable(arr_c, 'covariance'))
        arr_ens = np.repeat(np.random.rand(1, 4), 2, axis=0)
        simdict.append(name=key, data=api.Observable(arr_ens, 'simulated'), otype='plain')
        result_simple = api.B200SimpleLikelihood(meadict, covdict)(simdict)
        result_ensemble = api.B200EnsembleLikelihood(meadict, covdict, use_trace_approximation=trace)(simdict)
        assert result_ensemble == result_simple


def test_diag(api):
    simdict, meadict, covdict = api.Simulations(), api.Measurements(), api.Covariances()
    key = ('test', None, 40, None)
    meadict.append(name=key, data=api.Observable(np.random.rand(1, 40), 'measured'), otype='plain')
    covdict.append(name=key, cov_data=api.Observable(np.diag(np.random.rand(40)), 'covariance'))
    simdict.append(name=key, data=api.Observable(np.random.rand(10, 40), 'simulated'), otype='plain')

    def diagonal_mcov(data):                  # the reference passes the function object
        raise AssertionError('mapped by name, never called on the host')
    result_ens = api.B200EnsembleLikelihood(meadict, covdict, cov_func=diagonal_mcov)(simdict)
    result_diag = api.B200EnsembleLikelihoodDiagonal(meadict, covdict)(simdict)
    assert np.allclose(result_diag, result_ens)
    with pytest.raises(NotImplementedError):
        api.B200EnsembleLikelihood(meadict, covdict, cov_func=lambda d: (d.mean(0), np.eye(40)))


def test_goldens_through_the_plugin_api(api, golden_likelihood):
    cls = {'ensemble': api.B200EnsembleLikelihood, 'ensemble_diagonal': api.B200EnsembleLikelihoodDiagonal,
           'simple': api.B200SimpleLikelihood}
    for name, c in golden_likelihood.items():
        m = c['sims'].shape[1]
        key = ('test', None, m, None)
        M, C, S = api.Measurements(), api.Covariances(), api.Simulations()
        M.append(name=key, data=c['meas'].reshape(1, -1), otype='plain')
        S.append(name=key, data=c['sims'], otype='plain')
        if c['cov'] is not None:
            C.append(name=key, cov_data=c['cov'], otype='plain')
        masks = None
        if c['mask'] is not None:
            masks = api.Masks()
            masks.append(name=key, data=c['mask'].reshape(1, -1), otype='plain')
        kw = {}
        if c['kind'] == 'ensemble':
            kw = dict(cov_func=str(c['estimator']), use_trace_approximation=bool(c['use_trace_approximation']))
        lk = cls[c['kind']](M, C if c['cov'] is not None else None, mask_dict=masks, **kw)
        assert abs(lk(S) - c['value']) <= 1e-8, name


def test_bootstrap_dispersion_and_pickle(api):
    rng = np.random.default_rng(1)
    key = ('fd', None, 2, None)
    M, S = api.Measurements(), api.Simulations()
    M.append(name=key, data=rng.normal(size=(1, 48)), cov_data=np.full(48, 0.5), otype='HEALPix')
    S.append(name=key, data=rng.normal(size=(12, 48)), otype='HEALPix')
    lk = api.B200EnsembleLikelihood(M, compute_dispersion=True, n_bootstrap=20)
    np.random.seed(3)
    value, disp = lk(S)
    assert np.isfinite(value) and disp > 0
    assert abs(value - OK.ensemble_loglike(S[key].data, M[key].data, np.full(48, 0.5))) <= 1e-8
    lk2 = pickle.loads(pickle.dumps(lk))
    lk2.compute_dispersion = False
    assert lk2(S) == value


# ------------------------------------------------------------ template pipeline (TestSimulator + SimpleLikelihood)
def test_template_pipeline_golden(api):
    u = api.units

    class FakeRandomTE(api.ThermalElectronDensityField):     # foreign plugin: no device descriptor
        NAME = 'fake_rnd_TE'
        STOCHASTIC_FIELD = True
        DEPENDENCIES_LIST = []
        PARAMETER_NAMES = ['param']

        def compute_field(self, seed):
            return np.ones(self.data_shape) * u.cm ** (-3)

    fd_units = u.microgauss * u.cm ** -3
    x = np.arange(5.0)
    coords = {'x': x * u.kpc, 'y': np.zeros(5) * u.kpc, 'z': np.zeros(5) * u.kpc}
    key = ('test', None, 'tab', None)
    meas, covs = api.Measurements(), api.Covariances()
    meas.append(name=key, data=np.ones((1, 5)) * fd_units, otype='tabular', coords=coords)
    covs.append(name=key, cov_data=(np.ones(5) * 0.1) ** 2 * fd_units ** 2, otype='tabular')
    likelihood = api.B200SimpleLikelihood(meas, covs)
    grid = api.UniformGrid(box=[[0, 2 * np.pi] * u.kpc, [0, 0] * u.kpc, [0, 0] * u.kpc], resolution=[30, 1, 1])
    simulator = api.B200TestSimulator(meas)

    def loglike(cube):                                        # Pipeline._core_likelihood, pipeline.py:905-943
        te = FakeRandomTE(grid, parameters={'param': cube[0]}, ensemble_size=2, ensemble_seeds=[11, 12])
        b = api.ConstantMagneticField(grid, parameters={'Bx': cube[1] * u.microgauss, 'By': cube[2] * u.microgauss,
                                                        'Bz': 3. * u.microgauss}, ensemble_size=2)
        return likelihood(simulator([te, b]))

    assert np.allclose(loglike(np.array([0.5, 10.5, 0.5])), -55.581767201053125, rtol=0, atol=1e-11)
    # per-realisation contract through the base-class loop gives the same rows
    te = FakeRandomTE(grid, parameters={'param': 1.0}, ensemble_size=2, ensemble_seeds=[11, 12])
    b = api.ConstantMagneticField(grid, parameters={'Bx': 10.5 * u.microgauss, 'By': 0.5 * u.microgauss,
                                                    'Bz': 3. * u.microgauss}, ensemble_size=2)
    batched = simulator([te, b])[key].data
    looped = api.Simulator.__call__(simulator, [te, b])[key].data
    assert np.array_equal(batched, looped)
    with pytest.raises(AssertionError):
        simulator([b])                                        # missing required field


# ------------------------------------------------------------ LOS simulator vs the oracle, through the API
def _los_setup(api, nside=4, N=4, res=(20, 18, 12)):
    u = api.units
    m = 12 * nside * nside
    grid = api.UniformGrid(box=[[-20, 20] * u.kpc, [-20, 20] * u.kpc, [-4, 4] * u.kpc], resolution=list(res))
    rng = np.random.default_rng(0)
    meas = api.Measurements()
    keys = [('fd', None, nside, None), ('dm', None, nside, None), ('sync', 23.0, nside, 'I'),
            ('sync', 23.0, nside, 'Q'), ('sync', 23.0, nside, 'U'), ('sync', 1.4, nside, 'PI'),
            ('sync', 1.4, nside, 'PA')]
    units = [u.rad / u.m ** 2, u.pc / u.cm ** 3, u.K, u.K, u.K, u.K, u.rad]
    for k, un in zip(keys, units):
        meas.append(name=k, data=rng.normal(size=(1, m)) * un, cov_data=np.ones(m) * un ** 2, otype='HEALPix')
    seeds_b = list(range(101, 101 + N))
    fields = [
        api.ConstantMagneticField(grid, parameters={'Bx': 0 * u.microgauss, 'By': 1 * u.microgauss,
                                                    'Bz': 0 * u.microgauss}, ensemble_size=N),
        api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 2 * u.microgauss},
                                       ensemble_size=N, ensemble_seeds=seeds_b),
        api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3,
                                                          'scale_radius': 3 * u.kpc, 'scale_height': 0.5 * u.kpc},
                                        ensemble_size=N),
        api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5 * u.kpc, 'z0': 1 * u.kpc,
                                                         'r_sun': 8.3 * u.kpc}, ensemble_size=N)]
    return grid, meas, keys, fields, seeds_b


def test_b200_simulator_matches_oracle(api):
    nside, N, res = 4, 4, (20, 18, 12)
    grid, meas, keys, fields, seeds_b = _los_setup(api, nside, N, res)
    sim = api.B200Simulator(meas, radial_step=0.2, temperature='rj')     # the oracle's (and the kernels') scale
    sims = sim(fields)
    assert isinstance(sims, api.Simulations) and set(sims.keys()) == set(keys)
    # default: thermodynamic temperature like hammurabiX -- the same maps times the K_RJ -> K_CMB factor
    from imagine_b200.constants import rj_to_cmb
    cmb = api.B200Simulator(meas, radial_step=0.2)(fields)
    assert np.array_equal(cmb[keys[0]].data, sims[keys[0]].data)                        # fd: untouched
    assert np.allclose(cmb[keys[2]].data, sims[keys[2]].data * rj_to_cmb(23.0), rtol=1e-14, atol=0)
    assert np.allclose(cmb[keys[5]].data, sims[keys[5]].data * rj_to_cmb(1.4), rtol=1e-14, atol=0)
    assert np.array_equal(cmb[keys[6]].data, sims[keys[6]].data)                        # PA: an angle
    og = OF.Grid([[-20, 20], [-20, 20], [-4, 4]], res)
    stream = fields[1].device_descriptor()[2]
    descs = [(OF.CONST_B, [0, 1, 0], 0), (OF.GAUSS_B, [0, 2], stream), (OF.EXP_NE, [0.01, 3, 0.5], 0),
             (OF.CRE_ANALYTIC, [0.0217, 5, 1, 8.3], 0)]
    of = OF.evaluate_fields(og, descs, seeds_b)
    dirs = OH.pix2vec_ring(nside)
    o23 = OL.integrate(og, OL.LosConfig(dr=0.2, nu_ghz=23.0), dirs, of['magnetic_field'],
                       of['thermal_electron_density'], of['cosmic_ray_electron_density'])
    o14 = OL.integrate(og, OL.LosConfig(dr=0.2, nu_ghz=1.4), dirs, of['magnetic_field'],
                       of['thermal_electron_density'], of['cosmic_ray_electron_density'])
    want = {keys[0]: o23['fd'], keys[1]: o23['dm'], keys[2]: o23['I'], keys[3]: o23['Q'], keys[4]: o23['U'],
            keys[5]: np.sqrt(o14['Q'] ** 2 + o14['U'] ** 2), keys[6]: 0.5 * np.arctan2(o14['U'], o14['Q'])}
    for k in keys:
        got = sims[k].data
        assert sims[k].shape == (N, 12 * nside ** 2) and sims[k].otype == 'HEALPix'
        scale = np.abs(want[k]).max()
        tol = 1e-9 if k[3] != 'PA' else 1e-7          # the angle amplifies Q,U rounding near |P| ~ 0
        assert np.abs(got - want[k]).max() <= tol * scale, k
    # per-realisation contract: base-class loop over host fields reproduces the batched rows
    looped = api.Simulator.__call__(sim, fields)
    for k in keys[:5]:
        assert np.allclose(looped[k].data, sims[k].data, rtol=1e-9, atol=1e-12 * np.abs(sims[k].data).max()), k


def test_full_drop_in_evaluation_and_units(api):
    u = api.units
    nside, N = 4, 4
    grid, meas, keys, fields, _ = _los_setup(api, nside, N)
    sim = api.B200Simulator(meas, radial_step=0.2)
    lk = api.B200EnsembleLikelihood(meas)
    sims = sim(fields)
    value = lk(sims)
    want = sum(OK.ensemble_loglike(sims[k].data, meas[k].data, np.ones(12 * nside ** 2)) for k in keys)
    assert abs(value - want) <= 1e-8 * max(1.0, abs(want))
    # measurements in other units: the simulator converts its maps (simulator.py:417-419)
    meas2 = api.Measurements()
    k = ('dm', None, nside, None)
    meas2.append(name=k, data=np.ones((1, 192)) * (u.kpc / u.cm ** 3), otype='HEALPix')
    s2 = api.B200Simulator(meas2, radial_step=0.2)(fields)
    assert np.allclose(s2[k].data * 1e3, sims[k].data, rtol=1e-14)
    # masked likelihood: masked keys are renamed (name, freq, n_kept, tag)
    mask = api.Masks()
    keep = (np.arange(192) % 3 != 0).astype(float).reshape(1, -1)
    for key in keys:
        mask.append(name=key, data=keep, otype='HEALPix')
    lkm = api.B200EnsembleLikelihood(meas, mask_dict=mask)
    b = keep[0].astype(bool)
    want = sum(OK.ensemble_loglike(sims[key].data[:, b], meas[key].data[:, b], np.ones(int(b.sum()))) for key in keys)
    assert abs(lkm(sims) - want) <= 1e-8 * max(1.0, abs(want))
    # checkpointing: drop-ins survive pickling without device handles (io.py:25-75)
    sim2, lk2 = pickle.loads(pickle.dumps(sim)), pickle.loads(pickle.dumps(lk))
    assert lk2(sim2(fields)) == value
    host_sims = pickle.loads(pickle.dumps(sims[keys[0]]))
    assert np.array_equal(host_sims.data, sims[keys[0]].data)
    # asynchronous re-upload from pinned host memory (overlaps the simulator) changes nothing,
    # before or after the first evaluation, and the likelihood still pickles afterwards
    lk3 = api.B200EnsembleLikelihood(meas)
    for _ in range(3):
        lk3.upload_measurements()
        assert lk3(sim(fields)) == value
    assert pickle.loads(pickle.dumps(lk3))(sims) == value
    lkm.upload_measurements()
    assert abs(lkm(sims) - want) <= 1e-8 * max(1.0, abs(want))


def test_gpu_fields_serve_host_consumers(api):
    u = api.units
    grid = api.UniformGrid(box=[[-15, 15] * u.kpc] * 3, resolution=[15, 15, 15])
    f = api.ExponentialThermalElectrons(grid, parameters={'central_density': 1.0 * u.cm ** -3,
                                                          'scale_radius': 3.3 * u.kpc, 'scale_height': 3.3 * u.kpc})
    d = f.get_data()
    want = [0.00189093, 0.00361981, 0.00289678, 0.0055453, 0.00289678, 0.0055453, 0.00473694, 0.00906791]
    assert np.allclose(d.value[3:5, 3:5, 3:5].ravel(), want, rtol=0, atol=5e-9)     # tutorial_fields.ipynb:249-256
    g = api.NaiveGaussianMagneticField(grid, parameters={'a0': 3 * u.microgauss, 'b0': 6 * u.microgauss},
                                       ensemble_size=2, ensemble_seeds=[5, 6])
    a, b = g.get_data(0), g.get_data(1)
    assert a.shape == (15, 15, 15, 3) and not np.array_equal(a.value, b.value)
    assert np.array_equal(g.get_data(0).value, a.value)
    assert abs(a.value.mean() - 3) < 0.5 and abs(a.value.std() - 6) < 0.5


def test_four_parameter_lsa_fit_prefers_the_truth(api):
    """configs[4] in miniature: the drop-ins behind a sampler-style loglike(theta) callback."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'gmf_fit_lsa.py')
    spec = importlib.util.spec_from_file_location('gmf_fit_lsa', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    loglike = mod.build(grid_n=32, nside=8, n_ens=8, noise=0.05)
    truth = loglike([6.0, 27.0, 0.9, 25.0])
    assert np.isfinite(truth)
    assert loglike([6.0, 27.0, 0.9, 25.0]) == truth                  # deterministic given the seeds
    for theta in ([4.0, 27.0, 0.9, 25.0], [6.0, 10.0, 0.9, 25.0], [6.0, 27.0, 0.9, -10.0]):
        assert loglike(theta) < truth
    assert loglike([11.0, 27.0, 0.9, 25.0]) == -np.inf               # outside the prior box
    chain, acc = mod.metropolis(loglike, [5.5, 25.0, 1.0, 22.0], 30, np.array([0.05, 0.3, 0.05, 0.5]),
                                np.random.default_rng(0))
    assert chain.shape == (30, 4) and 0.0 <= acc <= 1.0


def test_config1_basic_pipeline_shapes(api):
    """
    BASELINE.json configs[0] = examples/basic_pipeline.py:42-162: grid [100,1,1] on x in [0, 2 pi],
    10 tabular points, ensemble 48, CosThermalElectronDensity (a = 1, beta = gamma = pi/2) x
    NaiveGaussianMagneticField, TestSimulator + EnsembleLikelihood -- through the GPU drop-ins,
    checked against the oracle fed with the GPU-generated realisations.
    """
    u = api.units
    x = np.linspace(0.01, 2. * np.pi - 0.01, 10)
    np.random.seed(233)
    signal = (1 + np.cos(x)) * np.random.normal(loc=3., scale=6., size=10)
    fd = signal + np.random.normal(loc=0., scale=0.1, size=10)              # tutorial_one golden data
    key = ('test', None, 'tab', None)
    fd_units = u.microgauss * u.cm ** -3
    coords = {'x': x * u.kpc, 'y': np.zeros(10) * u.kpc, 'z': np.zeros(10) * u.kpc}
    meas = api.Measurements()
    meas.append(name=key, data=fd.reshape(1, 10) * fd_units, cov_data=np.full(10, 0.1 ** 2) * fd_units ** 2,
                otype='tabular', coords=coords)
    grid = api.UniformGrid(box=[[0, 2 * np.pi] * u.kpc, [0, 0] * u.kpc, [0, 0] * u.kpc], resolution=[100, 1, 1])
    N = 48
    np.random.seed(1)
    seeds = [int(s) for s in np.random.randint(1, 1431655765, N)]
    sim, lik = api.B200TestSimulator(meas), api.B200EnsembleLikelihood(meas)

    def fields(a0, b0):
        return [api.CosThermalElectronDensity(grid, parameters={
                    'n0': 1 * u.cm ** -3, 'a': 1 * u.rad / u.kpc, 'alpha': 0 * u.rad, 'b': 0 * u.rad / u.kpc,
                    'beta': np.pi / 2 * u.rad, 'c': 0 * u.rad / u.kpc, 'gamma': np.pi / 2 * u.rad}, ensemble_size=N),
                api.NaiveGaussianMagneticField(grid, parameters={'a0': a0 * u.microgauss, 'b0': b0 * u.microgauss},
                                               ensemble_size=N, ensemble_seeds=seeds)]

    fl = fields(3.0, 6.0)
    sims = sim(fl)
    assert sims[key].shape == (N, 10)
    # oracle on the same realisations: fields through the host contract (get_data), sampling + likelihood in NumPy
    og = OF.Grid([[0, 2 * np.pi], [0, 0], [0, 0]], [100, 1, 1])
    ne = OF.cos_ne(og, 1, 1, 0, 0, np.pi / 2, 0, np.pi / 2)
    pts = np.stack([x, np.zeros(10), np.zeros(10)], 1)
    rows = np.stack([OF.test_simulator_sample(og, fl[1].get_data(i).value, ne, pts, 1) for i in range(N)])
    assert np.allclose(sims[key].data, rows, rtol=1e-12, atol=0)
    want = OK.ensemble_loglike(rows, fd, np.full(10, 0.01))
    got = lik(sims)
    assert abs(got - want) <= 1e-8 * max(1.0, abs(want))
    assert lik(sim(fields(3.0, 6.0))) == got                      # same seeds, same answer
    assert lik(sim(fields(-4.0, 2.5))) != got


# ------------------------------------------------------------ cross-checks against the reference's OWN classes
def test_reference_likelihoods_agree_on_gpu_simulated_maps(api):
    """GPU simulator -> DeviceObservable -> the reference's EnsembleLikelihood / Diagonal / Simple classes
    (unmodified, imagine.likelihoods) and the GPU drop-ins give the same numbers (1e-8 absolute)."""
    import imagine.likelihoods as RL
    nside, N = 4, 6
    grid, meas, keys, fields, _ = _los_setup(api, nside, N)
    sims = api.B200Simulator(meas, radial_step=0.2)(fields)
    for ref_cls, gpu_cls, kw in ((RL.EnsembleLikelihood, api.B200EnsembleLikelihood, {}),
                                 (RL.EnsembleLikelihoodDiagonal, api.B200EnsembleLikelihoodDiagonal, {}),
                                 (RL.SimpleLikelihood, api.B200SimpleLikelihood, {})):
        want, got = ref_cls(meas, **kw)(sims), gpu_cls(meas, **kw)(sims)
        assert abs(got - want) <= 1e-8, (ref_cls.__name__, got, want)


def test_bootstrap_dispersion_matches_the_reference_seed_for_seed(api):
    """likelihood.py:88-100: np.random.randint resampling; with the same NumPy seed the GPU drop-in draws
    the same row indices and must give the same value and the same dispersion."""
    import imagine.likelihoods as RL
    rng = np.random.default_rng(11)
    key = ('fd', None, 2, None)
    M, S = api.Measurements(), api.Simulations()
    M.append(name=key, data=rng.normal(size=(1, 48)), cov_data=np.full(48, 0.5), otype='HEALPix')
    S.append(name=key, data=rng.normal(size=(12, 48)), otype='HEALPix')
    np.random.seed(7)
    want_v, want_d = RL.EnsembleLikelihood(M, compute_dispersion=True, n_bootstrap=25)(S)
    np.random.seed(7)
    got_v, got_d = api.B200EnsembleLikelihood(M, compute_dispersion=True, n_bootstrap=25)(S)
    assert abs(got_v - want_v) <= 1e-8
    assert abs(got_d - want_d) <= 1e-8 * max(1.0, want_d)


def test_gpu_fields_equal_the_reference_field_classes(api):
    """Deterministic fields: the reference's own compute_field bodies (basic_fields.py, test_field.py) vs
    the CUDA kernels, 1e-9 relative; constants exact."""
    import imagine.fields as RF
    u = api.units
    box = [[-15, 15] * u.kpc, [-15, 15] * u.kpc, [-3, 3] * u.kpc]
    grid = api.UniformGrid(box=box, resolution=[15, 13, 9])
    pairs = [
        (api.ConstantMagneticField, RF.ConstantMagneticField,
         {'Bx': 1.5 * u.microgauss, 'By': -2 * u.microgauss, 'Bz': 0.25 * u.microgauss}),
        (api.ExponentialThermalElectrons, RF.ExponentialThermalElectrons,
         {'central_density': 0.03 * u.cm ** -3, 'scale_radius': 3.3 * u.kpc, 'scale_height': 0.7 * u.kpc}),
        (api.ConstantThermalElectrons, RF.ConstantThermalElectrons, {'ne': 0.02 * u.cm ** -3}),
        (api.CosThermalElectronDensity, RF.CosThermalElectronDensity,
         {'n0': 0.5 * u.cm ** -3, 'a': 1.0 * u.rad / u.kpc, 'alpha': 0.1 * u.rad, 'b': 0.5 * u.rad / u.kpc,
          'beta': 0.2 * u.rad, 'c': 0.25 * u.rad / u.kpc, 'gamma': 0.3 * u.rad}),
    ]
    for gpu_cls, ref_cls, params in pairs:
        got, want = gpu_cls(grid, parameters=params).get_data(), ref_cls(grid, parameters=params).get_data()
        assert got.unit.is_equivalent(want.unit)
        g, w = np.asarray(got.to_value(want.unit)), np.asarray(want.value)
        assert g.shape == w.shape
        if 'Constant' in gpu_cls.__name__:
            assert np.array_equal(g, w)
        else:
            assert np.allclose(g, w, rtol=1e-9, atol=0), gpu_cls.__name__
    # stochastic fields match the reference's MT19937 stream in distribution: moments of one realisation
    ref = RF.NaiveGaussianMagneticField(grid, parameters={'a0': 0.5 * u.microgauss, 'b0': 2 * u.microgauss},
                                        ensemble_size=1, ensemble_seeds=[3]).get_data(0).value
    gpu = api.NaiveGaussianMagneticField(grid, parameters={'a0': 0.5 * u.microgauss, 'b0': 2 * u.microgauss},
                                         ensemble_size=1, ensemble_seeds=[3]).get_data(0).value
    n = ref.size
    assert abs(gpu.mean() - 0.5) < 5 * 2 / np.sqrt(n) and abs(ref.mean() - 0.5) < 5 * 2 / np.sqrt(n)
    assert abs(gpu.std() - 2.0) < 5 * 2 / np.sqrt(2 * n) and abs(ref.std() - 2.0) < 5 * 2 / np.sqrt(2 * n)


def test_reference_test_simulator_equals_the_gpu_test_simulator(api):
    """imagine.simulators.TestSimulator (scipy RegularGridInterpolator, nearest) vs B200TestSimulator on the
    same reference Field objects (foreign host plugins for the GPU simulator)."""
    import imagine.fields as RF
    import imagine.simulators as RS
    u = api.units
    grid = api.UniformGrid(box=[[0, 2 * np.pi] * u.kpc, [0, 1] * u.kpc, [0, 1] * u.kpc], resolution=[30, 3, 3])
    x = np.linspace(0.1, 6.0, 7)
    coords = {'x': x * u.kpc, 'y': np.full(7, 0.4) * u.kpc, 'z': np.full(7, 0.6) * u.kpc}
    M = api.Measurements()
    key = ('test', None, 'tab', None)
    from imagine.observables import TabularDataset
    ds = TabularDataset({'v': np.ones(7), 'e': np.full(7, 0.1), 'x': x, 'y': np.full(7, 0.4), 'z': np.full(7, 0.6)},
                        name='test', data_col='v', err_col='e', x_col='x', y_col='y', z_col='z',
                        units=u.microgauss * u.cm ** -3)
    M.append(dataset=ds)
    key = list(M.keys())[0]
    fields = [RF.ConstantMagneticField(grid, parameters={'Bx': 10.5 * u.microgauss, 'By': 0.5 * u.microgauss,
                                                         'Bz': 3 * u.microgauss}, ensemble_size=2),
              RF.CosThermalElectronDensity(grid, parameters={'n0': 0.5 * u.cm ** -3, 'a': 1.0 * u.rad / u.kpc,
                                                             'alpha': 0 * u.rad, 'b': 0 * u.rad / u.kpc,
                                                             'beta': np.pi / 2 * u.rad, 'c': 0 * u.rad / u.kpc,
                                                             'gamma': np.pi / 2 * u.rad}, ensemble_size=2)]
    want = RS.TestSimulator(M)(fields)[key].data
    got = api.B200TestSimulator(M)(fields)[key].data
    assert got.shape == want.shape == (2, 7)
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_masks_inside_the_simulator(api):
    """hammurabi.py:281-325: one identical mask for every observable, applied while simulating -- masked
    pixels are not marched (zeros), kept pixels are bit-identical to the unmasked run."""
    nside, N = 4, 4
    grid, meas, keys, fields, _ = _los_setup(api, nside, N)
    m = 12 * nside ** 2
    keep = (np.arange(m) % 5 != 0).astype(float).reshape(1, -1)
    masks = api.Masks()
    for k in keys:
        masks.append(name=k, data=keep, otype='HEALPix')
    full = api.B200Simulator(meas, radial_step=0.2)(fields)
    masked = api.B200Simulator(meas, radial_step=0.2, masks=masks)(fields)
    b = keep[0].astype(bool)
    for k in keys:
        assert masked[k].shape == (N, m)
        assert np.array_equal(masked[k].data[:, b], full[k].data[:, b]), k
        assert (masked[k].data[:, ~b] == 0).all(), k
    # the masked likelihood sees the same numbers either way
    lk = api.B200EnsembleLikelihood(meas, mask_dict=masks)
    assert lk(masked) == lk(full)
    partial = api.Masks()
    partial.append(name=keys[0], data=keep, otype='HEALPix')
    with pytest.raises(AssertionError):
        api.B200Simulator(meas, masks=partial)
