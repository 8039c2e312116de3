"""
GPU: the drop-ins inside the reference's OWN `Pipeline` class (imagine/pipelines/pipeline.py, unmodified):
field factories, priors, the prior transform, ensemble seeds and `_likelihood_function` are the reference's;
only the Field / Simulator / Likelihood classes are swapped.  The samplers IMAGINE wraps (MultiNest, UltraNest,
dynesty, emcee) are not installed in this image, so `call()` is a few Metropolis steps driving the SAME
callback a sampler adapter would (`Pipeline._likelihood_function`, ultranest_pipeline.py:160).
"""
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    import torch
    assert torch.cuda.is_available()
    import imagine_b200.api as api
    from imagine_b200 import _lib
    _lib.init(0)
    import imagine as img
    return api, img


def _metropolis_pipeline(img):
    class MetropolisPipeline(img.pipelines.Pipeline):
        """Minimal sampler adapter: unit-cube random walk on `_likelihood_function`."""
        SUPPORTS_MPI = False

        def call(self, n_steps=20, step=0.08, seed=0, start=None, **kwargs):
            rng = np.random.default_rng(seed)
            n = len(self.active_parameters)
            cube = np.full(n, 0.5) if start is None else np.asarray(start, dtype=float)
            loglike = lambda c: self._likelihood_function(self.prior_transform(c))   # what a sampler does
            cur = loglike(cube)
            chain = [(cube.copy(), cur)]
            for _ in range(n_steps):
                prop = np.clip(cube + step * rng.normal(size=n), 0.0, 1.0)
                val = loglike(prop)
                if np.log(rng.uniform()) < val - cur:
                    cube, cur = prop, val
                chain.append((cube.copy(), cur))
            # what the reference's sampler adapters leave behind (e.g. ultranest_pipeline.py:186-190)
            self._samples_array = np.array([self.prior_transform(c) for c, _ in chain])
            self._evidence, self._evidence_err = float('nan'), float('nan')
            self.results = chain
            return chain
    return MetropolisPipeline


def _tabular_measurements(img, u, x, values, err):
    ds = img.observables.TabularDataset(
        {'meas': u.Quantity(values, u.microgauss * u.cm ** -3), 'err': np.full_like(values, err), 'x': x,
         'y': np.zeros_like(values), 'z': np.zeros_like(values)}, name='test', data_col='meas', err_col='err')
    return img.observables.Measurements(ds)


def test_reference_pipeline_gives_the_same_likelihood_with_gpu_dropins(env):
    """Deterministic fields (exact parity is possible): the reference Pipeline with the reference's
    TestSimulator + EnsembleLikelihood vs the same Pipeline with the GPU classes, at several cubes."""
    api, img = env
    u = api.units
    import imagine.fields as RF
    from imagine.simulators import TestSimulator
    x = np.linspace(0.01, 2 * np.pi - 0.01, 10)
    meas = _tabular_measurements(img, u, x, 3.0 * (1 + np.cos(x)), 0.1)
    grid = img.fields.UniformGrid(box=[[0, 2 * np.pi] * u.kpc, [0, 0] * u.kpc, [0, 0] * u.kpc],
                                  resolution=[100, 1, 1])
    ne_params = {'n0': 1 * u.cm ** -3, 'a': 1 * u.rad / u.kpc, 'alpha': 0 * u.rad, 'b': 0 * u.rad / u.kpc,
                 'beta': np.pi / 2 * u.rad, 'c': 0 * u.rad / u.kpc, 'gamma': np.pi / 2 * u.rad}
    b_defaults = {'Bx': 0 * u.microgauss, 'By': 3 * u.microgauss, 'Bz': 0 * u.microgauss}
    priors = {'By': img.priors.FlatPrior(xmin=0 * u.microgauss, xmax=6 * u.microgauss),
              'Bz': img.priors.FlatPrior(xmin=-2 * u.microgauss, xmax=2 * u.microgauss)}

    def build(ne_cls, b_cls, sim_cls, lik_cls, rundir):
        ne_f = img.fields.FieldFactory(field_class=ne_cls, grid=grid, default_parameters=ne_params)
        b_f = img.fields.FieldFactory(field_class=b_cls, grid=grid, default_parameters=b_defaults,
                                      active_parameters=('By', 'Bz'), priors=priors)
        return _metropolis_pipeline(img)(simulator=sim_cls(meas), factory_list=[ne_f, b_f],
                                        likelihood=lik_cls(meas), ensemble_size=4, run_directory=rundir,
                                        show_summary_reports=False)

    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
        ref = build(RF.CosThermalElectronDensity, RF.ConstantMagneticField, TestSimulator,
                    img.likelihoods.EnsembleLikelihood, d1)
        gpu = build(api.CosThermalElectronDensity, api.ConstantMagneticField, api.B200TestSimulator,
                    api.B200EnsembleLikelihood, d2)
        assert gpu.active_parameters == ref.active_parameters
        for cube in ([0.5, 0.5], [0.1, 0.9], [0.77, 0.25], [1.0, 0.0]):
            want = ref._likelihood_function(ref.prior_transform(np.array(cube)))
            got = gpu._likelihood_function(gpu.prior_transform(np.array(cube)))
            assert abs(got - want) <= 1e-8 * max(1.0, abs(want) * 1e-8), (cube, got, want)
        # the truth (By = 3 uG = cube 0.5) beats a wrong field strength
        ll = lambda c: gpu._likelihood_function(gpu.prior_transform(np.array(c)))
        assert ll([0.5, 0.5]) > ll([0.9, 0.5]) and ll([0.5, 0.5]) > ll([0.1, 0.5])


def test_four_parameter_gmf_fit_through_the_reference_pipeline(env):
    """BASELINE config 5 in spirit: WMAP LSA regular field (b0, psi0, psi1, chi0 free) + Gaussian random
    field, Faraday depth + synchrotron maps from B200Simulator, B200EnsembleLikelihood, all inside the
    reference's Pipeline (priors, prior transform, seeds); the random walk must climb towards the truth."""
    api, img = env
    u = api.units
    nside, res, N = 4, [24, 24, 12], 8
    m = 12 * nside ** 2
    grid = img.fields.UniformGrid(box=[[-20, 20] * u.kpc, [-20, 20] * u.kpc, [-4, 4] * u.kpc], resolution=res)
    truth = {'b0': 3.0 * u.microgauss, 'psi0': 27.0 * u.deg, 'psi1': 0.9 * u.deg, 'chi0': 25.0 * u.deg}
    ne = api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.02 * u.cm ** -3,
                                                           'scale_radius': 5 * u.kpc, 'scale_height': 1 * u.kpc},
                                         ensemble_size=1)
    cre = api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5 * u.kpc, 'z0': 1 * u.kpc,
                                                           'r_sun': 8.3 * u.kpc}, ensemble_size=1)
    keys = [('fd', None, nside, None), ('sync', 23.0, nside, 'I'), ('sync', 23.0, nside, 'Q'),
            ('sync', 23.0, nside, 'U')]
    units = [u.rad / u.m ** 2, u.K, u.K, u.K]
    blank = api.Measurements()
    for k, un in zip(keys, units):
        blank.append(name=k, data=np.zeros((1, m)) * un, otype='HEALPix')
    sim0 = api.B200Simulator(blank, radial_step=0.25)
    mock = sim0([ne, cre, api.LSAMagneticField(grid, parameters=truth, ensemble_size=1)])
    rng = np.random.default_rng(5)
    meas = api.Measurements()
    for k, un in zip(keys, units):
        row = mock[k].data[0]
        sigma = 0.05 * np.abs(row).mean()
        meas.append(name=k, data=(row + sigma * rng.normal(size=m)).reshape(1, -1) * un,
                    cov_data=np.full(m, sigma ** 2) * un ** 2, otype='HEALPix')
    fac = lambda obj: img.fields.FieldFactory(field_class=obj)      # fixed fields: factories from objects
    lsa = img.fields.FieldFactory(
        field_class=api.LSAMagneticField, grid=grid, default_parameters=truth,
        active_parameters=('b0', 'psi0', 'psi1', 'chi0'),
        priors={'b0': img.priors.FlatPrior(xmin=0 * u.microgauss, xmax=6 * u.microgauss),
                'psi0': img.priors.FlatPrior(xmin=0 * u.deg, xmax=54 * u.deg),
                'psi1': img.priors.FlatPrior(xmin=-2 * u.deg, xmax=4 * u.deg),
                'chi0': img.priors.FlatPrior(xmin=0 * u.deg, xmax=50 * u.deg)})
    brnd = img.fields.FieldFactory(field_class=api.NaiveGaussianMagneticField, grid=grid,
                                   default_parameters={'a0': 0 * u.microgauss, 'b0': 0.5 * u.microgauss})
    with tempfile.TemporaryDirectory() as rundir:
        pipe = _metropolis_pipeline(img)(simulator=api.B200Simulator(meas, radial_step=0.25),
                                         factory_list=[fac(ne), fac(cre), lsa, brnd],
                                         likelihood=api.B200EnsembleLikelihood(meas), ensemble_size=N,
                                         run_directory=rundir, show_summary_reports=False)
        assert len(pipe.active_parameters) == 4
        ll = lambda c: pipe._likelihood_function(pipe.prior_transform(np.array(c)))
        at_truth = ll([0.5, 0.5, 0.4833333333333333, 0.5])
        far = ll([0.95, 0.05, 0.9, 0.05])
        assert np.isfinite(at_truth) and at_truth > far
        # several cubes per call (what a vectorised sampler hands over): fields, priors, seeds and rescaler
        # are the pipeline's own, the points share one ray march over (point, realisation) lanes
        vec = api.VectorizedLikelihood(pipe)
        cubes = np.array([pipe.prior_transform(np.array(c)) for c in
                          ([0.5, 0.5, 0.48, 0.5], [0.95, 0.05, 0.9, 0.05], [0.3, 0.6, 0.2, 0.7],
                           [0.6, 0.4, 0.5, 0.5], [0.1, 0.9, 0.4, 0.2])])
        one_by_one = np.array([pipe._likelihood_function(c) for c in cubes])
        together = vec(cubes)
        assert vec._ev.batch_capacity() == 4 and together.shape == (5,)
        assert np.allclose(together, one_by_one, rtol=1e-9, atol=1e-8), (together, one_by_one)
        chain = pipe(n_steps=25, step=0.1, seed=1, start=[0.8, 0.2, 0.7, 0.2], save_pipeline_state=False)
        assert chain[-1][1] > chain[0][1]                    # the walk climbs
        assert max(v for _, v in chain) <= at_truth + 50.0   # and nothing beats the truth by a wide margin


def test_likelihood_convergence_report_stays_on_the_device(env):
    """Pipeline.prepare_likelihood_convergence_report (pipelines/pipeline.py:1035-1124) with the drop-ins:
    the reference's own loop -- simulate max_Nens * n_seeds realisations once, `Simulations.sub_sim` for
    every ensemble size, likelihood + bootstrap dispersion for each -- where `sub_sim` hands out device
    row subsets (no host copy of the maps is ever made).  Deterministic fields: identical to the
    all-reference pipeline, NumPy draw for NumPy draw; Gaussian fields: a non-zero dispersion."""
    api, img = env
    u = api.units
    import imagine.fields as RF
    from imagine.simulators import TestSimulator
    x = np.linspace(0.01, 2 * np.pi - 0.01, 10)
    meas = _tabular_measurements(img, u, x, 3.0 * (1 + np.cos(x)), 0.1)
    grid = img.fields.UniformGrid(box=[[0, 2 * np.pi] * u.kpc, [0, 0] * u.kpc, [0, 0] * u.kpc],
                                  resolution=[100, 1, 1])
    ne_params = {'n0': 1 * u.cm ** -3, 'a': 1 * u.rad / u.kpc, 'alpha': 0 * u.rad, 'b': 0 * u.rad / u.kpc,
                 'beta': np.pi / 2 * u.rad, 'c': 0 * u.rad / u.kpc, 'gamma': np.pi / 2 * u.rad}

    def build(ne_cls, b_cls, b_defaults, active, priors, sim_cls, lik_cls, rundir):
        ne_f = img.fields.FieldFactory(field_class=ne_cls, grid=grid, default_parameters=ne_params)
        b_f = img.fields.FieldFactory(field_class=b_cls, grid=grid, default_parameters=b_defaults,
                                      active_parameters=active, priors=priors)
        return _metropolis_pipeline(img)(simulator=sim_cls(meas), factory_list=[ne_f, b_f],
                                        likelihood=lik_cls(meas), ensemble_size=4, run_directory=rundir,
                                        show_summary_reports=False)

    const = dict(b_defaults={'Bx': 0 * u.microgauss, 'By': 3 * u.microgauss, 'Bz': 0 * u.microgauss},
                 active=('By',), priors={'By': img.priors.FlatPrior(xmin=0 * u.microgauss, xmax=6 * u.microgauss)})
    kw = dict(min_Nens=3, max_Nens=5, n_seeds=2, n_points=2)
    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2, tempfile.TemporaryDirectory() as d3:
        ref = build(RF.CosThermalElectronDensity, RF.ConstantMagneticField, sim_cls=TestSimulator,
                    lik_cls=img.likelihoods.EnsembleLikelihood, rundir=d1, **const)
        gpu = build(api.CosThermalElectronDensity, api.ConstantMagneticField, sim_cls=api.B200TestSimulator,
                    lik_cls=api.B200EnsembleLikelihood, rundir=d2, **const)
        np.random.seed(11)
        want = ref.prepare_likelihood_convergence_report(**kw)
        np.random.seed(11)
        got = gpu.prepare_likelihood_convergence_report(**kw)
        assert list(got['ensemble_size']) == list(want['ensemble_size']) and len(got) == 2 * 2 * 3
        assert np.allclose(got['likelihood'], want['likelihood'], rtol=0, atol=1e-8)
        assert np.allclose(got['likelihood_std'], want['likelihood_std'], rtol=0, atol=1e-8)
        assert gpu.ensemble_size == 4 and gpu.likelihood.compute_dispersion is False      # settings restored

        gauss = dict(b_defaults={'a0': 3 * u.microgauss, 'b0': 0.5 * u.microgauss}, active=('a0',),
                     priors={'a0': img.priors.FlatPrior(xmin=0 * u.microgauss, xmax=6 * u.microgauss)})
        rnd = build(api.CosThermalElectronDensity, api.NaiveGaussianMagneticField, sim_cls=api.B200TestSimulator,
                    lik_cls=api.B200EnsembleLikelihood, rundir=d3, **gauss)
        seen = []
        inner = rnd._get_observables
        rnd._get_observables = lambda values: seen.append(inner(values)) or seen[-1]
        report = rnd.prepare_likelihood_convergence_report(min_Nens=4, max_Nens=8, n_seeds=1, n_points=1)
        assert np.all(np.isfinite(report['likelihood'])) and np.all(report['likelihood_std'] > 0)
        key = next(iter(seen[0].keys()))
        sub = seen[0].sub_sim(slice(0, 3))[key]
        assert type(sub).__name__ == 'DeviceObservable' and sub.shape[0] == 3
        assert seen[0][key]._host is None               # the maps never left the device during the report
