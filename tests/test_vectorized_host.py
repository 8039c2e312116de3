"""
CPU: host logic of api.VectorizedLikelihood -- the fields of every parameter point come from the reference
Pipeline's own `_get_observables` (a recorder stands in for the simulator), their descriptors and seeds are
what `EnsembleEvaluator.evaluate_batch` consumes.  No kernel runs here; the numerical side is
tests/test_gpu_pipeline.py.
"""
import tempfile

import numpy as np
import pytest


@pytest.fixture(scope='module')
def env():
    import imagine_b200.api as api
    import imagine as img
    return api, img


def _pipeline(api, img, rundir, with_masks=False, likelihood_cls=None):
    u = api.units
    nside, N = 2, 4
    m = 12 * nside ** 2
    grid = img.fields.UniformGrid(box=[[-20, 20] * u.kpc, [-20, 20] * u.kpc, [-4, 4] * u.kpc], resolution=[4, 4, 4])
    meas = api.Measurements()
    for k, un in ((('fd', None, nside, None), u.rad / u.m ** 2), (('sync', 23.0, nside, 'I'), u.K)):
        meas.append(name=k, data=np.ones((1, m)) * un, cov_data=np.full(m, 0.1) * un ** 2, otype='HEALPix')
    ne = api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.02 * u.cm ** -3,
                                                           'scale_radius': 5 * u.kpc, 'scale_height': 1 * u.kpc},
                                         ensemble_size=1)
    cre = api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5 * u.kpc, 'z0': 1 * u.kpc,
                                                           'r_sun': 8.3 * u.kpc}, ensemble_size=1)
    lsa = img.fields.FieldFactory(
        field_class=api.LSAMagneticField, grid=grid,
        default_parameters={'b0': 3 * u.microgauss, 'psi0': 27 * u.deg, 'psi1': 0.9 * u.deg, 'chi0': 25 * u.deg},
        active_parameters=('b0', 'psi0'),
        priors={'b0': img.priors.FlatPrior(xmin=0 * u.microgauss, xmax=6 * u.microgauss),
                'psi0': img.priors.FlatPrior(xmin=0 * u.deg, xmax=54 * u.deg)})
    brnd = img.fields.FieldFactory(field_class=api.NaiveGaussianMagneticField, grid=grid,
                                   default_parameters={'a0': 0 * u.microgauss, 'b0': 0.5 * u.microgauss})
    fac = lambda obj: img.fields.FieldFactory(field_class=obj)

    class OneShot(img.pipelines.Pipeline):
        SUPPORTS_MPI = False

        def call(self, **kwargs):
            return None
    lik = (likelihood_cls or api.B200EnsembleLikelihood)(meas)
    return OneShot(simulator=api.B200Simulator(meas, radial_step=0.5), factory_list=[fac(ne), fac(cre), lsa, brnd],
                   likelihood=lik, ensemble_size=N, run_directory=rundir, show_summary_reports=False)


def test_points_are_built_by_the_pipeline_itself(env):
    api, img = env
    with tempfile.TemporaryDirectory() as rundir:
        pipe = _pipeline(api, img, rundir)
        sim = pipe._simulator
        vec = api.VectorizedLikelihood(pipe)
        assert vec.quantities == ('fd', 'I') and vec.nside == 2 and vec.settings.nu_ghz == 23.0
        values = pipe.prior_transform(np.array([0.5, 0.25]))
        descs, seeds, grid, n_real = vec._point(values)
        assert pipe._simulator is sim                                   # the recorder is gone again
        assert n_real == 4 and seeds.shape == (1, 4)
        kinds = [d[0] for d in descs]
        from imagine_b200 import _lib as L
        assert kinds == [L.FIELD_EXP_NE, L.FIELD_CRE_ANALYTIC, L.FIELD_LSA_B, L.FIELD_GAUSS_B]
        lsa = descs[2][1]
        assert lsa[0] == pytest.approx(3.0) and lsa[1] == pytest.approx(np.deg2rad(13.5))   # cube -> prior -> radians
        assert descs[3][3] == 0 and descs[2][3] == 0                    # the one seed row belongs to the Gaussian field
        # the same ensemble seeds as an ordinary evaluation of the pipeline would use
        assert list(seeds[0]) == list(pipe.ensemble_seeds[pipe._factory_list[3]])
        other = vec._point(pipe.prior_transform(np.array([0.9, 0.75])))[0]
        assert other[2][1][0] == pytest.approx(5.4) and other[0] == descs[0]      # only the active parameters moved


def test_scope_is_checked_up_front(env):
    api, img = env
    with tempfile.TemporaryDirectory() as rundir:
        pipe = _pipeline(api, img, rundir, likelihood_cls=api.B200EnsembleLikelihoodDiagonal)
        with pytest.raises(NotImplementedError):
            api.VectorizedLikelihood(pipe)
        pipe = _pipeline(api, img, rundir)
        pipe.likelihood.compute_dispersion = True
        with pytest.raises(NotImplementedError):
            api.VectorizedLikelihood(pipe)
