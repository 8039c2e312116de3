"""
not-gpu: the drop-ins are built on the REFERENCE's own classes.  `imagine_b200.reference`
imports the unmodified package (baseline/_ref or /root/reference) with stand-ins only for absent
third-party modules; everything below runs the reference's code, not a restatement of it.
"""
import pickle

import numpy as np
import pytest

import imagine_b200.api as api
from imagine_b200 import reference

u = api.units


def test_reference_package_is_the_backend():
    import imagine
    assert api.BACKEND in ('installed', 'sources+standins')
    assert imagine.__version__.startswith('2.')
    assert reference.root() is not None
    import imagine.simulators.simulator as rs
    import imagine.likelihoods.likelihood as rl
    import imagine.fields.base_fields as rb
    import imagine.observables.observable as ro
    assert api.B200Simulator.__mro__[1] is rs.Simulator and api.B200TestSimulator.__mro__[1] is rs.Simulator
    for cls in (api.B200EnsembleLikelihood, api.B200EnsembleLikelihoodDiagonal, api.B200SimpleLikelihood):
        assert issubclass(cls, rl.Likelihood)
    assert issubclass(api.ConstantMagneticField, rb.MagneticField)
    assert issubclass(api.ExponentialThermalElectrons, rb.ThermalElectronDensityField)
    assert issubclass(api.DeviceObservable, ro.Observable)


def _measurements(nside=2, with_sync=True):
    m = 12 * nside ** 2
    rng = np.random.default_rng(0)
    M = api.Measurements()
    M.append(name=('fd', None, nside, None), data=rng.normal(size=(1, m)) * (u.rad / u.m ** 2),
             cov_data=np.full(m, 0.25) * (u.rad / u.m ** 2) ** 2, otype='HEALPix')
    if with_sync:
        M.append(name=('sync', 23.0, nside, 'I'), data=rng.normal(size=(1, m)) * u.K,
                 cov_data=np.full(m, 0.5) * u.K ** 2, otype='HEALPix')
    return M


def test_simulator_and_likelihoods_construct_on_reference_bases_without_a_gpu():
    M = _measurements()
    sim = api.B200Simulator(M)                     # imagine.simulators.Simulator.__init__ registers the keys
    assert set(sim.observables) == set(M.keys())
    assert sim.output_units[('sync', 23.0, 2, 'I')] == u.K
    assert sim.simulated_quantities == ['fd', 'dm', 'sync']
    lk = api.B200EnsembleLikelihood(M)             # imagine.likelihoods.Likelihood.__init__ + unit checks
    assert lk.covariance_dict is M.cov
    bad = api.Covariances()
    bad.append(name=('fd', None, 2, None), cov_data=np.full(48, 0.25) * u.K ** 2, otype='HEALPix')
    bad.append(name=('sync', 23.0, 2, 'I'), cov_data=np.full(48, 0.5) * u.K ** 2, otype='HEALPix')
    with pytest.raises(AssertionError):            # likelihood.py:161-172 _check_units, the reference's own
        api.B200EnsembleLikelihood(M, bad)
    with pytest.raises(NotImplementedError):
        api.B200EnsembleLikelihood(M, cov_func=lambda x: x)
    # pickling (Pipeline.save, tools/io.py) keeps working with the reference bases
    sim2, lk2 = pickle.loads(pickle.dumps(sim)), pickle.loads(pickle.dumps(lk))
    assert set(sim2.observables) == set(M.keys()) and lk2.estimator == 'oas'


def test_gpu_fields_are_reference_fields():
    grid = api.UniformGrid(box=[[-1, 1] * u.kpc] * 3, resolution=[4, 4, 4])
    f = api.ConstantMagneticField(grid, parameters={'Bx': 1 * u.microgauss, 'By': 2e-6 * u.gauss, 'Bz': 0 * u.microgauss})
    assert f.type == 'magnetic_field' and f.data_shape == (4, 4, 4, 3) and f.units == u.microgauss
    kind, params, _ = f.device_descriptor()
    assert params == pytest.approx([1.0, 2.0, 0.0])
    g = api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 1 * u.microgauss},
                                       ensemble_size=3, ensemble_seeds=[5, 6, 7])
    assert g.stochastic_field and list(g.ensemble_seeds) == [5, 6, 7]
    with pytest.raises(AssertionError):            # field.py:213-218: unknown parameter names are rejected by the base
        api.ConstantMagneticField(grid, parameters={'Bx': 1 * u.microgauss, 'Bw': 1 * u.microgauss})


def test_reference_likelihood_consumes_a_device_observable():
    """`Simulations.append(data=<Observable>)` stores the object as given (observable_dict.py:306-307):
    the reference's own EnsembleLikelihood reads a DeviceObservable through `.data` (here backed by a
    host tensor, so the test runs without a GPU) and gets the same number as from plain arrays."""
    import torch
    import imagine.likelihoods as RL
    M = _measurements(with_sync=False)
    key = ('fd', None, 2, None)
    rows = np.random.default_rng(1).normal(size=(6, 48))
    S_dev, S_host = api.Simulations(), api.Simulations()
    S_dev.append(name=key, data=api.DeviceObservable(torch.from_numpy(rows.copy()), u.rad / u.m ** 2, None, 'HEALPix'))
    S_host.append(name=key, data=rows, otype='HEALPix')
    assert S_dev[key].shape == (6, 48) and np.array_equal(S_dev[key].data, rows)
    lk = RL.EnsembleLikelihood(M)
    assert lk(S_dev) == lk(S_host)
    assert RL.EnsembleLikelihoodDiagonal(M)(S_dev) == RL.EnsembleLikelihoodDiagonal(M)(S_host)
    assert np.allclose(S_dev.sub_sim([0, 0, 3])[key].data, rows[[0, 0, 3]]) if hasattr(S_dev, 'sub_sim') else True


def test_reference_fields_reproduce_the_tutorial_goldens_with_the_units_stand_in():
    """The reference's own grid / field code, unmodified (tutorials/tutorial_fields.ipynb:93, 249-256)."""
    import imagine.fields as RF
    grid = RF.UniformGrid(box=[[-15, 15] * u.kpc] * 3, resolution=[15, 15, 15])
    assert float(grid.r_spherical[5, 5, 5].to_value(u.kpc)) == 7.423074889580904
    ne = RF.ExponentialThermalElectrons(grid, parameters={'central_density': 1.0 * u.cm ** -3,
                                                          'scale_radius': 3.3 * u.kpc, 'scale_height': 3.3 * u.kpc})
    got = np.asarray(ne.get_data()[3:5, 3:5, 3:5].to_value(u.cm ** -3)).ravel()
    want = [0.00189093, 0.00361981, 0.00289678, 0.0055453, 0.00289678, 0.0055453, 0.00473694, 0.00906791]
    assert np.allclose(got, want, rtol=0, atol=5e-9)


def test_simulator_can_report_thermodynamic_temperature():
    """hammurabiX (and the reference's key convention, observable_dict.py:15-19) reports K-cmb; the kernels
    integrate Rayleigh-Jeans brightness temperature.  `temperature='cmb'` applies the conversion to the
    synchrotron maps through the same per-key scale as unit conversions."""
    import imagine_b200.api as api
    from imagine_b200.constants import rj_to_cmb
    u = api.units
    M = api.Measurements()
    keys = [('fd', None, 2, None), ('sync', 23.0, 2, 'I'), ('sync', 23.0, 2, 'Q'), ('sync', 1.4, 2, 'PA')]
    units = [u.rad / u.m ** 2, u.K, u.K, u.rad]
    for k, un in zip(keys, units):
        M.append(name=k, data=np.zeros((1, 48)) * un, otype='HEALPix')
    rj, cmb = api.B200Simulator(M, temperature='rj'), api.B200Simulator(M)          # K_CMB is the default
    assert [rj._scale(k) for k in keys] == [1.0, 1.0, 1.0, 1.0]
    assert [cmb._scale(k) for k in keys] == [1.0, rj_to_cmb(23.0), rj_to_cmb(23.0), 1.0]
    assert abs(rj_to_cmb(23.0) - 1.01375) < 1e-5 and abs(rj_to_cmb(1.4) - 1.00005) < 1e-5
    with pytest.raises(ValueError):
        api.B200Simulator(M, temperature='antenna')
