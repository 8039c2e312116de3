"""
not-gpu: the host-side mirror of the reference plugin interface (containers, key conventions,
assertions).  Modelled on the reference's imagine/tests/test_observable.py,
test_observabledict.py and test_templates.py:40-77.
"""
import pickle

import numpy as np
import pytest

import imagine_b200.api as M
from imagine_b200.api import units as u


def test_basedict_is_abstract():
    with pytest.raises(TypeError):
        __import__("imagine.observables").observables.ObservableDict()


def test_observable_semantics():
    arr = np.random.rand(1, 128)
    obs = M.Observable(arr, 'measured')
    assert obs.shape == (1, 128) and obs.size == 128 and obs.dtype == 'measured'
    assert np.array_equal(obs.ensemble_mean, arr)
    with pytest.raises(ValueError):
        M.Observable(arr, None)
    with pytest.raises(ValueError):
        M.Observable([1, 2, 3], 'measured')
    with pytest.raises(AssertionError):
        M.Observable(np.random.rand(2, 3), 'measured')          # measured = single row
    with pytest.raises(AssertionError):
        M.Observable(np.random.rand(2, 3), 'covariance')        # must be square
    sim = M.Observable(np.random.rand(2, 5), 'simulated')
    sim.append(np.random.rand(3, 5))
    assert sim.shape == (5, 5)
    sim.rw_flag = True
    sim.append(np.random.rand(1, 5))
    assert sim.shape == (1, 5) and sim.rw_flag is False
    var = M.Observable(np.arange(1.0, 4.0), 'variance')
    assert np.array_equal(var.data, np.diag([1.0, 2.0, 3.0]))   # observable.py:91-94
    assert np.array_equal(var.var, [1.0, 2.0, 3.0])
    q = M.Observable(np.ones((1, 3)) * u.K, 'measured')
    assert q.unit == u.K and isinstance(q.data, np.ndarray)


def test_measurements_simulations_covariances():
    meas = M.Measurements()
    arr, hrr = np.random.rand(1, 3), np.random.rand(1, 48)
    meas.append(name=('test', None, 3, None), data=arr, otype='plain')
    meas.append(name=('test', None, 2, None), data=hrr, otype='HEALPix')
    assert np.array_equal(meas[('test', None, 2, None)].data, hrr)
    with pytest.raises(AssertionError):
        meas.append(name=('test', None, 4, None), data=hrr, otype='HEALPix')   # 12*Nside^2 check
    with pytest.raises(AssertionError):
        meas.append(name=('bad', None), data=arr)
    sims = M.Simulations()
    sims.append(name=('test', None, 3, None), data=np.random.rand(2, 3), otype='plain')
    sims.append(name=('test', None, 3, None), data=np.random.rand(2, 3), otype='plain')
    assert sims[('test', None, 3, None)].shape == (4, 3)
    sub = sims.sub_sim([0, 0, 3])
    assert sub[('test', None, 3, None)].shape == (3, 3)
    covs = M.Covariances()
    covs.append(name=('test', None, 3, None), cov_data=np.random.rand(3, 3))
    covs.append(name=('v', None, 3, None), cov_data=np.random.rand(3))
    assert covs[('test', None, 3, None)].dtype == 'covariance' and covs[('v', None, 3, None)].dtype == 'variance'
    m2 = M.Measurements()
    m2.append(name=('fd', None, 2, None), data=hrr, cov_data=np.ones(48), otype='HEALPix')
    assert m2.cov[('fd', None, 2, None)].dtype == 'variance'


def test_masks_rename_keys_like_the_reference():
    key = ('test', None, 4, None)
    meas, cov, mask = M.Measurements(), M.Covariances(), M.Masks()
    meas.append(name=key, data=np.array([[1, 2, 3, 4.]]), otype='plain')
    mask.append(name=key, data=np.array([[1, 1, 0, 0.]]), otype='plain')
    cov.append(name=key, cov_data=np.diag((1, 2, 3, 4.)), otype='plain')
    assert np.array_equal(mask(meas)[('test', None, 2, None)].data, [[1., 2.]])
    assert np.array_equal(mask(cov)[('test', None, 2, None)].data, [[1., 0.], [0., 2.]])
    var = M.Covariances()
    var.append(name=key, cov_data=np.array([1, 2, 3, 4.]), otype='plain')
    assert np.array_equal(mask(var)[('test', None, 2, None)].var, [1., 2.])
    other = M.Measurements()
    other.append(name=('other', None, 4, None), data=np.ones((1, 4)), otype='plain')
    assert ('other', None, 4, None) in mask(other).keys()            # unmasked entries pass through


def test_uniform_grid_and_field_contract():
    grid = M.UniformGrid(box=[[0, 4] * u.kpc, [1, 2] * u.kpc, [1, 1] * u.kpc], resolution=[5, 2, 1])
    assert np.array_equal(grid.x.value[:, 0, 0], [0, 1, 2, 3, 4])
    assert np.array_equal(grid.y.value[0, :, 0], [1, 2])
    assert tuple(grid.shape) == (5, 2, 1)

    class Bz(M.MagneticField):                               # test_templates.py:40-60 analogue
        NAME = 'bz'
        PARAMETER_NAMES = ['amp']

        def compute_field(self, seed):
            B = np.zeros(self.data_shape)
            B[..., 2] = self.parameters['amp'] * np.sign(self.grid.x.value - 2)
            return B * u.microgauss

    f = Bz(grid, parameters={'amp': 42.0})
    d = f.get_data()
    assert d.value[0, 0, 0, 2] == -42 and d.value[4, 1, 0, 2] == 42
    assert f.get_data() is d                                 # deterministic cache (field.py:162-171)
    with pytest.raises(AssertionError):
        Bz(grid, parameters={'nope': 1})

    class Bad(M.ThermalElectronDensityField):
        NAME = 'bad'
        PARAMETER_NAMES = []

        def compute_field(self, seed):
            return np.ones((2, 2, 2)) * u.cm ** -3

    with pytest.raises(AssertionError):
        Bad(grid).get_data()                                 # shape check (field.py:181-191)

    class NoName(M.MagneticField):
        PARAMETER_NAMES = []

        def compute_field(self, seed):
            return None

    with pytest.raises(AttributeError):
        NoName(grid)                                         # req_attr (class_tools.py:33-47)
    f2 = Bz(grid, parameters={'amp': 1.0}, ensemble_size=4)
    assert len(f2.ensemble_seeds) == 4


def test_likelihood_base_unit_check_and_covariance_lookup():
    class L0(M.Likelihood):
        def call(self, observable_dict):
            return 0.0

    meas = M.Measurements()
    meas.append(name=('fd', None, 1, None), data=np.ones((1, 12)) * (u.rad / u.m ** 2),
                cov_data=np.ones(12) * (u.rad / u.m ** 2) ** 2, otype='HEALPix')
    lk = L0(meas)
    assert lk.covariance_dict is meas.cov
    bad = M.Covariances()
    bad.append(name=('fd', None, 1, None), cov_data=np.ones(12) * u.K ** 2, otype='HEALPix')
    with pytest.raises(AssertionError):
        L0(meas, bad)


def test_units_stand_in():
    q = [0, 2 * np.pi] * u.kpc
    assert q.to_value(u.pc)[1] == pytest.approx(2e3 * np.pi)
    assert (3 * u.microgauss).to_value(u.gauss) == pytest.approx(3e-6)
    assert (u.microgauss * u.cm ** -3).is_equivalent(u.gauss / u.cm ** 3)
    assert not u.K.is_equivalent(u.kpc)
    with pytest.raises(u.UnitConversionError):
        (1 * u.K).to(u.kpc)
    assert ((1 * u.rad / u.kpc) * (2 * u.kpc)).to_value(u.rad) == pytest.approx(2.0)
    assert u.unit_scale(u.pc / u.cm ** 3, u.kpc / u.cm ** 3) == pytest.approx(1e-3)


def test_gpu_field_descriptors_without_gpu():
    from imagine_b200.api import fields as F
    from imagine_b200 import _lib
    grid = M.UniformGrid(box=[[-1, 1] * u.kpc] * 3, resolution=[4, 4, 4])
    f = F.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3,
                                                        'scale_radius': 3000 * u.pc,
                                                        'scale_height': 0.1 * u.kpc})
    kind, p, stream = f.device_descriptor()
    assert kind == _lib.FIELD_EXP_NE and p == pytest.approx([0.01, 3.0, 0.1])
    g = F.NaiveGaussianMagneticField(grid, parameters={'a0': 1 * u.microgauss, 'b0': 2e-6 * u.gauss},
                                     ensemble_size=3, ensemble_seeds=[1, 2, 3])
    assert g.stochastic_field and g.device_descriptor()[1] == pytest.approx([1.0, 2.0])
    assert g.NAME == 'naive_gaussian_magnetic_field' and g.PARAMETER_NAMES == ['a0', 'b0']
    c = F.CosThermalElectronDensity(grid, parameters={'n0': 1 * u.cm ** -3, 'a': 1 * u.rad / u.kpc,
                                                      'alpha': 0 * u.rad, 'b': 0 / u.kpc, 'beta': 90 * u.deg,
                                                      'c': 0 * u.rad / u.kpc, 'gamma': np.pi / 2})
    assert c.device_descriptor()[1] == pytest.approx([1, 1, 0, 0, np.pi / 2, 0, np.pi / 2])
    pickle.loads(pickle.dumps(g))                            # checkpointable (io.py:25-75)


def test_hammurabi_field_dump_layout(tmp_path):
    """Raw C-order float64, component innermost: the format hammurabi.py:267 writes with tofile."""
    from imagine_b200.api.simulators import dump_hammurabi_fields
    B = np.arange(2 * 3 * 4 * 3, dtype=np.float64).reshape(2, 3, 4, 3) * u.microgauss
    ne = np.arange(24, dtype=np.float64).reshape(2, 3, 4) * u.cm ** -3
    out = dump_hammurabi_fields({'magnetic_field': B, 'thermal_electron_density': ne, 'dummy': {}}, str(tmp_path))
    assert set(out) == {'breg', 'tereg'}
    back = np.fromfile(out['breg'][0], dtype=np.float64)
    assert np.array_equal(back, np.arange(72.0)) and out['breg'][1] == {'nx': 2, 'ny': 3, 'nz': 4}
    assert np.array_equal(np.fromfile(out['tereg'][0]), np.arange(24.0))
