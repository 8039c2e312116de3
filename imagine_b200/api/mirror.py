"""
Host-side mirror of the part of IMAGINE's plugin interface that the hot path touches.

The reference package (`imagine`) cannot be imported in the build / GPU image (astropy, mpi4py,
e13tools ... absent; `np.int` at fields/grid.py:59).  When it *is* importable the drop-ins in
this package subclass the real classes (see api/__init__.py); otherwise they subclass the
classes below, which restate the same names, argument meaning and error behaviour:

  UniformGrid                      fields/grid.py:189-275
  Field, MagneticField, ThermalElectronDensityField, CosmicRayElectronDensityField
                                   fields/field.py:19-219, fields/base_fields.py:34-113
  Observable                       observables/observable.py:43-211
  Measurements, Simulations, Covariances, Masks
                                   observables/observable_dict.py:52-406
  Simulator                        simulators/simulator.py:18-422
  Likelihood                       likelihoods/likelihood.py:37-172

Only host bookkeeping lives here (dictionaries, shape/unit assertions, key conventions).  All
arithmetic of the hot path is in the CUDA library; nothing in this module evaluates a field,
integrates a line of sight or factorises a matrix.
"""
import abc

import numpy as np

from . import units as u


# ----------------------------------------------------------------------------------- grid
class UniformGrid:
    """Vertex-centred, end-inclusive uniform grid (np.mgrid semantics, grid.py:256-260)."""

    def __init__(self, box, resolution, grid_type='cartesian'):
        self.box = box
        self.resolution = np.empty((3,), dtype=np.int64)
        self.resolution[:] = resolution
        self.grid_type = grid_type
        self._coordinates = None

    @property
    def shape(self):
        return self.resolution

    def box_values(self, unit=u.kpc):
        """[[lo, hi]] x 3 as floats in `unit` (rows may carry their own units)."""
        return np.array([[float(u.value_in(b, unit)) for b in row] for row in self.box])

    def _axis(self, a):
        lo, hi = self.box_values()[a]
        n = int(self.resolution[a])
        step = (hi - lo) / float(n - 1) if n != 1 else 1.0
        return np.arange(n, dtype=np.float64) * step + lo

    @property
    def coordinates(self):
        if self._coordinates is None:
            if self.grid_type != 'cartesian':
                raise NotImplementedError('only cartesian grids are on the hot path '
                                          '(test_simulator.py:35, hammurabi.py:61)')
            ax = [self._axis(a) for a in range(3)]
            full = np.meshgrid(*ax, indexing='ij')
            self._coordinates = {k: v * u.kpc for k, v in zip('xyz', full)}
        return self._coordinates

    x = property(lambda self: self.coordinates['x'])
    y = property(lambda self: self.coordinates['y'])
    z = property(lambda self: self.coordinates['z'])

    @property
    def r_cylindrical(self):
        c = self.coordinates
        return np.sqrt(c['x'].value ** 2 + c['y'].value ** 2) * u.kpc

    @property
    def r_spherical(self):
        c = self.coordinates
        return np.sqrt(c['x'].value ** 2 + c['y'].value ** 2 + c['z'].value ** 2) * u.kpc


# ---------------------------------------------------------------------------------- fields
def _class_attr(obj, name):
    if not hasattr(obj, name):
        raise AttributeError('Required class attribute %r is not defined!' % name)
    return getattr(obj, name)


class Field(metaclass=abc.ABCMeta):
    """Field base: parameters, ensemble seeds, cached deterministic data (field.py:43-219)."""

    def __init__(self, grid, *, parameters={}, ensemble_size=None, ensemble_seeds=None, dependencies={}):
        for attr in ('TYPE', 'NAME', 'UNITS', 'PARAMETER_NAMES'):
            _class_attr(self, attr)
        self.grid = grid
        self._parameters = {}
        self.parameters = parameters
        if ensemble_size is None and ensemble_seeds is None:
            ensemble_size = 1
        self.ensemble_size = ensemble_size
        self.ensemble_seeds = ensemble_seeds
        self._deterministic_data = None
        self.dependencies = dependencies
        assert self.type not in self.dependencies_list, 'Field cannot depend on its own field type'

    type = property(lambda self: self.TYPE)
    name = property(lambda self: self.NAME)
    units = property(lambda self: self.UNITS)
    parameter_names = property(lambda self: self.PARAMETER_NAMES)
    dependencies_list = property(lambda self: getattr(self, 'DEPENDENCIES_LIST', []))
    stochastic_field = property(lambda self: getattr(self, 'STOCHASTIC_FIELD', False))

    @property
    @abc.abstractmethod
    def data_description(self):
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def data_shape(self):
        raise NotImplementedError

    @abc.abstractmethod
    def compute_field(self, seed):
        raise NotImplementedError

    def get_data(self, i_realization=0, dependencies={}):
        if self.stochastic_field:
            assert i_realization < self.ensemble_size
            self._update_dependencies(dependencies)
            field = self.compute_field(self.ensemble_seeds[i_realization])
            self._check_realisation(field)
        elif self._deterministic_data is None:
            self._update_dependencies(dependencies)
            field = self.compute_field(None)
            self._check_realisation(field)
            self._deterministic_data = field
        else:
            field = self._deterministic_data
        return field

    def _update_dependencies(self, dependencies):
        for dep in self.dependencies_list:
            if dep not in dependencies:
                raise KeyError('Missing field dependency {}'.format(dep))
        self.dependencies = dependencies

    def _check_realisation(self, field):
        assert self.units.is_equivalent(field.unit), 'Field units should be %r' % self.units
        assert tuple(field.shape) == tuple(self.data_shape), \
            'Incorrect shape %s, it should be %s' % (tuple(field.shape), tuple(self.data_shape))

    @property
    def ensemble_seeds(self):
        return self._ensemble_seeds

    @ensemble_seeds.setter
    def ensemble_seeds(self, seeds):
        if seeds is None:
            self._ensemble_seeds = np.random.randint(0, 2 ** 31 - 1, self.ensemble_size)
        else:
            if self.ensemble_size is None:
                self.ensemble_size = len(seeds)
            assert len(seeds) == self.ensemble_size
            self._ensemble_seeds = seeds

    @property
    def parameters(self):
        return self._parameters

    @parameters.setter
    def parameters(self, parameters):
        for k in parameters:
            assert k in self.parameter_names
        self._parameters.update(parameters)


class MagneticField(Field):
    TYPE = 'magnetic_field'
    UNITS = u.microgauss
    data_description = property(lambda self: ['grid_x', 'grid_y', 'grid_z', 'component (x,y,z)'])
    data_shape = property(lambda self: (*self.grid.shape, 3))


class ThermalElectronDensityField(Field):
    TYPE = 'thermal_electron_density'
    UNITS = u.cm ** (-3)
    data_description = property(lambda self: ['grid_x', 'grid_y', 'grid_z'])
    data_shape = property(lambda self: tuple(self.grid.shape))


class CosmicRayElectronDensityField(Field):
    """
    The reference class is a stub that raises NotImplementedError (base_fields.py:104-113);
    the TYPE string is the one the Hammurabi wrapper anticipates (hammurabi.py:58-60).  Here:
    gridded CRE flux normalisation at the reference energy, in [GeV m^2 s sr]^-1.
    """
    TYPE = 'cosmic_ray_electron_density'
    UNITS = (u.GeV * u.m ** 2 * u.s * u.sr) ** (-1)
    data_description = property(lambda self: ['grid_x', 'grid_y', 'grid_z'])
    data_shape = property(lambda self: tuple(self.grid.shape))


# ----------------------------------------------------------------------------- observables
class Observable:
    """[rows, size] float64 block + dtype/otype/unit bookkeeping (observable.py:58-211)."""

    def __init__(self, data=None, dtype=None, coords=None, otype=None):
        self.dtype = dtype
        self.otype = otype
        if isinstance(data, u.Quantity):
            self.data = np.asarray(data.value, dtype=np.float64)
            self.unit = data.unit
        elif isinstance(data, np.ndarray):
            self.data = data
            self.unit = None
        else:
            raise ValueError
        self.coords = coords
        self.rw_flag = False

    @property
    def dtype(self):
        return self._dtype

    @dtype.setter
    def dtype(self, dtype):
        if dtype is None:
            raise ValueError('dtype cannot be None')
        assert dtype in ('measured', 'simulated', 'variance', 'covariance')
        self._dtype = str(dtype)

    @property
    def data(self):
        if self._dtype == 'variance':
            return np.diag(self._data)      # a stored variance still reads back as a covariance
        return self._data

    @data.setter
    def data(self, data):
        if data is None:
            self._data = None
            return
        assert isinstance(data, np.ndarray)
        if self._dtype != 'variance':
            assert len(data.shape) == 2
        if self._dtype == 'measured':
            assert data.shape[0] == 1
        self._data = np.copy(data)
        if self._dtype == 'covariance':
            assert data.shape[0] == data.shape[1]

    @property
    def var(self):
        if self._dtype == 'variance':
            return self._data
        if self._dtype == 'covariance':
            return self._data.diagonal()
        raise TypeError("Needs 'variance' or 'covariance' observables for this")

    @property
    def shape(self):
        return self._data.shape

    @property
    def global_data(self):
        return self.data

    @property
    def size(self):
        return self._data.shape[1]

    @property
    def ensemble_mean(self):
        if self._dtype == 'measured':
            assert self._data.shape[0] == 1
            return self._data
        if self._dtype == 'simulated':
            return np.mean(self._data, axis=0).reshape(1, -1)
        raise TypeError('unsupported data type')

    @property
    def rw_flag(self):
        return self._rw_flag

    @rw_flag.setter
    def rw_flag(self, flag):
        assert isinstance(flag, bool)
        self._rw_flag = flag

    def append(self, new_data):
        assert isinstance(new_data, (np.ndarray, Observable, u.Quantity))
        assert self._dtype == 'simulated'
        if isinstance(new_data, u.Quantity):
            assert self.unit == new_data.unit
            new_data = np.asarray(new_data.value)
        rows = new_data.data if isinstance(new_data, Observable) else new_data
        if self._rw_flag:
            self._data = np.copy(rows)
            self._rw_flag = False
        else:
            self._data = np.vstack([self._data, rows])


def _npix(nside):
    return 12 * int(nside) ** 2


class ObservableDict(metaclass=abc.ABCMeta):
    """Key convention (name, freq_GHz|None, Nside|'tab'|size, tag|None), observable_dict.py:11-25."""

    def __init__(self, *datasets):
        self._archive = {}
        for d in datasets:
            self.append(dataset=d)

    archive = property(lambda self: self._archive)

    def keys(self):
        return self._archive.keys()

    def __iter__(self):
        return iter(self._archive)

    def __getitem__(self, key):
        return self._archive[key]

    def __len__(self):
        return len(self._archive)

    @staticmethod
    def _unpack(dataset, name, data, cov_data, otype, coords):
        if dataset is not None:
            name, data, otype, coords = dataset.key, dataset.data, dataset.otype, dataset.coords
            cov_data = dataset.cov if dataset.cov is not None else dataset.var
        assert len(name) == 4, 'Wrong format for Observable key!'
        return name, data, cov_data, otype, coords

    @abc.abstractmethod
    def append(self, dataset=None, *, name=None, data=None, cov_data=None, otype=None, coords=None):
        raise NotImplementedError


class Measurements(ObservableDict):
    def __init__(self, *datasets):
        self.cov = None
        super().__init__(*datasets)

    def append(self, dataset=None, *, name=None, data=None, cov_data=None, otype=None, coords=None):
        name, data, cov, otype, coords = self._unpack(dataset, name, data, cov_data, otype, coords)
        if cov is not None:
            if self.cov is None:
                self.cov = Covariances()
            self.cov.append(name=name, cov_data=cov, otype=otype, coords=coords)
        if isinstance(data, Observable):
            assert data.dtype == 'measured'
            self._archive[name] = data
        elif isinstance(data, (np.ndarray, u.Quantity)):
            if otype == 'HEALPix':
                assert data.shape[1] == _npix(name[2])
            self._archive[name] = Observable(data=data, dtype='measured', coords=coords, otype=otype)
        else:
            raise TypeError('Unsupported data type')


class Simulations(ObservableDict):
    def append(self, dataset=None, *, name=None, data=None, cov_data=None, otype=None, coords=None):
        name, data, _, otype, coords = self._unpack(dataset, name, data, cov_data, otype, coords)
        if name in self._archive:
            self._archive[name].rw_flag = False
            self._archive[name].append(data)
        elif isinstance(data, Observable):
            self._archive[name] = data
        elif isinstance(data, (np.ndarray, u.Quantity)):
            self._archive[name] = Observable(data=data, dtype='simulated', coords=coords, otype=otype)
        else:
            raise TypeError('unsupported data type')

    def sub_sim(self, indices):
        sub = Simulations()
        for k in self:
            obs = self[k]
            if hasattr(obs, 'take_rows'):                      # device-resident block
                sub.append(name=k, data=obs.take_rows(indices), otype=obs.otype)
            else:
                sub.append(name=k, data=obs.global_data[indices, :], otype=obs.otype)
        return sub


class Covariances(ObservableDict):
    def append(self, dataset=None, *, name=None, data=None, cov_data=None, otype=None, coords=None):
        name, _, data, otype, coords = self._unpack(dataset, name, data, cov_data, otype, coords)
        if isinstance(data, Observable):
            self._archive[name] = data
        elif isinstance(data, (np.ndarray, u.Quantity)):
            if len(data.shape) == 2:
                if otype == 'HEALPix':
                    assert data.shape[1] == _npix(name[2])
                self._archive[name] = Observable(data, dtype='covariance', coords=coords, otype=otype)
            else:
                self._archive[name] = Observable(data, dtype='variance', coords=coords, otype=otype)
        else:
            raise TypeError('unsupported data type')


class Masks(ObservableDict):
    """0 = masked, 1 = kept; masked entries are renamed (name, freq, n_kept, tag), otype 'plain'."""

    def append(self, dataset=None, *, name=None, data=None, cov_data=None, otype=None, coords=None):
        name, data, _, otype, _ = self._unpack(dataset, name, data, cov_data, otype, coords)
        if isinstance(data, Observable):
            assert data.dtype == 'measured'
            if otype == 'HEALPix':
                assert data.size == _npix(name[2])
            self._archive[name] = data
        elif isinstance(data, np.ndarray):
            assert data.shape[0] == 1
            if otype == 'HEALPix':
                assert data.shape[1] == _npix(name[2])
            self._archive[name] = Observable(data, 'measured')
        else:
            raise TypeError('unsupported data type')

    def kept_indices(self, name):
        return np.flatnonzero(self._archive[name].data[0].astype(bool)).astype(np.int32)

    def masked_key(self, name):
        return (name[0], name[1], int(self.kept_indices(name).size), name[3])

    def __call__(self, observable_dict):
        """Host application to Measurements / Covariances / host Simulations (setup-time path)."""
        assert isinstance(observable_dict, (Measurements, Simulations, Covariances))
        out = type(observable_dict)()
        is_cov = isinstance(observable_dict, Covariances)
        for name, obs in observable_dict._archive.items():
            if name not in self._archive:
                out.append(name=name, **({'cov_data': obs} if is_cov else {'data': obs}))
                continue
            keep = self.kept_indices(name)
            if is_cov:
                masked = obs.var[keep] if obs.dtype == 'variance' else obs.data[np.ix_(keep, keep)]
                out.append(name=self.masked_key(name), cov_data=masked, otype='plain')
            else:
                if hasattr(obs, 'take_columns'):
                    out.append(name=self.masked_key(name), data=obs.take_columns(keep), otype='plain')
                else:
                    out.append(name=self.masked_key(name), data=obs.data[:, keep], otype='plain')
            out[self.masked_key(name)].unit = obs.unit
            out[self.masked_key(name)].coords = obs.coords
        return out


# ------------------------------------------------------------------------------- simulator
class Simulator(metaclass=abc.ABCMeta):
    """Simulator base (simulator.py:47-422): observables registry, field preparation, loop."""

    def __init__(self, measurements):
        for attr in ('SIMULATED_QUANTITIES', 'REQUIRED_FIELD_TYPES', 'ALLOWED_GRID_TYPES'):
            _class_attr(self, attr)
        self.grid = None
        self.grids = None
        self.fields = None
        self.field_parameters = {}
        self.observables = []
        self.output_coords = {}
        self.output_units = {}
        self.output_type = {}
        self._ensemble_size = None
        self.register_observables(measurements)

    simulated_quantities = property(lambda self: self.SIMULATED_QUANTITIES)
    required_field_types = property(lambda self: self.REQUIRED_FIELD_TYPES)
    optional_field_types = property(lambda self: getattr(self, 'OPTIONAL_FIELD_TYPES', []))
    allowed_grid_types = property(lambda self: self.ALLOWED_GRID_TYPES)
    use_common_grid = property(lambda self: getattr(self, 'USE_COMMON_GRID', True))
    ensemble_size = property(lambda self: self._ensemble_size)

    def register_observables(self, measurements):
        assert isinstance(measurements, Measurements)
        for key in measurements.keys():
            if key[0] in self.simulated_quantities:
                self.observables.append(key)
                self.output_coords[key] = measurements[key].coords
                self.output_type[key] = measurements[key].otype
                self.output_units[key] = measurements[key].unit
        assert len(self.observables) > 0, 'No valid observable was requested!'

    def register_ensemble_size(self, field_list):
        sizes = {f.ensemble_size for f in field_list}
        assert len(sizes) == 1, 'All fields should have the same ensemble size'
        self._ensemble_size = sizes.pop()

    def _accepted(self, field_list):
        """Fields this simulator consumes, after the grid checks of simulator.py:128-147."""
        wanted = (*self.required_field_types, *self.optional_field_types)
        self.grid, self.grids = None, None
        used = []
        for f in field_list:
            if f.type not in wanted:
                continue
            if f.grid is not None and self.allowed_grid_types is not None:
                assert f.grid.grid_type in self.allowed_grid_types, 'Grid type not allowed'
            if self.use_common_grid:
                if self.grid is None:
                    self.grid = f.grid
                assert self.grid is f.grid, 'Multiple grids when a common grid is required'
            used.append(f)
        assert set(self.required_field_types) <= {f.type for f in used}, 'Missing required field'
        return used

    def prepare_fields(self, field_list, i):
        """Host path: evaluate realisation i of every field and sum same-type fields."""
        self.fields = {}
        for f in self._accepted(field_list):
            deps = {d: self.fields[d] for d in f.dependencies_list if isinstance(d, str)}
            data = f.get_data(i, deps)
            if f.type not in self.fields:
                self.fields[f.type] = data
                self.field_parameters[f.type] = f.parameter_names
            else:
                self.fields[f.type] = self.fields[f.type] + data      # out of place, like the reference

    @abc.abstractmethod
    def simulate(self, key, coords_dict, realization_id, output_units):
        raise NotImplementedError

    def __call__(self, field_list):
        sims = Simulations()
        self.register_ensemble_size(field_list)
        for i in range(self._ensemble_size):
            self.prepare_fields(field_list, i)
            for key in self.observables:
                sim = self.simulate(key=key, coords_dict=self.output_coords[key], realization_id=i,
                                    output_units=self.output_units[key])
                row = sim[np.newaxis, :]
                if self.output_units[key] is not None:
                    row = row.to(self.output_units[key])
                sims.append(name=key, data=row, coords=self.output_coords[key],
                            otype=self.output_type[key])
        return sims


# ------------------------------------------------------------------------------ likelihood
class Likelihood(metaclass=abc.ABCMeta):
    """Likelihood base (likelihood.py:65-172): masks, covariance lookup, bootstrap dispersion."""

    def __init__(self, measurement_dict, covariance_dict=None, mask_dict=None,
                 compute_dispersion=False, n_bootstrap=150):
        self._check_units(measurement_dict, covariance_dict)
        self.mask_dict = mask_dict
        self.measurement_dict = measurement_dict
        if covariance_dict is None:
            covariance_dict = measurement_dict.cov
        self.covariance_dict = covariance_dict
        self.compute_dispersion = compute_dispersion
        self.n_bootstrap = n_bootstrap

    def __call__(self, observable_dict, **kwargs):
        if self.mask_dict is not None:
            observable_dict = self.mask_dict(observable_dict)
        value = self.call(observable_dict, **kwargs)
        if not self.compute_dispersion:
            return value
        sample = [self._bootstrapped_likelihood(observable_dict, **kwargs)
                  for _ in range(self.n_bootstrap)]
        return value, np.std(sample)

    def _bootstrapped_likelihood(self, simulations, **kwargs):
        size = simulations[next(iter(simulations.keys()))].shape[0]
        idx = np.random.randint(0, size, size)
        return self.call(simulations.sub_sim(idx), **kwargs)

    @property
    def mask_dict(self):
        return self._mask_dict

    @mask_dict.setter
    def mask_dict(self, mask_dict):
        if mask_dict is not None:
            assert isinstance(mask_dict, Masks)
        self._mask_dict = mask_dict

    @property
    def measurement_dict(self):
        return self._measurement_dict

    @measurement_dict.setter
    def measurement_dict(self, measurement_dict):
        assert isinstance(measurement_dict, Measurements)
        self._measurement_dict = measurement_dict
        if self._mask_dict is not None:
            self._measurement_dict = self._mask_dict(self._measurement_dict)

    @property
    def covariance_dict(self):
        return self._covariance_dict

    @covariance_dict.setter
    def covariance_dict(self, covariance_dict):
        if covariance_dict is not None:
            assert isinstance(covariance_dict, Covariances)
        self._covariance_dict = covariance_dict
        if self._mask_dict is not None and self._covariance_dict is not None:
            self._covariance_dict = self._mask_dict(self._covariance_dict)

    @abc.abstractmethod
    def call(self, observable_dict):
        raise NotImplementedError

    @staticmethod
    def _check_units(measurements, covariances):
        if covariances is None:
            return
        for k in measurements:
            if measurements[k].unit is None:
                assert covariances[k].unit is None
            else:
                assert (measurements[k].unit) ** 2 == covariances[k].unit
