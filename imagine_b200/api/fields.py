"""
GPU-aware Field plugins: same NAME / PARAMETER_NAMES / TYPE / UNITS as the reference classes,
plus `device_descriptor()`, which lets `B200Simulator` evaluate the whole ensemble in one fused
launch instead of calling `get_data` once per realisation.  `compute_field` stays valid for any
other (host) Simulator: it evaluates ONE realisation on the GPU and returns a host Quantity --
there is no NumPy implementation behind it.

Reference classes mirrored:
  ConstantMagneticField            fields/basic_fields.py:16-40
  ConstantThermalElectrons         fields/basic_fields.py:43-56
  ExponentialThermalElectrons      fields/basic_fields.py:59-88
  RandomThermalElectrons           fields/basic_fields.py:91-127
  CosThermalElectronDensity        fields/test_field.py:19-43
  NaiveGaussianMagneticField       fields/test_field.py:68-96
  AnalyticCosmicRayElectrons, ConstantCosmicRayElectrons
                                   new (reference CRE class is a stub, base_fields.py:104-113);
                                   parameter names from fields/hamx/cre_analytic.py:19-27
Stochastic fields draw from Philox4x32-10 keyed on the realisation seed (never on the GPU rank),
so they match the reference's MT19937 stream in distribution, not bit for bit.
"""
import zlib

import numpy as np

from . import units as u
from ._bases import Field, MagneticField, ThermalElectronDensityField
from .. import _lib as L

FLUX_UNIT = (u.GeV * u.m ** 2 * u.s * u.sr) ** (-1)


class CosmicRayElectronDensityField(Field):
    """
    Gridded cosmic-ray electron flux normalisation at the reference energy, [GeV m^2 s sr]^-1.
    The reference's class of this name is a stub whose constructor raises (base_fields.py:104-113);
    the TYPE string is the one the Hammurabi wrapper anticipates (hammurabi.py:58-60).
    """
    TYPE = 'cosmic_ray_electron_density'
    UNITS = FLUX_UNIT

    @property
    def data_description(self):
        return ['grid_x', 'grid_y', 'grid_z']

    @property
    def data_shape(self):
        return tuple(self.grid.shape)


def _stream_id(name):
    return zlib.crc32(name.encode()) & 0x7fffffff


def device_grid(grid):
    """engine.DeviceGrid for a (mirror or reference) UniformGrid; cached on the object."""
    from .. import engine
    dg = getattr(grid, '_imb_device_grid', None)
    if dg is None:
        if getattr(grid, 'grid_type', 'cartesian') != 'cartesian':
            raise ValueError('only cartesian UniformGrids are supported on the GPU path')
        box = [[float(u.value_in(b, u.kpc)) for b in row] for row in grid.box]
        dg = engine.DeviceGrid(box, [int(n) for n in grid.resolution])
        try:
            grid._imb_device_grid = dg
        except AttributeError:
            pass
    return dg


class _DeviceFieldMixin:
    """Shared machinery: descriptor -> single-realisation evaluation for host consumers."""
    KIND = None
    PARAM_UNITS = ()          # canonical unit of each entry of PARAMETER_NAMES

    def _canonical_parameters(self):
        vals = []
        params = self.parameters             # the reference's getter validates on every access: once
        for name, unit in zip(self.PARAMETER_NAMES, self.PARAM_UNITS):
            p = params[name]
            if isinstance(p, u.Quantity) and p.unit is unit:
                p = p.value                  # already canonical (the usual case): no conversion
            elif isinstance(p, u.Quantity):
                try:
                    p = p.to_value(unit)
                except Exception:            # e.g. 1/kpc given where rad/kpc is expected
                    try:
                        p = p.to_value(unit, u.dimensionless_angles())
                    except Exception:
                        # FieldFactory wraps bare numbers as dimensionless Quantities
                        # (field_factory.py:210): a bare number means "in canonical units"
                        if p.unit != u.dimensionless_unscaled:
                            raise
                        p = p.value
            vals.append(float(p))
        return vals

    def device_descriptor(self):
        """(kind, params[<=8], philox_stream) in canonical units (kpc, uG, cm^-3, rad)."""
        return (self.KIND, self._canonical_parameters(), _stream_id(self.NAME))

    def compute_field(self, seed):
        from .. import engine
        dg = device_grid(self.grid)
        out = engine.field_eval(dg, [self.device_descriptor()], [0 if seed is None else int(seed)],
                                n_real=1, b_stride=3)
        (arr,) = out.values()
        host = arr.cpu().numpy().reshape(self.data_shape)
        return host * self.UNITS

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_deterministic_data'] = None
        return state


class ConstantMagneticField(_DeviceFieldMixin, MagneticField):
    NAME = 'constant_B'
    PARAMETER_NAMES = ['Bx', 'By', 'Bz']
    PARAM_UNITS = (u.microgauss,) * 3
    KIND = L.FIELD_CONST_B


class NaiveGaussianMagneticField(_DeviceFieldMixin, MagneticField):
    NAME = 'naive_gaussian_magnetic_field'
    STOCHASTIC_FIELD = True
    PARAMETER_NAMES = ['a0', 'b0']
    PARAM_UNITS = (u.microgauss,) * 2
    KIND = L.FIELD_GAUSS_B


class LSAMagneticField(_DeviceFieldMixin, MagneticField):
    """
    WMAP-3yr logarithmic-spiral-arm regular field evaluated on the grid: the gridded counterpart of
    the `BregLSA` dummy (fields/hamx/breg_lsa.py:11-31: NAME 'breg_lsa', parameters b0 [uG],
    psi0, psi1, chi0 [deg]) -- the 4-parameter GMF model of BASELINE config 5.
    """
    NAME = 'breg_lsa'
    PARAMETER_NAMES = ['b0', 'psi0', 'psi1', 'chi0']
    PARAM_UNITS = (u.microgauss, u.deg, u.deg, u.deg)
    KIND = L.FIELD_LSA_B
    R0_KPC, Z0_KPC, R_MIN_KPC, R_MAX_KPC = 8.0, 1.0, 3.0, 20.0

    def _canonical_parameters(self):
        b0, psi0, psi1, chi0 = super()._canonical_parameters()       # plain numbers are degrees
        d2r = np.pi / 180.0
        return [b0, psi0 * d2r, psi1 * d2r, chi0 * d2r, self.R0_KPC, self.Z0_KPC, self.R_MIN_KPC,
                self.R_MAX_KPC]


class JF12MagneticField(_DeviceFieldMixin, MagneticField):
    """
    Jansson & Farrar (2012) regular Galactic magnetic field evaluated on the grid: the gridded
    counterpart of the `BregJF12` dummy (fields/hamx/breg_jf12.py:9-39: NAME 'breg_jf12', the same 20
    parameter names, defaults of `BregJF12Factory`).
    """
    NAME = 'breg_jf12'
    PARAMETER_NAMES = ['b_arm1', 'b_arm2', 'b_arm3', 'b_arm4', 'b_arm5', 'b_arm6', 'b_arm7', 'b_ring',
                       'h_disk', 'w_disk', 'Bn', 'Bs', 'rn', 'rs', 'wh', 'z0', 'B0_X', 'Xtheta', 'rpc_X', 'r0_X']
    PARAM_UNITS = ((u.microgauss,) * 8 + (u.kpc, u.kpc, u.microgauss, u.microgauss) + (u.kpc,) * 4
                   + (u.microgauss, u.deg, u.kpc, u.kpc))
    KIND = L.FIELD_JF12_B
    DEFAULTS = {'b_arm1': 0.1, 'b_arm2': 3.0, 'b_arm3': -0.9, 'b_arm4': -0.8, 'b_arm5': -2.0, 'b_arm6': -4.2,
                'b_arm7': 0.0, 'b_ring': 0.1, 'h_disk': 0.40, 'w_disk': 0.27, 'Bn': 1.4, 'Bs': -1.1, 'rn': 9.22,
                'rs': 16.7, 'wh': 0.20, 'z0': 5.3, 'B0_X': 4.6, 'Xtheta': 49, 'rpc_X': 4.8, 'r0_X': 2.9}


class ConstantThermalElectrons(_DeviceFieldMixin, ThermalElectronDensityField):
    NAME = 'constant_TE'
    PARAMETER_NAMES = ['ne']
    PARAM_UNITS = (u.cm ** -3,)
    KIND = L.FIELD_CONST_NE


class ExponentialThermalElectrons(_DeviceFieldMixin, ThermalElectronDensityField):
    NAME = 'exponential_disc_thermal_electrons'
    PARAMETER_NAMES = ['central_density', 'scale_radius', 'scale_height']
    PARAM_UNITS = (u.cm ** -3, u.kpc, u.kpc)
    KIND = L.FIELD_EXP_NE


class RandomThermalElectrons(_DeviceFieldMixin, ThermalElectronDensityField):
    NAME = 'random_thermal_electrons'
    STOCHASTIC_FIELD = True
    PARAMETER_NAMES = ['mean', 'std', 'min_ne']
    PARAM_UNITS = (u.cm ** -3,) * 3
    KIND = L.FIELD_RANDOM_NE


class CosThermalElectronDensity(_DeviceFieldMixin, ThermalElectronDensityField):
    NAME = 'cos_therm_electrons'
    PARAMETER_NAMES = ['n0', 'a', 'alpha', 'b', 'beta', 'c', 'gamma']
    PARAM_UNITS = (u.cm ** -3, u.rad / u.kpc, u.rad, u.rad / u.kpc, u.rad, u.rad / u.kpc, u.rad)
    KIND = L.FIELD_COS_NE


class AnalyticCosmicRayElectrons(_DeviceFieldMixin, CosmicRayElectronDensityField):
    """j0 exp(-(R - r_sun)/r0) sech^2(z/z0): CRE flux normalisation at E0 (WMAP-style disc)."""
    NAME = 'cre_analytic'
    PARAMETER_NAMES = ['j0', 'r0', 'z0', 'r_sun']
    PARAM_UNITS = (FLUX_UNIT, u.kpc, u.kpc, u.kpc)
    KIND = L.FIELD_CRE_ANALYTIC


class ConstantCosmicRayElectrons(_DeviceFieldMixin, CosmicRayElectronDensityField):
    NAME = 'cre_constant'
    PARAMETER_NAMES = ['j0']
    PARAM_UNITS = (FLUX_UNIT,)
    KIND = L.FIELD_CRE_CONST


def descriptor_of(field):
    """Descriptor of a GPU-aware field, or None for a foreign (host-only) Field plugin."""
    fn = getattr(field, 'device_descriptor', None)
    return fn() if callable(fn) else None
