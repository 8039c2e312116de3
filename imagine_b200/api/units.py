"""
Unit handling for the host-side API mirror.

IMAGINE's plugin interface passes `astropy.units.Quantity` objects (field parameters, grid
boxes, coordinates, observables).  When astropy is importable it is used as is.  It is not
installed in the build/GPU image, so this module carries a deliberately small stand-in with
the same spelling for the subset the hot path touches: `value`, `unit`, `to`, `to_value`,
`<<`, `*`, `/`, `**`, `is_equivalent`, indexing and `shape`.  Units are (scale, dimension
vector) pairs over the base (cm, gauss, K, rad, s, GeV, sr).
"""
import numpy as np

try:                                     # pragma: no cover - astropy absent in this image
    import astropy.units as _au
    HAVE_ASTROPY = True
except Exception:
    _au = None
    HAVE_ASTROPY = False

_NDIM = 7


class Unit:
    __array_priority__ = 1000

    def __init__(self, scale=1.0, dims=None, name=None):
        self.scale = float(scale)
        self.dims = tuple(dims) if dims is not None else (0,) * _NDIM
        self.name = name

    # algebra ------------------------------------------------------------------------------
    def _combine(self, other, sign):
        return Unit(self.scale * other.scale ** sign,
                    tuple(a + sign * b for a, b in zip(self.dims, other.dims)))

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, 1)
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * self)
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.value, self / other.unit)
        return Quantity(1.0 / np.asarray(other, dtype=np.float64), self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.scale, tuple(-d for d in self.dims))
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * inv)
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit(self.scale ** p, tuple(d * p for d in self.dims))

    def __eq__(self, other):
        return (isinstance(other, Unit) and self.dims == other.dims
                and np.isclose(self.scale, other.scale, rtol=1e-12, atol=0))

    def __hash__(self):
        return hash(self.dims)

    def is_equivalent(self, other, equivalencies=None):
        other = as_unit(other)
        if equivalencies == 'dimensionless_angles':
            strip = lambda d: tuple(0 if i == 3 else v for i, v in enumerate(d))
            return strip(self.dims) == strip(other.dims)
        return self.dims == other.dims

    def to(self, other):
        """Conversion factor self -> other."""
        other = as_unit(other)
        if self.dims != other.dims:
            raise UnitConversionError('%r and %r are not convertible' % (self, other))
        return self.scale / other.scale

    def __repr__(self):
        return self.name or 'Unit(%g, %s)' % (self.scale, self.dims)


class UnitConversionError(ValueError):
    pass


def _d(i, p=1):
    return tuple(p if k == i else 0 for k in range(_NDIM))


dimensionless_unscaled = Unit(1.0, name='')
cm = Unit(1.0, _d(0), 'cm')
m = Unit(1.0e2, _d(0), 'm')
pc = Unit(3.0856775814913673e18, _d(0), 'pc')
kpc = Unit(3.0856775814913673e21, _d(0), 'kpc')
gauss = Unit(1.0, _d(1), 'G')
microgauss = Unit(1.0e-6, _d(1), 'uG')
K = Unit(1.0, _d(2), 'K')
rad = Unit(1.0, _d(3), 'rad')
deg = Unit(np.pi / 180.0, _d(3), 'deg')
s = Unit(1.0, _d(4), 's')
GHz = Unit(1.0e9, _d(4, -1), 'GHz')
GeV = Unit(1.0, _d(5), 'GeV')
sr = Unit(1.0, _d(6), 'sr')


def dimensionless_angles():
    return 'dimensionless_angles'


def as_unit(u):
    if u is None:
        return dimensionless_unscaled
    if isinstance(u, Unit):
        return u
    if isinstance(u, Quantity):
        return Unit(u.unit.scale * float(u.value), u.unit.dims)
    raise TypeError('not a unit: %r' % (u,))


class Quantity:
    """value (numpy array or scalar) x Unit."""
    __array_priority__ = 2000

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit_in = value.unit
            value = value.value if unit is None else value.to_value(unit)
            unit = unit_in if unit is None else unit
        elif isinstance(value, (list, tuple)) and any(isinstance(v, Quantity) for v in value):
            unit = unit or value[0].unit
            value = [Quantity(v).to_value(unit) for v in value]
        self.value = np.asarray(value, dtype=np.float64) if not np.isscalar(value) else float(value)
        self.unit = as_unit(unit)

    # conversions ----------------------------------------------------------------------------
    def to_value(self, unit=None, equivalencies=None):
        if unit is None:
            return self.value
        unit = as_unit(unit)
        if equivalencies == 'dimensionless_angles':
            strip = lambda u: Unit(u.scale, tuple(0 if i == 3 else v for i, v in enumerate(u.dims)))
            return self.value * strip(self.unit).to(strip(unit))
        f = self.unit.to(unit)
        return self.value if f == 1.0 else self.value * f

    def to(self, unit, equivalencies=None):
        return Quantity(self.to_value(unit, equivalencies), unit)

    def __lshift__(self, unit):
        return self.to(unit)

    def __rlshift__(self, other):                     # ndarray << Quantity-unit is not used
        return NotImplemented

    # container behaviour ------------------------------------------------------------------
    @property
    def shape(self):
        return np.shape(self.value)

    @property
    def size(self):
        return np.size(self.value)

    @property
    def T(self):
        return Quantity(np.transpose(self.value), self.unit)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, idx):
        return Quantity(self.value[idx], self.unit)

    def __setitem__(self, idx, val):
        self.value[idx] = val.to_value(self.unit) if isinstance(val, Quantity) else val

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __float__(self):
        return float(self.to_value(dimensionless_unscaled)) if self.unit.dims == (0,) * _NDIM else float(self.value)

    def copy(self):
        return Quantity(np.copy(self.value), self.unit)

    # arithmetic ----------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        if isinstance(other, Quantity):
            return Quantity(self.value * other.value, self.unit * other.unit)
        return Quantity(self.value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        if isinstance(other, Quantity):
            return Quantity(self.value / other.value, self.unit / other.unit)
        return Quantity(self.value / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / self.value, dimensionless_unscaled / self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def _other_value(self, other):
        if isinstance(other, Quantity):
            return other.to_value(self.unit)
        if self.unit.dims != (0,) * _NDIM and np.any(np.asarray(other) != 0):
            raise UnitConversionError('cannot combine a dimensional Quantity with a plain number')
        return other

    def __add__(self, other):
        return Quantity(self.value + self._other_value(other), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.value - self._other_value(other), self.unit)

    def __rsub__(self, other):
        return Quantity(self._other_value(other) - self.value, self.unit)

    def __eq__(self, other):
        if isinstance(other, Quantity):
            if other.unit.dims != self.unit.dims:
                return False
            return self.value == other.to_value(self.unit)
        return self.value == other

    def __lt__(self, other):
        return self.value < self._other_value(other)

    def __gt__(self, other):
        return self.value > self._other_value(other)

    def __abs__(self):
        return Quantity(np.abs(self.value), self.unit)

    def __repr__(self):
        return '<Quantity %r %r>' % (self.value, self.unit)


if HAVE_ASTROPY:                          # pragma: no cover
    Quantity = _au.Quantity               # noqa: F811
    Unit = _au.UnitBase                   # noqa: F811
    UnitConversionError = _au.UnitConversionError   # noqa: F811
    dimensionless_unscaled = _au.dimensionless_unscaled
    cm, m, pc, kpc, gauss, microgauss, K, rad, deg, s, GHz, GeV, sr = (
        _au.cm, _au.m, _au.pc, _au.kpc, _au.gauss, _au.microgauss, _au.K, _au.rad, _au.deg,
        _au.s, _au.GHz, _au.GeV, _au.sr)
    dimensionless_angles = _au.dimensionless_angles

    def as_unit(u):                       # noqa: F811
        return _au.dimensionless_unscaled if u is None else _au.Unit(u)


def unit_scale(from_unit, to_unit):
    """Multiplicative factor converting a value in `from_unit` into `to_unit`."""
    if HAVE_ASTROPY:                      # pragma: no cover
        return float((1.0 * from_unit).to_value(to_unit))
    return as_unit(from_unit).to(to_unit)


def value_in(q, unit):
    """Plain float/ndarray of `q` (Quantity or number) expressed in `unit`."""
    if isinstance(q, Quantity):
        return q.to_value(unit)
    return q
