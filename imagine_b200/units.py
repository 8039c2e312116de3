"""
Unit handling of the host-side plugin layer.

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
    # imagine_b200.reference registers THIS module as `astropy.units` when astropy is absent
    HAVE_ASTROPY = not getattr(_au, '__imb_standin__', False)
except Exception:
    _au = None
    HAVE_ASTROPY = False

_NDIM = 7


class Unit:
    __array_priority__ = 20000
    __array_ufunc__ = None           # ndarray <op> Unit defers to the reflected methods below

    def __init__(self, scale=1.0, dims=None, name=None, s=None):
        if s is not None or not isinstance(scale, (int, float)):             # astropy's Unit(s) spelling
            src = as_unit(scale if s is None else s)
            scale, dims, name = src.scale, src.dims, src.name
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

    def __rlshift__(self, other):                     # ndarray << unit (fields/field.py:214)
        if isinstance(other, Quantity):
            return other.to(self)
        return Quantity(other, self)

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


_NAMED = {'': dimensionless_unscaled, 'cm': cm, 'm': m, 'pc': pc, 'kpc': kpc, 'G': gauss, 'gauss': gauss,
          'uG': microgauss, 'microgauss': microgauss, 'K': K, 'rad': rad, 'deg': deg, 's': s, 'GHz': GHz,
          'GeV': GeV, 'sr': sr}


def _parse_unit(text):
    """'kpc', 'rad / m2', 'pc cm-3', 'uG**2' ...: products of named units with integer powers."""
    out, sign = dimensionless_unscaled, 1
    for tok in text.replace('*', ' ').replace('/', ' / ').split():
        if tok == '/':
            sign = -1
            continue
        tok = tok.replace('^', '').strip('()')
        k = len(tok)
        while k > 0 and (tok[k - 1].isdigit() or tok[k - 1] in '+-'):
            k -= 1
        name, power = tok[:k], int(tok[k:]) if tok[k:] else 1
        if name not in _NAMED:
            raise ValueError('unknown unit %r in %r' % (name, text))
        out = out * _NAMED[name] ** (sign * power)
    return out


def dimensionless_angles():
    return 'dimensionless_angles'


def as_unit(u):
    if u is None:
        return dimensionless_unscaled
    if isinstance(u, Unit):
        return u
    if isinstance(u, str):
        return _parse_unit(u)
    if isinstance(u, Quantity):
        return Unit(u.unit.scale * float(u.value), u.unit.dims)
    raise TypeError('not a unit: %r' % (u,))


_SAME_UNIT = {np.add, np.subtract, np.maximum, np.minimum, np.fmax, np.fmin, np.hypot, np.fmod, np.remainder}
_COMPARE = {np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal}
_KEEP_UNIT = {np.negative, np.positive, np.absolute, np.fabs, np.rint, np.floor, np.ceil, np.trunc, np.conjugate}
_PLAIN_OUT = {np.isfinite, np.isnan, np.isinf, np.sign, np.signbit}
_DIMLESS = {np.exp, np.exp2, np.expm1, np.log, np.log2, np.log10, np.log1p, np.sinh, np.cosh, np.tanh,
            np.arcsinh, np.arccosh, np.arctanh}
_TRIG = {np.sin, np.cos, np.tan}
_INV_TRIG = {np.arcsin, np.arccos, np.arctan}
_ZERO_DIMS = (0,) * _NDIM


class Quantity(np.ndarray):
    """
    value x Unit as an ndarray subclass, like `astropy.units.Quantity` (the reference's containers test
    `isinstance(data, np.ndarray)` and then `isinstance(data, u.Quantity)`, observable_dict.py:280-288,
    observable.py:62-69).  Units propagate through the ufuncs the likelihood-evaluation path uses.
    """
    __array_priority__ = 10000

    def __new__(cls, value, unit=None, dtype=None, copy=True):
        if isinstance(value, Quantity):
            unit_in = value.unit
            value = value.value if unit is None else value.to_value(unit)
            unit = unit_in if unit is None else unit
        elif isinstance(value, (list, tuple)) and any(isinstance(v, Quantity) for v in value):
            unit = unit or value[0].unit
            value = [Quantity(v).to_value(unit) for v in value]
        arr = np.array(value, dtype=np.float64 if dtype is None else dtype, copy=True if copy else None)
        obj = arr.view(cls)
        obj._unit = as_unit(unit)
        return obj

    def __array_finalize__(self, obj):
        self._unit = getattr(obj, '_unit', dimensionless_unscaled)

    def __reduce__(self):
        fn, args, state = super().__reduce__()
        return fn, args, (state, self._unit)

    def __setstate__(self, state):
        super().__setstate__(state[0])
        self._unit = state[1]

    # conversions ----------------------------------------------------------------------------
    @property
    def unit(self):
        return self._unit

    @property
    def value(self):
        v = self.view(np.ndarray)
        return v.item() if v.ndim == 0 else v

    def to_value(self, unit=None, equivalencies=None):
        if unit is None:
            return self.value
        unit = as_unit(unit)
        if equivalencies == 'dimensionless_angles':
            strip = lambda u: Unit(u.scale, tuple(0 if i == 3 else v for i, v in enumerate(u.dims)))
            f = strip(self._unit).to(strip(unit))
        else:
            f = self._unit.to(unit)
        return self.value if f == 1.0 else self.value * f

    def to(self, unit, equivalencies=None):
        return Quantity(self.to_value(unit, equivalencies), unit)

    def __lshift__(self, unit):
        return self.to(unit)

    def __float__(self):
        return float(self.to_value(dimensionless_unscaled)) if self._unit.dims == _ZERO_DIMS else float(self.value)

    # container behaviour ------------------------------------------------------------------
    def __getitem__(self, idx):
        out = super().__getitem__(idx)
        return out if isinstance(out, Quantity) else Quantity(out, self._unit)

    def __setitem__(self, idx, val):
        super().__setitem__(idx, val.to_value(self._unit) if isinstance(val, Quantity) else val)

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]

    # arithmetic with bare units; everything else goes through the ufuncs ---------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self._unit * other)
        return super().__mul__(other)

    def __rmul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, other * self._unit)
        return super().__rmul__(other)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self._unit / other)
        return super().__truediv__(other)

    def __rtruediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(1.0 / self.value, other / self._unit)
        return super().__rtruediv__(other)

    def __eq__(self, other):
        if isinstance(other, Quantity) and other._unit.dims != self._unit.dims:
            return False
        return super().__eq__(other)

    __hash__ = None

    def _plain(self, other):
        """`other` as plain values in my unit (a bare number only combines with a dimensionless me,
        or when it is zero / non-finite)."""
        if isinstance(other, Quantity):
            return other.to_value(self._unit)
        o = np.asarray(other)
        if self._unit.dims != _ZERO_DIMS and np.any((o != 0) & np.isfinite(o)):
            raise UnitConversionError('cannot combine a dimensional Quantity with a plain number')
        if self._unit.dims == _ZERO_DIMS and self._unit.scale != 1.0:
            return o / self._unit.scale
        return o

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        vals = [i.view(np.ndarray) if isinstance(i, Quantity) else i for i in inputs]
        units = [i._unit if isinstance(i, Quantity) else None for i in inputs]
        if out is not None:
            kwargs['out'] = tuple(o.view(np.ndarray) if isinstance(o, Quantity) else o for o in out)
        first = next(i for i in inputs if isinstance(i, Quantity))

        def run(*v):
            return getattr(ufunc, method)(*v, **kwargs)

        def wrap(res, unit):
            if out is not None:
                for o in out:
                    if isinstance(o, Quantity):
                        o._unit = as_unit(unit)
                return out[0] if len(out) == 1 else out
            return Quantity(res, unit)

        if method in ('reduce', 'accumulate'):
            if ufunc in (np.add, np.maximum, np.minimum, np.fmax, np.fmin):
                return wrap(run(*vals), units[0])
            if ufunc in (np.logical_or, np.logical_and):
                return run(*vals)
            raise TypeError('%s.%s is not defined for Quantities' % (ufunc.__name__, method))
        if method != '__call__':
            raise TypeError('%s.%s is not defined for Quantities' % (ufunc.__name__, method))
        if ufunc is np.multiply:
            u0 = units[0] or dimensionless_unscaled
            return wrap(run(*vals), u0 * (units[1] or dimensionless_unscaled))
        if ufunc in (np.true_divide, np.divide):
            u0 = units[0] or dimensionless_unscaled
            return wrap(run(*vals), u0 / (units[1] or dimensionless_unscaled))
        if ufunc in _SAME_UNIT or ufunc in _COMPARE or ufunc is np.arctan2:
            ref = first
            v = [ref.view(np.ndarray) if i is ref else ref._plain(i) for i in inputs]
            res = run(*v)
            if ufunc in _COMPARE:
                return res
            return wrap(res, rad if ufunc is np.arctan2 else ref._unit)
        if ufunc is np.power:
            p = inputs[1].to_value(dimensionless_unscaled) if isinstance(inputs[1], Quantity) else inputs[1]
            if units[0] is None or units[0].dims == _ZERO_DIMS:               # number ** dimensionless
                base = inputs[0].to_value(dimensionless_unscaled) if isinstance(inputs[0], Quantity) else inputs[0]
                return wrap(run(base, p), dimensionless_unscaled)
            if np.ndim(p) != 0:
                raise TypeError('a dimensional Quantity can only be raised to a scalar power')
            return wrap(run(vals[0], p), units[0] ** float(p))
        if ufunc is np.square:
            return wrap(run(*vals), units[0] ** 2)
        if ufunc is np.sqrt:
            return wrap(run(*vals), units[0] ** 0.5)
        if ufunc is np.cbrt:
            return wrap(run(*vals), units[0] ** (1.0 / 3.0))
        if ufunc is np.reciprocal:
            return wrap(run(*vals), units[0] ** -1)
        if ufunc in _KEEP_UNIT:
            return wrap(run(*vals), units[0])
        if ufunc in _PLAIN_OUT:
            return run(*vals)
        if ufunc in _DIMLESS:
            return wrap(run(first.to_value(dimensionless_unscaled)), dimensionless_unscaled)
        if ufunc in _TRIG:
            return wrap(run(first.to_value(rad, 'dimensionless_angles')), dimensionless_unscaled)
        if ufunc in _INV_TRIG:
            return wrap(run(first.to_value(dimensionless_unscaled)), rad)
        if all(un is None or un.dims == _ZERO_DIMS for un in units):
            v = [i.to_value(dimensionless_unscaled) if isinstance(i, Quantity) else i for i in inputs]
            return run(*v)
        raise TypeError('ufunc %s is not defined for Quantities by this stand-in' % ufunc.__name__)

    def __repr__(self):
        return '<Quantity %r %r>' % (self.value, self._unit)

    __str__ = __repr__


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
    if hasattr(q, 'to_value'):
        return q.to_value(unit)
    return q
