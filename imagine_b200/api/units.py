"""Units facade of the plugin layer: the names of `imagine_b200.units` (astropy when installed)."""
from ..units import *                    # noqa: F401,F403
from ..units import (Quantity, Unit, UnitConversionError, HAVE_ASTROPY, as_unit,       # noqa: F401
                     dimensionless_angles, dimensionless_unscaled, unit_scale, value_in)
