"""
Base classes of the plugin layer: the REFERENCE's own classes (``imagine_b200.reference`` makes the
package importable; stand-ins only for third-party modules that are absent).
"""
from ..reference import ensure_imagine, backend

ensure_imagine()

from imagine.fields import Field, MagneticField, ThermalElectronDensityField, UniformGrid      # noqa: E402,F401
from imagine.observables import Observable, Measurements, Simulations, Covariances, Masks      # noqa: E402,F401
from imagine.simulators import Simulator                                                      # noqa: E402,F401
from imagine.likelihoods import Likelihood                                                    # noqa: E402,F401

BACKEND = backend()          # 'installed' | 'sources+standins'
