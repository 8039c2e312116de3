"""
Reference-facing plugin layer: GPU drop-ins for IMAGINE's Field / Simulator / Likelihood API.

The base classes are the reference's own (`imagine.fields.Field`, `imagine.simulators.Simulator`,
`imagine.likelihoods.Likelihood`, the Observable dictionaries): `imagine_b200.reference` imports the
unmodified package, with stand-ins only for third-party modules this image lacks (astropy.units,
mpi4py, e13tools, plotting / sampler back ends).
"""
from ._bases import (Field, MagneticField, ThermalElectronDensityField, UniformGrid,   # noqa: F401
                     Observable, Measurements, Simulations, Covariances, Masks, Simulator,
                     Likelihood, BACKEND)

from . import units                                                                    # noqa: F401
from .device_observable import DeviceObservable, ShardedDeviceObservable                                        # noqa: F401
from .fields import (CosmicRayElectronDensityField, ConstantMagneticField, NaiveGaussianMagneticField, LSAMagneticField, JF12MagneticField,                # noqa: F401
                     ConstantThermalElectrons, ExponentialThermalElectrons,
                     RandomThermalElectrons, CosThermalElectronDensity,
                     AnalyticCosmicRayElectrons, ConstantCosmicRayElectrons)
from .simulators import B200Simulator, B200TestSimulator                               # noqa: F401
from .vectorized import VectorizedLikelihood                                           # noqa: F401
from .likelihoods import (B200EnsembleLikelihood, B200EnsembleLikelihoodDiagonal,      # noqa: F401
                          B200SimpleLikelihood)
