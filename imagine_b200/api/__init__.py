"""
Reference-facing plugin layer: GPU drop-ins for IMAGINE's Field / Simulator / Likelihood API.

Base classes come from the real `imagine` package when it can be imported; otherwise from
`mirror.py`, which restates the same interface (the reference needs astropy, mpi4py,
e13tools ... that are not in this image).
"""
from ._bases import (Field, MagneticField, ThermalElectronDensityField,              # noqa: F401
                     CosmicRayElectronDensityField, UniformGrid, Observable, Measurements,
                     Simulations, Covariances, Masks, Simulator, Likelihood, BACKEND)

from . import units                                                                    # noqa: F401
from .device_observable import DeviceObservable, ShardedDeviceObservable                                        # noqa: F401
from .fields import (ConstantMagneticField, NaiveGaussianMagneticField, LSAMagneticField,                # noqa: F401
                     ConstantThermalElectrons, ExponentialThermalElectrons,
                     RandomThermalElectrons, CosThermalElectronDensity,
                     AnalyticCosmicRayElectrons, ConstantCosmicRayElectrons)
from .simulators import B200Simulator, B200TestSimulator                               # noqa: F401
from .likelihoods import (B200EnsembleLikelihood, B200EnsembleLikelihoodDiagonal,      # noqa: F401
                          B200SimpleLikelihood)
