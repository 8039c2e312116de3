"""Resolves the base classes once (real `imagine` if importable, else the mirror)."""
try:                                            # pragma: no cover
    from imagine.fields import Field, MagneticField, ThermalElectronDensityField, UniformGrid
    from imagine.observables import Observable, Measurements, Simulations, Covariances, Masks
    from imagine.simulators import Simulator
    from imagine.likelihoods import Likelihood
    from .mirror import CosmicRayElectronDensityField
    BACKEND = 'imagine'
except Exception:
    from .mirror import (Field, MagneticField, ThermalElectronDensityField,          # noqa: F401
                         CosmicRayElectronDensityField, UniformGrid, Observable, Measurements,
                         Simulations, Covariances, Masks, Simulator, Likelihood)
    BACKEND = 'mirror'
