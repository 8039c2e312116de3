"""
imagine_b200 -- B200-native implementation of IMAGINE's likelihood-evaluation hot path
(fields on a UniformGrid -> line-of-sight sky maps -> ensemble Gaussian log-likelihood).

Layout
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/imagine_b200.h)
  _lib.py          ctypes binding of libimagine_b200.so (no CPU fallback)
  engine.py        tensor-level face of the C ABI
  api/             host-side mirror of the reference plugin interface (Field / Simulator /
                   Likelihood / Observable dictionaries) and the GPU drop-ins
"""
__version__ = '0.1.0'

from . import _lib, engine            # noqa: F401
from ._build import build             # noqa: F401
