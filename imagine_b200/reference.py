"""
Makes the reference package ``imagine`` importable, so that the GPU drop-ins of ``imagine_b200.api``
subclass the reference's OWN ``Field`` / ``Simulator`` / ``Likelihood`` / ``Observable`` classes.

Resolution order of :func:`ensure_imagine`:

1. ``import imagine`` works as is (a user environment with astropy, mpi4py, ... installed): use it.
2. Otherwise the unmodified reference sources are looked up at ``$IMAGINE_REFERENCE_ROOT``,
   ``<repo>/baseline/_ref`` (the offline ``pip install --no-deps --target`` made by
   ``__graft_entry__.build()``; git-ignored, travels to the GPU box) and ``/root/reference``.  The
   third-party modules the reference imports but this image lacks are replaced by stand-ins that cover
   exactly what the likelihood-evaluation path touches:

   ============================  =========================================================
   ``astropy.units``             ``imagine_b200.units`` (same spelling, small subset)
   ``mpi4py.MPI``                a single-rank ``COMM_WORLD`` (the reference's MPI mode is
                                 replaced by the NCCL exchange of ``hotpath.py``)
   ``e13tools``                  ``add_to_all`` decorator, ``q2tex``
   plotting / sampler / I/O      empty modules (``matplotlib``, ``healpy``, ``hickle``,
   back ends never called here   ``pymultinest``, ``ultranest``, ``dynesty``, ``emcee``,
                                 ``hampyx``, ``corner``, ``cmasher``, ``schwimmbad``, ...)
   ``numpy.int``                 alias of ``int`` (``fields/grid.py:59`` predates NumPy 1.24)
   ``ensemble_seed_generator``   ``tools/random_seed.py:60`` evaluates ``np.uint32(-1)//3``, which
                                 NumPy >= 2 refuses; replaced by the same draw with the literal bound
                                 1431655765 (identical seeds under NumPy 1)
   ============================  =========================================================

   Only ABSENT modules are replaced; anything installed is left alone.

No arithmetic lives here and nothing of the reference is copied: its modules are imported from where
they lie.
"""
import importlib
import importlib.util
import os
import sys
import types
import warnings

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STATE = {'backend': None, 'root': None}

# absent third-party modules that may be replaced by an empty module (never called on this path)
_INERT = ('matplotlib', 'matplotlib.pyplot', 'cmasher', 'healpy', 'hickle', 'corner', 'schwimmbad',
          'pymultinest', 'ultranest', 'dynesty', 'emcee', 'hampyx', 'IPython', 'IPython.display',
          'astropy.table')


def candidate_roots():
    roots = [os.environ.get('IMAGINE_REFERENCE_ROOT'), os.path.join(_REPO, 'baseline', '_ref'),
             '/root/reference']
    return [r for r in roots if r and os.path.isfile(os.path.join(r, 'imagine', '__init__.py'))]


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__imb_standin__ = True
    sys.modules[name] = mod
    if '.' in name:
        parent, child = name.rsplit('.', 1)
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
    return mod


class SingleRankComm:
    """``mpi4py.MPI.COMM_WORLD`` for exactly one rank."""

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def Barrier(self):
        return None

    def Bcast(self, buf, root=0):
        return None

    def bcast(self, obj, root=0):
        return obj

    def allgather(self, obj):
        return [obj]

    def Allreduce(self, send, recv, op=None):
        recv[...] = send

    def allreduce(self, obj, op=None):
        return obj


def _install_standins():
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int                                   # fields/grid.py:59
    if _missing('e13tools'):
        def add_to_all(obj):
            sys.modules[obj.__module__].__all__.append(obj.__name__)
            return obj

        def q2tex(q, *a, **k):
            return str(q)
        _module('e13tools', add_to_all=add_to_all, q2tex=q2tex)
    if _missing('mpi4py'):
        mpi = types.ModuleType('mpi4py.MPI')
        mpi.COMM_WORLD = SingleRankComm()
        mpi.SUM, mpi.MAX, mpi.MIN = 'sum', 'max', 'min'
        mpi.__imb_standin__ = True
        sys.modules['mpi4py.MPI'] = mpi
        _module('mpi4py', MPI=mpi)
    if _missing('astropy'):
        from . import units
        astropy = _module('astropy')
        units_mod = sys.modules[units.__name__]
        units_mod.__imb_standin__ = True
        sys.modules['astropy.units'] = units_mod
        astropy.units = units_mod
    for name in _INERT:
        if name.split('.')[0] == 'astropy' and not getattr(sys.modules.get('astropy'), '__imb_standin__', False):
            continue
        if _missing(name):
            attrs = {}
            if name == 'corner':
                attrs['corner'] = None
            if name == 'hampyx':
                attrs['Hampyx'] = type('Hampyx', (), {})
            if name == 'astropy.table':
                attrs['QTable'] = type('QTable', (), {})
            _module(name, **attrs)


def ensure_imagine():
    """Import ``imagine`` (see module docstring); returns the package.  Raises ImportError if neither
    the package nor its sources can be found."""
    if _STATE['backend'] is not None:
        return sys.modules['imagine']
    try:
        import imagine                                  # noqa: F401
        _patch_numpy2(imagine)
        _STATE['backend'], _STATE['root'] = 'installed', os.path.dirname(os.path.dirname(imagine.__file__))
        return imagine
    except Exception:
        for k in [k for k in sys.modules if k == 'imagine' or k.startswith('imagine.')]:
            del sys.modules[k]
    roots = candidate_roots()
    if not roots:
        raise ImportError(
            'the reference package `imagine` is not importable and its sources were not found at '
            '$IMAGINE_REFERENCE_ROOT, %s or /root/reference; run `python -c "import __graft_entry__ as g; '
            'g.build()"` where /root/reference exists, or install imagine' % os.path.join(_REPO, 'baseline', '_ref'))
    root = roots[0]
    _install_standins()
    sys.path.insert(0, root)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            imagine = importlib.import_module('imagine')
    finally:
        sys.path.remove(root)
    _patch_numpy2(imagine)
    _STATE['backend'], _STATE['root'] = 'sources+standins', root
    return imagine


def _patch_numpy2(imagine):
    """tools/random_seed.py:60 under NumPy >= 2: same draw, literal bound (= np.uint32(-1) // 3)."""
    import numpy as np
    try:
        np.uint32(-1)
        return
    except OverflowError:
        pass

    def ensemble_seed_generator(size):
        return np.random.randint(low=1, high=1431655765, size=int(size))
    ensemble_seed_generator.__doc__ = 'imagine.tools.random_seed.ensemble_seed_generator with the literal bound'
    for name in ('imagine.tools.random_seed', 'imagine.tools', 'imagine.pipelines.pipeline'):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, 'ensemble_seed_generator'):
            mod.ensemble_seed_generator = ensemble_seed_generator


def backend():
    """'installed' | 'sources+standins' (after ensure_imagine)."""
    return _STATE['backend']


def root():
    return _STATE['root']
