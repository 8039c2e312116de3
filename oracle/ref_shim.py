"""
Loader that runs the reference's OWN likelihood / covariance code, unmodified, from
``/root/reference`` inside the build container (oracle, test infrastructure only).

The reference package cannot be imported normally here: ``astropy``, ``mpi4py``,
``e13tools``, ``healpy``, ``hickle`` ... are not installed and the ``__init__`` chains
pull all of them in.  This module registers tiny stand-in modules for those imports,
registers *empty* parent packages whose ``__path__`` points at the read-only reference
tree (so sub-modules load from where they lie, never copied), and then imports

    imagine/tools/{config,class_tools,parallel_ops,covariance_estimator,masker}.py
    imagine/observables/{observable,observable_dict}.py
    imagine/likelihoods/{likelihood,ensemble_likelihood,simple_likelihood}.py

It is used (a) by ``tests/golden/make_golden.py`` to generate the committed golden
vectors and (b) by not-gpu tests, when ``/root/reference`` exists, to re-validate
``oracle/likelihood.py`` against the real thing.  ``/root/reference`` does not exist on
the GPU box: nothing that runs there may call :func:`load`.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('IMAGINE_REFERENCE_ROOT', '/root/reference')

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'imagine', 'likelihoods'))


def _fake_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _SingleRankComm:
    """mpi4py.MPI.COMM_WORLD for exactly one rank."""

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


def load():
    """Import the reference likelihood stack; returns a namespace object."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)

    def add_to_all(obj):
        sys.modules[obj.__module__].__all__.append(obj.__name__)
        return obj

    _fake_module('e13tools', add_to_all=add_to_all)
    mpi = _fake_module('mpi4py.MPI', COMM_WORLD=_SingleRankComm())
    _fake_module('mpi4py', MPI=mpi)

    class Quantity:  # only ever used in isinstance() on this path
        pass

    units = _fake_module('astropy.units', Quantity=Quantity)
    _fake_module('astropy', units=units)
    for name in ('matplotlib', 'matplotlib.pyplot', 'cmasher', 'healpy', 'hickle'):
        if name not in sys.modules:
            _fake_module(name)
    _fake_module('corner', corner=None)

    for pkg in ('imagine', 'imagine.tools', 'imagine.observables', 'imagine.likelihoods'):
        mod = types.ModuleType(pkg)
        mod.__path__ = [os.path.join(REFERENCE_ROOT, *pkg.split('.'))]
        sys.modules[pkg] = mod

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        tools = sys.modules['imagine.tools']
        tools.rc = importlib.import_module('imagine.tools.config').rc
        ct = importlib.import_module('imagine.tools.class_tools')
        tools.BaseClass, tools.req_attr = ct.BaseClass, ct.req_attr
        for sub in ('parallel_ops', 'covariance_estimator', 'masker'):
            m = importlib.import_module('imagine.tools.' + sub)
            for n in m.__all__:
                setattr(tools, n, getattr(m, n))
        obs_dict = importlib.import_module('imagine.observables.observable_dict')
        obs = importlib.import_module('imagine.observables.observable')
        lk = importlib.import_module('imagine.likelihoods.likelihood')
        sys.modules['imagine.likelihoods'].Likelihood = lk.Likelihood
        ens = importlib.import_module('imagine.likelihoods.ensemble_likelihood')
        simple = importlib.import_module('imagine.likelihoods.simple_likelihood')
        cov = sys.modules['imagine.tools.covariance_estimator']

    ns = types.SimpleNamespace(
        Observable=obs.Observable,
        Measurements=obs_dict.Measurements,
        Simulations=obs_dict.Simulations,
        Covariances=obs_dict.Covariances,
        Masks=obs_dict.Masks,
        EnsembleLikelihood=ens.EnsembleLikelihood,
        EnsembleLikelihoodDiagonal=ens.EnsembleLikelihoodDiagonal,
        SimpleLikelihood=simple.SimpleLikelihood,
        oas_mcov=cov.oas_mcov, oas_cov=cov.oas_cov,
        empirical_mcov=cov.empirical_mcov, empirical_cov=cov.empirical_cov,
        diagonal_mcov=cov.diagonal_mcov,
        mask_obs=tools.mask_obs, mask_var=tools.mask_var, mask_cov=tools.mask_cov,
    )
    _loaded = ns
    return ns
