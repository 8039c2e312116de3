"""
GPU: the reference's OWN test files, run with the CUDA drop-ins bound to the names they import.

`__graft_entry__.build()` places imagine/tests/*.py next to the installed reference in baseline/_ref
(git-ignored, travels to the GPU box; nothing of it is committed).  Each file is loaded as a module --
its `from imagine.likelihoods import ...` binds the reference classes -- and then the names are re-bound
to the GPU classes / C-ABI wrappers, so the reference's test bodies (mock data from `np.random.rand`,
hand-computed baselines, `np.allclose`, and the exact `==` between EnsembleLikelihood and SimpleLikelihood)
decide whether the drop-ins are drop-ins.
"""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_tests_dir():
    from imagine_b200 import reference
    reference.ensure_imagine()
    path = os.path.join(reference.root(), 'imagine', 'tests')
    if not os.path.isfile(os.path.join(path, 'test_likelihood.py')):
        pytest.skip('the reference test files are not next to the reference package (%s); run '
                    '__graft_entry__.build() where /root/reference exists' % path)
    return path


def _load(name, **rebinding):
    path = os.path.join(_reference_tests_dir(), name + '.py')
    spec = importlib.util.spec_from_file_location('reference_' + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for k, v in rebinding.items():
        assert hasattr(mod, k), 'the reference test module does not import %s' % k
        setattr(mod, k, v)
    return mod


@pytest.fixture(scope='module')
def api():
    import torch
    assert torch.cuda.is_available()
    import imagine_b200.api as api
    from imagine_b200 import _lib
    _lib.init(0)
    return api


LIKELIHOOD_TESTS = [('TestSimpleLikeli', 'test_without_cov'), ('TestSimpleLikeli', 'test_with_cov'),
                    ('TestEnsembleLikeli', 'test'), ('TestEnsembleLikeli', 'test_with_trace_approximation'),
                    ('TestEnsembleLikeli', 'test_diag')]


@pytest.mark.parametrize('cls,method', LIKELIHOOD_TESTS)
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_reference_test_likelihood_py(api, cls, method, seed):
    """imagine/tests/test_likelihood.py with SimpleLikelihood / EnsembleLikelihood / EnsembleLikelihoodDiagonal
    replaced by the B200 classes (non-symmetric np.random.rand 'covariances' included: the LU path)."""
    mod = _load('test_likelihood', SimpleLikelihood=api.B200SimpleLikelihood,
                EnsembleLikelihood=api.B200EnsembleLikelihood,
                EnsembleLikelihoodDiagonal=api.B200EnsembleLikelihoodDiagonal)
    np.random.seed(seed)
    getattr(getattr(mod, cls)(), method)()


def test_every_test_of_the_reference_file_is_listed(api):
    mod = _load('test_likelihood')
    found = [(c, m) for c in dir(mod) if c.startswith('Test') for m in dir(getattr(mod, c)) if m.startswith('test')]
    assert sorted(found) == sorted(LIKELIHOOD_TESTS)


def _tool_wrappers():
    """The reference's array-in / array-out tool functions on top of the C ABI."""
    import torch
    from imagine_b200 import engine
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    host = lambda t: t.cpu().numpy()

    def keep(mask):
        return torch.from_numpy(np.flatnonzero(mask[0].astype(bool)).astype(np.int32)).cuda()
    return dict(
        mpi_mean=lambda arr: host(engine.ensemble_mean(dev(arr))),
        mask_obs=lambda obs, mask: host(engine.gather_columns(dev(obs), keep(mask))),
        mask_cov=lambda cov, mask: host(engine.gather_cov(dev(cov), keep(mask))),
        empirical_cov=lambda arr: host(engine.mcov(dev(arr), 'empirical')[1]),
        oas_cov=lambda arr: host(engine.mcov(dev(arr), 'oas')[1]),
        oas_mcov=lambda arr: tuple(host(t) for t in engine.mcov(dev(arr), 'oas')[:2]))


@pytest.mark.parametrize('method', ['test_mean', 'test_mask', 'test_empirical_cov', 'test_oas_cov', 'test_oas_mcov'])
def test_reference_test_tools_py(api, method):
    """imagine/tests/test_tools.py: pmean, the maskers and the covariance estimators through imb_ensemble_mean,
    imb_gather_columns / imb_gather_cov and imb_mcov."""
    mod = _load('test_tools', **_tool_wrappers())
    np.random.seed(3)
    getattr(mod.TestTools(), method)()
