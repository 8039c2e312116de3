"""
Runs the reference's OWN likelihood / covariance / field code, unmodified (oracle, test infrastructure
only): used (a) by ``tests/golden/make_golden.py`` to generate the committed golden vectors and (b) by
the tests, to re-validate ``oracle/likelihood.py`` and the CUDA kernels against the real thing.

The import machinery lives in ``imagine_b200.reference`` (the plugin layer needs the same classes):
the package is imported from ``baseline/_ref`` (offline install made by ``__graft_entry__.build()``;
travels to the GPU box) or ``/root/reference``, with stand-ins only for third-party modules this image
lacks.  Nothing of the reference is copied.
"""
import types

_loaded = None


def available():
    from imagine_b200 import reference
    try:
        import imagine  # noqa: F401
        return True
    except Exception:
        return bool(reference.candidate_roots())


def load():
    """Import the reference likelihood stack; returns a namespace object."""
    global _loaded
    if _loaded is not None:
        return _loaded
    from imagine_b200 import reference
    reference.ensure_imagine()
    import imagine.likelihoods as lk
    import imagine.observables as obs
    import imagine.tools as tools
    import imagine.tools.covariance_estimator as cov
    _loaded = types.SimpleNamespace(
        Observable=obs.Observable,
        Measurements=obs.Measurements,
        Simulations=obs.Simulations,
        Covariances=obs.Covariances,
        Masks=obs.Masks,
        EnsembleLikelihood=lk.EnsembleLikelihood,
        EnsembleLikelihoodDiagonal=lk.EnsembleLikelihoodDiagonal,
        SimpleLikelihood=lk.SimpleLikelihood,
        oas_mcov=cov.oas_mcov, oas_cov=cov.oas_cov,
        empirical_mcov=cov.empirical_mcov, empirical_cov=cov.empirical_cov,
        diagonal_mcov=cov.diagonal_mcov,
        mask_obs=tools.mask_obs, mask_var=tools.mask_var, mask_cov=tools.mask_cov,
    )
    return _loaded
