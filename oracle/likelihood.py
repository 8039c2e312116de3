"""
CPU oracle, stage 3: ensemble covariance estimators and Gaussian log-likelihoods on raw
arrays.  NumPy restatement (test infrastructure only) of

  pmean / pvar / ptrace / pslogdet / plu_solve   tools/parallel_ops.py:45-176
  empirical_mcov      tools/covariance_estimator.py:106-148   (1/N normalisation)
  oas_mcov            tools/covariance_estimator.py:194-237
  diagonal_mcov       tools/covariance_estimator.py:48-65
  EnsembleLikelihood.call          likelihoods/ensemble_likelihood.py:68-107
  EnsembleLikelihoodDiagonal.call  likelihoods/ensemble_likelihood.py:137-173
  SimpleLikelihood.call            likelihoods/simple_likelihood.py:33-84
  mask_obs / mask_var / mask_cov   tools/masker.py:25-116

Pinned against the reference's own code (run unmodified through oracle/ref_shim.py) by
tests/golden/likelihood_golden.npz and, in the build container, live.

Conventions: ``sims`` [N,m] (row = realisation), ``meas`` [m] or [1,m], ``cov`` is either
None, a 1-D variance [m] (the reference's 'variance' dtype, expanded with np.diag exactly
like observables/observable.py:91-94) or a dense [m,m] matrix.
"""
import numpy as np


def _as_dense(cov):
    cov = np.asarray(cov, dtype=np.float64)
    return np.diag(cov) if cov.ndim == 1 else cov


def _var_of(cov):
    cov = np.asarray(cov, dtype=np.float64)
    return cov if cov.ndim == 1 else cov.diagonal()


def pmean(data):
    return (np.mean(data, axis=0)).reshape(1, -1)


def empirical_mcov(data):
    n = data.shape[0]
    mean = pmean(data)
    u = data - mean
    return mean, (u.T @ u) / n


def oas_shrinkage(trs, trs2, n, m):
    """rho of covariance_estimator.py:228-234 (denominator == 0 -> rho = 1)."""
    num = (1.0 - 2.0 / m) * trs2 + trs * trs
    den = (n + 1.0 - 2.0 / m) * (trs2 - (trs * trs) / m)
    if den == 0:
        return 1.0
    return float(np.min([1, num / den]))


def oas_mcov(data):
    n, m = data.shape
    mean = pmean(data)
    u = data - mean
    s = (u.T @ u) / n
    trs = np.trace(s)
    trs2 = np.trace(s @ s)
    rho = oas_shrinkage(trs, trs2, n, m)
    return mean, (1. - rho) * s + np.eye(m) * rho * trs / m


def diagonal_mcov(data):
    return pmean(data), np.diag(data.var(axis=0))


ESTIMATORS = {'oas': oas_mcov, 'empirical': empirical_mcov, 'diagonal': diagonal_mcov}


def ensemble_loglike(sims, meas, cov, estimator='oas', use_trace_approximation=False):
    """One key of EnsembleLikelihood.call (dense route, LU like the reference)."""
    meas = np.asarray(meas, dtype=np.float64).reshape(1, -1)
    mean, sim_cov = ESTIMATORS[estimator](sims)
    diff = meas - mean
    full = _as_dense(cov) + sim_cov
    if not use_trace_approximation:
        sign, logdet = np.linalg.slogdet(full * 2. * np.pi)
    else:
        diag_sum = _var_of(cov) + sims.var(axis=0)
        sign, logdet = 1, (np.log(diag_sum * 2. * np.pi)).sum()
    return float(-0.5 * (np.vdot(diff, np.linalg.solve(full, diff.T)) + sign * logdet))


def ensemble_loglike_gram(sims, meas, var, estimator='oas'):
    """
    Same number as `ensemble_loglike` when the measurement covariance is diagonal, from the
    N x N Gram matrix (Woodbury + matrix-determinant lemma).  O(m N^2).
    Restates what the GPU Gram route computes, for debugging; the *pin* is the dense path.
    """
    n, m = sims.shape
    meas = np.asarray(meas, dtype=np.float64).reshape(-1)
    mean = np.mean(sims, axis=0)
    u = sims - mean
    g = u @ u.T
    if estimator == 'oas':
        trs = np.trace(g) / n
        trs2 = np.sum(g * g) / (n * n)
        rho = oas_shrinkage(trs, trs2, n, m)
    elif estimator == 'empirical':
        trs, rho = 0.0, 0.0
    else:
        raise ValueError(estimator)
    dvec = np.asarray(var, dtype=np.float64) + rho * trs / m
    d = meas - mean
    c = (1.0 - rho) / n
    ud = u / dvec
    kmat = np.eye(n) + c * (ud @ u.T)
    y = ud @ d
    sign_k, logdet_k = np.linalg.slogdet(kmat)
    sign_d = np.sign(dvec).prod()
    logdet = np.log(np.abs(dvec) * 2. * np.pi).sum() + logdet_k
    quad = np.dot(d / dvec, d) - c * np.dot(y, np.linalg.solve(kmat, y))
    return float(-0.5 * (quad + sign_d * sign_k * logdet))


def ensemble_loglike_gram_extended(sims, meas, var, estimator='oas'):
    """
    The Gram route evaluated in EXTENDED precision (np.longdouble: 64-bit mantissa on x86) wherever a long
    reduction is involved, as a comparator for absolute-tolerance tests at full size: at |log L| ~ 10^6-10^7 a
    double carries only ~2e-9, so a float64 NumPy comparator is itself uncertain at the 1e-8 level.  The
    N x N solve runs in float64 on matrices rounded from extended precision (condition number ~ 1 + O(1)).
    O(m N^2) longdouble operations: seconds at m = 49 152, N = 64.
    """
    ld = np.longdouble
    n, m = sims.shape
    x = np.asarray(sims, dtype=ld)
    meas = np.asarray(meas, dtype=ld).reshape(-1)
    var = np.asarray(var, dtype=ld).reshape(-1)
    mean = x.sum(axis=0) / ld(n)
    u = x - mean
    g = np.zeros((n, n), dtype=ld)
    for i in range(n):
        g[i, :i + 1] = (u[:i + 1] * u[i]).sum(axis=1)
        g[:i + 1, i] = g[i, :i + 1]
    if estimator == 'oas':
        trs = np.trace(g) / ld(n)
        trs2 = (g * g).sum() / ld(n * n)
        num = (ld(1) - ld(2) / ld(m)) * trs2 + trs * trs
        den = (ld(n) + ld(1) - ld(2) / ld(m)) * (trs2 - trs * trs / ld(m))
        rho = ld(1) if den == 0 else min(ld(1), num / den)
    elif estimator == 'empirical':
        trs, rho = ld(0), ld(0)
    else:
        raise ValueError(estimator)
    dvec = var + rho * trs / ld(m)
    d = meas - mean
    c = (ld(1) - rho) / ld(n)
    ud = u / dvec
    gw = np.zeros((n, n), dtype=ld)
    for i in range(n):
        gw[i, :i + 1] = (ud[:i + 1] * u[i]).sum(axis=1)
        gw[:i + 1, i] = gw[i, :i + 1]
    y = (ud * d).sum(axis=1)
    kmat = (np.eye(n, dtype=ld) + c * gw).astype(np.float64)
    sign_k, logdet_k = np.linalg.slogdet(kmat)
    z = np.linalg.solve(kmat, y.astype(np.float64))
    sign_d = np.sign(dvec).prod()
    logdet = np.log(np.abs(dvec) * ld(2) * ld(np.pi)).sum() + ld(logdet_k)
    quad = (d / dvec * d).sum() - c * (y * z.astype(ld)).sum()
    return float(ld(-0.5) * (quad + sign_d * ld(sign_k) * logdet))


def ensemble_diagonal_loglike(sims, meas, cov):
    meas = np.asarray(meas, dtype=np.float64).reshape(1, -1)
    diff = meas - pmean(sims)
    full_var = _var_of(cov) + sims.var(axis=0)
    sign = np.sign(full_var).prod()
    logdet = np.log(full_var * 2. * np.pi).sum()
    return float(-0.5 * np.vdot(diff, 1. / full_var * diff) - 0.5 * sign * logdet)


def simple_loglike(sims, meas, cov=None):
    meas = np.asarray(meas, dtype=np.float64).reshape(1, -1)
    diff = meas - pmean(sims)
    if cov is None:
        return float(-0.5 * np.vdot(diff, diff))
    cov = np.asarray(cov, dtype=np.float64)
    if cov.ndim == 1:
        sign = np.sign(cov).prod()
        logdet = np.log(cov * 2. * np.pi).sum()
        return float(-0.5 * np.vdot(diff, 1. / cov * diff) - 0.5 * sign * logdet)
    sign, logdet = np.linalg.slogdet(cov * 2 * np.pi)
    return float(-0.5 * np.vdot(diff, np.linalg.solve(cov, diff.T)) - 0.5 * sign * logdet)


def mask_obs(obs, mask):
    return obs[:, np.asarray(mask).reshape(-1).astype(bool)]


def mask_var(var, mask):
    return var[np.asarray(mask).reshape(-1).astype(bool)]


def mask_cov(cov, mask):
    b = np.asarray(mask).reshape(-1).astype(bool)
    return cov[np.ix_(b, b)]
