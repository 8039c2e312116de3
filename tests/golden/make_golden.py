"""
Generates tests/golden/likelihood_golden.npz by running the REFERENCE's own, unmodified
likelihood / covariance / masking code (loaded from /root/reference through
oracle/ref_shim.py) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The inputs are stored next to the outputs so the fixtures are self-contained on the GPU
box (where /root/reference does not exist).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle import ref_shim  # noqa: E402

R = ref_shim.load()
HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference(kind, sims, meas, cov, **kw):
    m = sims.shape[1]
    key = ('test', None, m, None)
    M, C, S = R.Measurements(), R.Covariances(), R.Simulations()
    M.append(name=key, data=meas.reshape(1, -1), otype='plain')
    S.append(name=key, data=sims, otype='plain')
    if cov is not None:
        C.append(name=key, cov_data=cov, otype='plain')
    mask = kw.pop('mask', None)
    masks = None
    if mask is not None:
        masks = R.Masks()
        masks.append(name=key, data=mask.reshape(1, -1), otype='plain')
    covs = C if cov is not None else None
    if kind == 'ensemble':
        est = kw.pop('estimator', 'oas')
        cov_func = {'oas': None, 'empirical': R.empirical_mcov, 'diagonal': R.diagonal_mcov}[est]
        lk = R.EnsembleLikelihood(M, covs, mask_dict=masks, cov_func=cov_func, **kw)
    elif kind == 'ensemble_diagonal':
        lk = R.EnsembleLikelihoodDiagonal(M, covs, mask_dict=masks)
    elif kind == 'simple':
        lk = R.SimpleLikelihood(M, covs, mask_dict=masks)
    else:
        raise ValueError(kind)
    return float(np.asarray(lk(S)).reshape(-1)[0])


def spd(rng, m, scale=1.0):
    a = rng.normal(size=(m, m))
    return scale * (a @ a.T / m + 0.5 * np.eye(m))


def main():
    rng = np.random.default_rng(20261019)
    cases = {}

    def add(name, kind, sims, meas, cov, **kw):
        val = run_reference(kind, sims, meas, cov, **dict(kw))
        cases[name] = dict(kind=kind, sims=sims, meas=meas, cov=cov, value=val, **kw)
        print('%-34s %-18s N=%-3d m=%-5d -> %.17g' % (name, kind, sims.shape[0], sims.shape[1], val))

    # --- small, every branch --------------------------------------------------------
    N, m = 6, 48
    sims = rng.normal(1.0, 0.5, size=(N, m))
    meas = rng.normal(1.0, 0.5, size=m)
    var = rng.uniform(0.05, 0.3, size=m)
    dense = spd(rng, m, 0.2)
    nonspd = rng.random((m, m))
    add('ens_oas_var', 'ensemble', sims, meas, var)
    add('ens_oas_dense_spd', 'ensemble', sims, meas, dense)
    add('ens_oas_dense_nonspd', 'ensemble', sims, meas, nonspd)
    add('ens_emp_var', 'ensemble', sims, meas, var, estimator='empirical')
    add('ens_diagfunc_var', 'ensemble', sims, meas, var, estimator='diagonal')
    add('ens_oas_trace', 'ensemble', sims, meas, np.diag(var), use_trace_approximation=True)
    add('ensdiag_var', 'ensemble_diagonal', sims, meas, var)
    add('ensdiag_dense', 'ensemble_diagonal', sims, meas, dense)
    add('simple_nocov', 'simple', sims, meas, None)
    add('simple_var', 'simple', sims, meas, var)
    add('simple_dense_spd', 'simple', sims, meas, dense)
    add('simple_dense_nonspd', 'simple', sims, meas, nonspd)
    mask = (rng.random(m) > 0.3).astype(np.float64)
    add('ens_oas_var_masked', 'ensemble', sims, meas, var, mask=mask)
    add('ens_oas_dense_masked', 'ensemble', sims, meas, dense, mask=mask)
    add('simple_var_masked', 'simple', sims, meas, var, mask=mask)

    # --- degenerate ensembles (tests/test_likelihood.py:81-149, test_tools.py:154-184) -
    same = np.repeat(rng.random((1, 4)), 2, axis=0)
    c4 = rng.random((4, 4))
    m4 = rng.random(4)
    add('ens_identical_nonspd4', 'ensemble', same, m4, c4)
    add('simple_identical_nonspd4', 'simple', same, m4, c4)
    add('ens_single_row', 'ensemble', sims[:1], meas, var)
    d4 = np.diag(rng.random(4))
    add('ens_identical_trace4', 'ensemble', same, m4, d4, use_trace_approximation=True)
    add('simple_identical_diag4', 'simple', same, m4, d4)

    # --- odd sizes / N > m / m not a multiple of anything -----------------------------
    s2 = rng.normal(size=(37, 13))
    add('ens_oas_N_gt_m', 'ensemble', s2, rng.normal(size=13), rng.uniform(0.5, 1, 13))
    add('ens_oas_N_gt_m_dense', 'ensemble', s2, rng.normal(size=13), spd(rng, 13))
    s3 = rng.normal(size=(5, 131))
    add('ens_oas_m131', 'ensemble', s3, rng.normal(size=131), rng.uniform(0.5, 1, 131))
    add('ens_oas_m131_dense', 'ensemble', s3, rng.normal(size=131), spd(rng, 131))

    # --- HEALPix-sized: Nside 8 (m = 768), N = 16, RM-like scales ----------------------
    N, m = 16, 768
    base = 30.0 * rng.normal(size=m)
    sims = base + 8.0 * rng.normal(size=(N, m))
    meas = base + 8.0 * rng.normal(size=m) + rng.normal(0, 0.3, size=m)
    var = np.full(m, 0.09)
    add('ens_oas_var_nside8', 'ensemble', sims, meas, var)
    # dense measurement covariance at Nside 4 (m = 192) keeps the fixture small
    add('ens_oas_dense_nside4', 'ensemble', sims[:, :192], meas[:192], spd(rng, 192, 0.5))
    add('ensdiag_var_nside8', 'ensemble_diagonal', sims, meas, var)
    add('simple_var_nside8', 'simple', sims, meas, var)

    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            if v is None:
                continue
            flat[name + '/' + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'likelihood_golden.npz'), **flat)

    # estimator-level goldens (mean/cov of oas_mcov / empirical_mcov)
    x = rng.normal(size=(7, 20))
    mean, cov = R.oas_mcov(x)
    mean_e, cov_e = R.empirical_mcov(x)
    np.savez_compressed(os.path.join(HERE, 'estimator_golden.npz'), x=x, oas_mean=mean, oas_cov=cov,
                        emp_mean=mean_e, emp_cov=cov_e)
    print('written', len(cases), 'cases')


if __name__ == '__main__':
    main()
