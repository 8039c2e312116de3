"""Oracle stage 3 vs the reference's own code (golden fixtures; live when the tree is present)."""
import numpy as np
import pytest

from oracle import likelihood as L
from oracle import ref_shim
from tests.helpers import GOLDEN_DIR, load_golden_cases, oracle_value

CASES = load_golden_cases()


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_reference_golden(name):
    case = CASES[name]
    got = oracle_value(case)
    # same LAPACK, same formulae -> last-digit agreement; 1e-8 abs is north_star's bar
    assert abs(got - case['value']) <= 1e-10 * max(1.0, abs(case['value'])), (got, case['value'])


def test_reference_structural_identities():
    """tests/test_likelihood.py:81-149: Ensemble == Simple exactly for identical rows."""
    assert CASES['ens_identical_nonspd4']['value'] == CASES['simple_identical_nonspd4']['value']
    assert CASES['ens_identical_trace4']['value'] == CASES['simple_identical_diag4']['value']
    # tests/test_likelihood.py:151-178
    assert np.allclose(CASES['ens_diagfunc_var']['value'], CASES['ensdiag_var']['value'])


def test_estimator_goldens():
    z = np.load(GOLDEN_DIR + '/estimator_golden.npz')
    mean, cov = L.oas_mcov(z['x'])
    assert np.array_equal(mean, z['oas_mean'])
    assert np.allclose(cov, z['oas_cov'], rtol=1e-14, atol=1e-16)
    mean, cov = L.empirical_mcov(z['x'])
    assert np.allclose(cov, z['emp_cov'], rtol=1e-14, atol=1e-16)


def test_single_row_ensemble_zero_covariance():
    """tests/test_tools.py:154-184: 1-row ensemble -> covariance == 0, mean == row."""
    x = np.random.default_rng(3).random((1, 32))
    for est in (L.empirical_mcov, L.oas_mcov):
        mean, cov = est(x)
        assert np.array_equal(mean, x)
        assert np.all(cov == 0.0)


@pytest.mark.parametrize('name', ['ens_oas_var', 'ens_emp_var', 'ens_oas_var_nside8',
                                  'ens_oas_N_gt_m', 'ens_oas_m131', 'ens_single_row'])
def test_gram_route_equals_dense(name):
    """Woodbury / determinant-lemma route == the reference dense LU route."""
    c = CASES[name]
    got = L.ensemble_loglike_gram(c['sims'], c['meas'], c['cov'], estimator=c['estimator'])
    assert abs(got - c['value']) <= 1e-8


def test_end_to_end_template_golden():
    """
    -55.581767201053125 (tests/mocks_for_templates.py:111-112; setup
    tests/test_templates.py:146-202): grid [30,1,1] x in [0,2pi], B = (10.5, 0.5, 3) uG,
    ne == 1, TestSimulator at x = 0..4, meas == 1, err 0.1 ('variance' = err^2), N = 2.
    """
    from oracle import fields as F
    g = F.Grid([[0, 2 * np.pi], [0, 0], [0, 0]], [30, 1, 1])
    B = F.constant_B(g, 10.5, 0.5, 3.0)
    ne = np.ones(g.shape)
    pts = np.stack([np.arange(5.0), np.zeros(5), np.zeros(5)], axis=1)
    row = F.test_simulator_sample(g, B, ne, pts, b_axis=1)
    sims = np.stack([row, row])
    val = L.simple_loglike(sims, np.ones(5), np.ones(5) * 0.1 ** 2)
    assert val == pytest.approx(-55.581767201053125, abs=1e-12)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present (GPU box)')
def test_oracle_live_against_reference_random():
    R = ref_shim.load()
    rng = np.random.default_rng(11)
    for N, m in [(3, 7), (9, 33), (20, 64)]:
        sims, meas = rng.normal(size=(N, m)), rng.normal(size=m)
        var = rng.uniform(0.2, 2.0, m)
        key = ('x', None, m, None)
        M, C, S = R.Measurements(), R.Covariances(), R.Simulations()
        M.append(name=key, data=meas[None], otype='plain')
        C.append(name=key, cov_data=var, otype='plain')
        S.append(name=key, data=sims, otype='plain')
        ref = R.EnsembleLikelihood(M, C)(S)
        assert abs(L.ensemble_loglike(sims, meas, var) - ref) <= 1e-10 * abs(ref)
        assert abs(L.ensemble_loglike_gram(sims, meas, var) - ref) <= 1e-9
        ref_d = R.EnsembleLikelihoodDiagonal(M, C)(S)
        assert abs(L.ensemble_diagonal_loglike(sims, meas, var) - ref_d) <= 1e-10 * abs(ref_d)
