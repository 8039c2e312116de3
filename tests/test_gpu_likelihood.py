"""GPU parity, stage 3: ensemble statistics and log-likelihoods vs the reference goldens / oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import likelihood as OK                      # noqa: E402
from tests.helpers import GOLDEN_DIR, load_golden_cases   # noqa: E402

CASES = load_golden_cases()


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    from imagine_b200 import engine, _lib
    _lib.init(0)
    return engine


def _dev(a):
    import torch
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _gpu_value(eng, case, route='auto'):
    import torch
    sims, meas, cov = _dev(case['sims']), _dev(case['meas']), _dev(case['cov'])
    if case['mask'] is not None:
        idx = torch.from_numpy(np.flatnonzero(case['mask']).astype(np.int32)).cuda()
        sims, meas = eng.gather_columns(sims, idx), eng.gather_columns(meas, idx)
        if cov is not None:
            cov = eng.gather_columns(cov, idx) if cov.dim() == 1 else eng.gather_cov(cov, idx)
    if case['kind'] == 'ensemble':
        return eng.loglike_ensemble(sims, meas, cov, estimator=str(case['estimator']), route=route,
                                    use_trace_approximation=bool(case['use_trace_approximation']),
                                    return_info=True)
    if case['kind'] == 'ensemble_diagonal':
        var = cov if cov.dim() == 1 else cov.diagonal().contiguous()
        return eng.loglike_ensemble_diagonal(sims, meas, var), (0, 0)
    return eng.loglike_simple(sims, meas, cov, return_info=True)


@pytest.mark.parametrize('name', sorted(CASES))
def test_golden_cases(eng, name):
    """north_star: log-likelihoods within 1e-8 absolute of the reference path."""
    got, info = _gpu_value(eng, CASES[name])
    assert abs(got - CASES[name]['value']) <= 1e-8, (got, CASES[name]['value'], info)


@pytest.mark.parametrize('name', [n for n, c in CASES.items()
                                  if c['kind'] == 'ensemble' and c['estimator'] != 'diagonal'
                                  and not c['use_trace_approximation']])
def test_dense_route_on_every_ensemble_case(eng, name):
    got, info = _gpu_value(eng, CASES[name], route='dense')
    assert info[0] & 2
    assert abs(got - CASES[name]['value']) <= 1e-8, (got, CASES[name]['value'], info)


def test_route_selection_and_factorisation_flags(eng):
    _, info = _gpu_value(eng, CASES['ens_oas_var'])
    assert info[0] & 1                                   # gram
    _, info = _gpu_value(eng, CASES['ens_oas_dense_spd'])
    assert (info[0] & 2) and (info[0] & 4)               # dense + Cholesky
    _, info = _gpu_value(eng, CASES['ens_oas_dense_nonspd'])
    assert (info[0] & 2) and (info[0] & 8)               # dense + LU (non-symmetric input)


def test_ensemble_equals_simple_exactly_for_identical_rows(eng):
    """tests/test_likelihood.py:81-149 asserts `==`, with and without the trace approximation."""
    e, _ = _gpu_value(eng, CASES['ens_identical_nonspd4'])
    s, _ = _gpu_value(eng, CASES['simple_identical_nonspd4'])
    assert e == s
    e, _ = _gpu_value(eng, CASES['ens_identical_trace4'])
    s, _ = _gpu_value(eng, CASES['simple_identical_diag4'])
    assert e == s
    rng = np.random.default_rng(0)           # the reference test itself, many draws
    for m in (4, 8, 48):
        for _ in range(5):
            row = rng.random((1, m))
            case = dict(kind='ensemble', sims=np.repeat(row, 2, 0), meas=rng.random(m), cov=rng.random((m, m)),
                        mask=None, estimator='oas', use_trace_approximation=False)
            e, _ = _gpu_value(eng, case)
            s, _ = _gpu_value(eng, dict(case, kind='simple'))
            assert e == s
            assert abs(e - OK.simple_loglike(case['sims'], case['meas'], case['cov'])) <= 1e-8 * max(1, abs(e))


def test_diagonal_estimator_equals_ensemble_diagonal(eng):
    """tests/test_likelihood.py:151-178."""
    a, _ = _gpu_value(eng, CASES['ens_diagfunc_var'])
    b, _ = _gpu_value(eng, CASES['ensdiag_var'])
    assert np.allclose(a, b)


def test_mean_var_bit_exact(eng):
    rng = np.random.default_rng(2)
    x = rng.normal(size=(13, 1000))
    assert np.array_equal(eng.ensemble_mean(_dev(x)).cpu().numpy(), np.mean(x, axis=0).reshape(1, -1))
    assert np.array_equal(eng.ensemble_var(_dev(x)).cpu().numpy(), x.var(axis=0))


@pytest.mark.parametrize('est', ['oas', 'empirical', 'diagonal'])
def test_mcov_matches_oracle(eng, est):
    z = np.load(GOLDEN_DIR + '/estimator_golden.npz')
    mean, cov, rho = eng.mcov(_dev(z['x']), est)
    omean, ocov = OK.ESTIMATORS[est](z['x'])
    assert np.array_equal(mean.cpu().numpy(), omean)
    assert np.allclose(cov.cpu().numpy(), ocov, rtol=1e-12, atol=1e-15)
    if est == 'oas':
        assert np.allclose(cov.cpu().numpy(), z['oas_cov'], rtol=1e-12, atol=1e-15)   # reference's own output


def test_single_row_ensemble_zero_covariance(eng):
    """tests/test_tools.py:154-184."""
    x = np.random.default_rng(3).random((1, 32))
    for est in ('oas', 'empirical'):
        mean, cov, _ = eng.mcov(_dev(x), est)
        assert np.array_equal(mean.cpu().numpy(), x)
        assert np.all(cov.cpu().numpy() == 0.0)


@pytest.mark.parametrize('N,m', [(16, 3072), (64, 12288), (3, 1000), (200, 517), (64, 49152), (65, 2048), (130, 4096),
                                 (256, 24576), (400, 1024)])
def test_gram_route_at_scale_vs_oracle_gram(eng, N, m):
    """Sizes where the dense oracle is too slow: compare with the O(mN^2) NumPy restatement.  The ensemble sizes
    walk the three Cholesky finishes of the Gram route: N <= 64 in shared memory, 64 < N <= ~370 panel by
    panel (65: one column in the second panel; 256 = the C4 shard), above that in global memory (400)."""
    rng = np.random.default_rng(N * m)
    base = 30 * rng.normal(size=m)
    sims = base + 8 * rng.normal(size=(N, m))
    meas = base + 8 * rng.normal(size=m)
    var = rng.uniform(0.05, 0.2, size=m)
    for est in ('oas', 'empirical'):
        got = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), estimator=est)
        want = OK.ensemble_loglike_gram(sims, meas, var, estimator=est)
        assert abs(got - want) <= 1e-8 * max(1.0, abs(want) * 1e-4), (got, want)


@pytest.mark.parametrize('N,m', [(16, 3072), (64, 12288), (64, 49152)])
def test_gram_route_1e8_absolute_vs_extended_precision(eng, N, m):
    """north_star's tolerance, 1e-8 ABSOLUTE, at the sizes of C2 / C3 where |log L| ~ 10^5-10^6: the comparator
    is the Gram route in extended precision (oracle/likelihood.py:ensemble_loglike_gram_extended), because a
    float64 NumPy evaluation is itself only good to a few 1e-9 there (the looser bound of the test above)."""
    rng = np.random.default_rng(N + m)
    base = 30 * rng.normal(size=m)
    sims = base + 8 * rng.normal(size=(N, m))
    meas = base + 8 * rng.normal(size=m)
    var = rng.uniform(0.05, 0.2, size=m)
    got = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), estimator='oas')
    want = OK.ensemble_loglike_gram_extended(sims, meas, var, estimator='oas')
    assert abs(want) > 1e4
    assert abs(got - want) <= 1e-8, (got, want, got - want)


@pytest.mark.parametrize('N,m', [(16, 768), (8, 1500), (6, 777), (10, 1290), (4, 2049)])
def test_dense_vs_gram_vs_reference_algebra(eng, N, m):
    """Dense DMMA + Cholesky route == Gram route == NumPy dense LU (the reference algebra).  The sizes walk
    the factorisation's code paths: whole 256-wide block columns (768), a ragged last panel (1500: 28
    columns, 1290: 10), an odd leading dimension (777, 2049: no 16-byte copies, 64 x 64 update tiles)."""
    rng = np.random.default_rng(m)
    sims = rng.normal(size=(N, m)) * 3 + rng.normal(size=m)
    meas = rng.normal(size=m)
    var = rng.uniform(0.5, 1.5, size=m)
    want = OK.ensemble_loglike(sims, meas, var)
    g = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), route='gram')
    d, info = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), route='dense', return_info=True)
    assert info[0] & 4
    assert abs(g - want) <= 1e-8 and abs(d - want) <= 1e-8, (g, d, want)
    a = rng.normal(size=(m, m))
    dense = a @ a.T / m + np.eye(m)
    want = OK.ensemble_loglike(sims, meas, dense)
    d = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(dense))
    assert abs(d - want) <= 1e-8, (d, want)
    s = eng.loglike_simple(_dev(sims), _dev(meas), _dev(dense))
    assert abs(s - OK.simple_loglike(sims, meas, dense)) <= 1e-8


def test_lu_fallback_medium_nonsymmetric(eng):
    rng = np.random.default_rng(9)
    m, N = 200, 6
    sims, meas = rng.normal(size=(N, m)), rng.normal(size=m)
    cov = rng.random((m, m)) + 5 * np.eye(m)
    got, info = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(cov), return_info=True)
    assert info[0] & 8
    want = OK.ensemble_loglike(sims, meas, cov)
    assert abs(got - want) <= 1e-8 * max(1, abs(want))


def test_symmetric_indefinite_falls_back_to_lu(eng):
    rng = np.random.default_rng(10)
    m, N = 24, 4
    sims, meas = rng.normal(size=(N, m)), rng.normal(size=m)
    a = rng.normal(size=(m, m))
    cov = a + a.T                                        # symmetric, indefinite
    got, info = eng.loglike_simple(_dev(sims), _dev(meas), _dev(cov), return_info=True)
    assert info[0] & 8 and info[1] > 0
    want = OK.simple_loglike(sims, meas, cov)
    assert abs(got - want) <= 1e-8 * max(1, abs(want))


def test_singular_matrix_raises(eng):
    m = 6
    with pytest.raises(np.linalg.LinAlgError):
        eng.loglike_simple(_dev(np.ones((2, m))), _dev(np.zeros(m)), _dev(np.zeros((m, m))))


def test_shape_errors(eng):
    with pytest.raises(ValueError):
        eng.loglike_ensemble(_dev(np.ones((2, 5))), _dev(np.ones(4)), _dev(np.ones(5)))
    with pytest.raises(ValueError):
        eng.loglike_ensemble(_dev(np.ones((2, 5))), _dev(np.ones(5)), _dev(np.ones(5)), estimator='diagonal',
                             route='gram')


@pytest.mark.parametrize('N,m,K', [(6, 48, 3), (64, 3072, 5), (17, 1001, 2), (130, 400, 2), (200, 517, 1)])
def test_batched_gram_route_matches_per_key_oracle(eng, N, m, K):
    """imb_loglike_ensemble_batch == the per-key values (any N, incl. N > 64 blocks and N > 158 global path)."""
    rng = np.random.default_rng(N + m + K)
    sims = [rng.normal(size=(N, m)) * (k + 1) + rng.normal(size=m) for k in range(K)]
    meas = [rng.normal(size=m) for _ in range(K)]
    var = [rng.uniform(0.3, 1.2, size=m) for _ in range(K)]
    for est in ('oas', 'empirical'):
        got, info = eng.loglike_ensemble_batch([_dev(s) for s in sims], [_dev(y) for y in meas],
                                               [_dev(v) for v in var], estimator=est, return_info=True)
        for k in range(K):
            want = OK.ensemble_loglike_gram(sims[k], meas[k], var[k], estimator=est)
            assert abs(got[k] - want) <= 1e-8 * max(1.0, abs(want) * 1e-4), (k, got[k], want)
            assert info[k][0] & 1
            single = eng.loglike_ensemble(_dev(sims[k]), _dev(meas[k]), _dev(var[k]), estimator=est)
            assert single == got[k]                       # batching does not change a key's value


def test_batched_route_falls_back_per_key_on_nonpositive_variance(eng):
    rng = np.random.default_rng(4)
    N, m = 5, 24
    sims = [rng.normal(size=(N, m)) for _ in range(2)]
    meas = [rng.normal(size=m) for _ in range(2)]
    var = [rng.uniform(0.5, 1.0, size=m), rng.uniform(0.5, 1.0, size=m)]
    var[1][3] = -2.0                                      # negative "variance": reference LU handles it
    got, info = eng.loglike_ensemble_batch([_dev(s) for s in sims], [_dev(y) for y in meas],
                                           [_dev(v) for v in var], return_info=True)
    assert info[0][0] & 1 and info[1][0] & 8
    for k in range(2):
        want = OK.ensemble_loglike(sims[k], meas[k], var[k])
        assert abs(got[k] - want) <= 1e-8 * max(1.0, abs(want))


@pytest.mark.parametrize('N,m,K,P', [(8, 96, 2, 4), (64, 3072, 3, 8), (12, 1000, 1, 5)])
def test_phased_gram_route_over_pixel_shards(eng, N, m, K, P):
    """
    imb_gram_phase1/2/3 on P pixel slices (what P ranks hold after the all-to-all, blocked row
    layout [src][K][n_local][m/P]) with the two SUM reductions done by hand == the batch result.
    """
    import torch
    rng = np.random.default_rng(N * m + P)
    n_local, mp = N // (P if N % P == 0 else 1), m // P
    S = P if N % P == 0 else 1                      # number of source blocks of rows
    sims = [rng.normal(size=(N, m)) * (k + 1) + rng.normal(size=m) for k in range(K)]
    meas = [rng.normal(size=m) for _ in range(K)]
    var = [rng.uniform(0.3, 1.2, size=m) for _ in range(K)]
    for est in ('oas', 'empirical'):
        want = eng.loglike_ensemble_batch([_dev(s) for s in sims], [_dev(y) for y in meas],
                                          [_dev(v) for v in var], estimator=est)
        shards = []
        for p in range(P):                          # what rank p receives: [S, K, n_local, mp]
            buf = np.empty((S, K, n_local, mp))
            for k in range(K):
                buf[:, k] = sims[k][:, p * mp:(p + 1) * mp].reshape(S, n_local, mp)
            shards.append(_dev(buf))
        layout = (n_local, K * n_local * mp, mp)
        meas_l = [[_dev(meas[k][p * mp:(p + 1) * mp]) for k in range(K)] for p in range(P)]
        var_l = [[_dev(var[k][p * mp:(p + 1) * mp]) for k in range(K)] for p in range(P)]
        G = torch.zeros((K, N, N), dtype=torch.float64, device='cuda')
        Gp = torch.empty_like(G)
        for p in range(P):
            eng.gram_phase1([shards[p][0, k] for k in range(K)], N, mp, layout, meas_l[p], Gp)
            G += Gp                                 # the all-reduce
        P2 = torch.zeros((K, N * N + N + 4), dtype=torch.float64, device='cuda')
        P2p = torch.empty_like(P2)
        for p in range(P):
            ptrs = [shards[p][0, k] for k in range(K)]
            eng.gram_phase1(ptrs, N, mp, layout, meas_l[p], Gp)          # restores rank p's scratch
            Gfull = G.clone()
            eng.gram_phase2(ptrs, N, mp, layout, meas_l[p], var_l[p], m, Gfull, P2p, estimator=est)
            P2 += P2p                               # the all-reduce
        got = eng.gram_phase3(K, N, mp, G.clone(), P2)
        for k in range(K):
            assert abs(got[k] - want[k]) <= 1e-9 * max(1.0, abs(want[k])), (est, k, got[k], want[k])


def test_phased_route_refuses_a_workspace_somebody_else_used(eng):
    """The phases of one sharded evaluation share the library's scratch (means, weights, scalars): a call that
    carves the workspace between phase 1 and phase 2 must turn phase 2 into an error (IMB_ESTATE), not into
    a wrong number."""
    import torch
    from imagine_b200 import _lib
    rng = np.random.default_rng(3)
    N, m, K = 8, 256, 2
    sims = [_dev(rng.normal(size=(N, m))) for _ in range(K)]
    meas = [_dev(rng.normal(size=m)) for _ in range(K)]
    var = [_dev(rng.uniform(0.5, 1.5, size=m)) for _ in range(K)]
    layout = (N, 0, m)
    G = torch.zeros((K, N, N), dtype=torch.float64, device='cuda')
    P2 = torch.zeros((K, N * N + N + 4), dtype=torch.float64, device='cuda')
    eng.gram_phase1(sims, N, m, layout, meas, G)
    eng.gram_phase2(sims, N, m, layout, meas, var, m, G, P2)          # undisturbed: fine
    want = eng.gram_phase3(K, N, m, G, P2)
    eng.gram_phase1(sims, N, m, layout, meas, G)
    eng.loglike_ensemble_diagonal(sims[0], meas[0], var[0])            # somebody else takes the workspace
    with pytest.raises(_lib.ImbError) as err:
        eng.gram_phase2(sims, N, m, layout, meas, var, m, G, P2)
    assert err.value.code == _lib.IMB_ESTATE
    eng.gram_phase1(sims, N, m, layout, meas, G)                       # starting over works
    eng.gram_phase2(sims, N, m, layout, meas, var, m, G, P2)
    assert eng.gram_phase3(K, N, m, G, P2) == want


def test_unit_scaling_law(eng):
    """Expressing maps, measurements and variances in a unit s times smaller shifts the log-likelihood by
    -m log s (the determinant term): what api.VectorizedLikelihood uses to work in the kernels' units."""
    rng = np.random.default_rng(21)
    N, m, s = 12, 300, 1e3
    sims = rng.normal(size=(N, m)) + 2.0
    meas = rng.normal(size=m) + 2.0
    var = rng.uniform(0.5, 1.5, size=m)
    for est in ('oas', 'empirical'):
        base = eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), estimator=est)
        scaled = eng.loglike_ensemble(_dev(sims * s), _dev(meas * s), _dev(var * s * s), estimator=est)
        assert abs(scaled - (base - m * np.log(s))) <= 1e-8 * max(1.0, abs(base))


def test_edge_shapes(eng):
    """Smallest / ragged inputs: one pixel, one realisation, empty masks, N >> m."""
    import torch
    rng = np.random.default_rng(8)
    for N, m in [(1, 1), (2, 1), (5, 1), (1, 7), (70, 3)]:
        sims, meas, var = rng.normal(size=(N, m)), rng.normal(size=m), rng.uniform(0.5, 1.5, m)
        want = OK.ensemble_loglike(sims, meas, var)
        assert abs(eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var)) - want) <= 1e-9 * max(1, abs(want))
        assert abs(eng.loglike_ensemble(_dev(sims), _dev(meas), _dev(var), route='dense') - want) <= 1e-9 * max(1, abs(want))
        assert abs(eng.loglike_ensemble_diagonal(_dev(sims), _dev(meas), _dev(var))
                   - OK.ensemble_diagonal_loglike(sims, meas, var)) <= 1e-9 * max(1, abs(want))
        assert abs(eng.loglike_simple(_dev(sims), _dev(meas), None) - OK.simple_loglike(sims, meas, None)) <= 1e-12 * max(1, abs(want))
    idx = torch.zeros(0, dtype=torch.int32, device='cuda')
    assert tuple(eng.gather_columns(_dev(rng.normal(size=(3, 5))), idx).shape) == (3, 0)
    assert tuple(eng.gather_cov(_dev(np.eye(5)), idx).shape) == (0, 0)
    with pytest.raises(ValueError):
        eng.loglike_ensemble(_dev(np.zeros((2, 0))), _dev(np.zeros(0)), _dev(np.zeros(0)))
