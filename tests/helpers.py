"""Shared test helpers (golden loaders, synthetic workloads)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden_cases():
    """tests/golden/likelihood_golden.npz -> {case: {kind, sims, meas, cov|None, value, ...}}."""
    z = np.load(os.path.join(GOLDEN_DIR, 'likelihood_golden.npz'), allow_pickle=False)
    cases = {}
    for k in z.files:
        name, field = k.split('/')
        v = z[k]
        if v.ndim == 0:
            v = v.item()
        cases.setdefault(name, {})[field] = v
    for c in cases.values():
        c.setdefault('cov', None)
        c.setdefault('mask', None)
        c.setdefault('estimator', 'oas')
        c.setdefault('use_trace_approximation', False)
    return cases


def oracle_value(case):
    """Evaluate a golden case with the NumPy oracle."""
    from oracle import likelihood as L
    sims, meas, cov = case['sims'], case['meas'], case['cov']
    if case['mask'] is not None:
        sims = L.mask_obs(sims, case['mask'])
        meas = L.mask_var(meas, case['mask'])
        if cov is not None:
            cov = L.mask_var(cov, case['mask']) if cov.ndim == 1 else L.mask_cov(cov, case['mask'])
    if case['kind'] == 'ensemble':
        return L.ensemble_loglike(sims, meas, cov, estimator=case['estimator'],
                                  use_trace_approximation=bool(case['use_trace_approximation']))
    if case['kind'] == 'ensemble_diagonal':
        return L.ensemble_diagonal_loglike(sims, meas, cov)
    if case['kind'] == 'simple':
        return L.simple_loglike(sims, meas, cov)
    raise ValueError(case['kind'])


def b_to_reference_layout(B, n_real, grid_shape):
    """Device B buffer (b_stride 3 | 4 | interleaved 24 / 48 / 96) -> NumPy [rows, nx, ny, nz, 3]."""
    a = B.cpu().numpy()
    if a.ndim == 4 and a.shape[2] == 3:               # [group][cell][x|y|z][realisation % RL]
        g, ncell, _, rl = a.shape
        a = a.transpose(0, 3, 1, 2).reshape(g * rl, ncell, 3)
        assert (a[n_real:] == 0.0).all()
        a = a[:n_real]
    rows = a.shape[0]
    return a.reshape(rows, *grid_shape, a.shape[-1])[..., :3]
