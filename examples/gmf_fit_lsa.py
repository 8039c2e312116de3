#!/usr/bin/env python
"""
BASELINE.json configs[4] in spirit: a 4-parameter Galactic-magnetic-field fit (WMAP LSA model:
b0, psi0, psi1, chi0) with the GPU Simulator + Likelihood drop-ins behind IMAGINE's plugin API.

The samplers IMAGINE wraps (MultiNest, UltraNest, dynesty, emcee) are not installed in this image,
so the driver below is a plain random-walk Metropolis loop that calls exactly what
`Pipeline._core_likelihood` calls (pipelines/pipeline.py:905-943): build the Field objects for a
parameter point, `simulator(field_list)`, `likelihood(simulations)`.  With a real IMAGINE install
the same `simulator` / `likelihood` objects are passed to `img.pipelines.MultinestPipeline`.

    python examples/gmf_fit_lsa.py [--steps 300] [--grid 64] [--nside 16] [--ensemble 8]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import imagine_b200.api as api           # noqa: E402

u = api.units
TRUE = dict(b0=6.0, psi0=27.0, psi1=0.9, chi0=25.0)          # fields/hamx/breg_lsa.py:46-49 defaults
PRIOR = dict(b0=(0., 10.), psi0=(0., 50.), psi1=(0., 5.), chi0=(-25., 50.))   # breg_lsa.py:50-53


def make_fields(grid, theta, n_ens, seeds):
    return [
        api.LSAMagneticField(grid, parameters={'b0': theta[0] * u.microgauss, 'psi0': theta[1] * u.deg,
                                               'psi1': theta[2] * u.deg, 'chi0': theta[3] * u.deg},
                             ensemble_size=n_ens),
        api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 1.0 * u.microgauss},
                                       ensemble_size=n_ens, ensemble_seeds=seeds),
        api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3,
                                                          'scale_radius': 3 * u.kpc, 'scale_height': 0.5 * u.kpc},
                                        ensemble_size=n_ens),
        api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5 * u.kpc, 'z0': 1 * u.kpc,
                                                         'r_sun': 8.3 * u.kpc}, ensemble_size=n_ens)]


def build(grid_n=64, nside=16, n_ens=8, noise=0.05, seed=1):
    """Mock data from the true parameters (one realisation, independent seed) + noise."""
    grid = api.UniformGrid(box=[[-20, 20] * u.kpc, [-20, 20] * u.kpc, [-4, 4] * u.kpc],
                           resolution=[grid_n] * 3)
    m = 12 * nside ** 2
    keys = [('fd', None, nside, None), ('sync', 23.0, nside, 'I'), ('sync', 23.0, nside, 'Q'),
            ('sync', 23.0, nside, 'U')]
    units = [u.rad / u.m ** 2, u.K, u.K, u.K]
    blank = api.Measurements()
    for k, un in zip(keys, units):
        blank.append(name=k, data=np.zeros((1, m)) * un, otype='HEALPix')
    truth = api.B200Simulator(blank)(make_fields(grid, list(TRUE.values()), 1, [987654321]))
    rng = np.random.default_rng(seed)
    meas = api.Measurements()
    for k, un in zip(keys, units):
        row = truth[k].data[0]
        sigma = noise * np.abs(row).mean()
        meas.append(name=k, data=(row + sigma * rng.normal(size=m)).reshape(1, m) * un,
                    cov_data=np.full(m, sigma ** 2) * un ** 2, otype='HEALPix')
    simulator = api.B200Simulator(meas)
    likelihood = api.B200EnsembleLikelihood(meas)
    np.random.seed(1)
    seeds = [int(s) for s in np.random.randint(1, 1431655765, n_ens)]

    def loglike(theta):
        for v, (lo, hi) in zip(theta, PRIOR.values()):
            if not lo <= v <= hi:
                return -np.inf
        return likelihood(simulator(make_fields(grid, theta, n_ens, seeds)))
    return loglike


def metropolis(loglike, start, steps, scale, rng):
    x, lx = np.array(start, dtype=float), loglike(start)
    chain, accepted = [], 0
    for _ in range(steps):
        y = x + scale * rng.normal(size=x.size)
        ly = loglike(list(y))
        if np.log(rng.random()) < ly - lx:
            x, lx, accepted = y, ly, accepted + 1
        chain.append(x.copy())
    return np.array(chain), accepted / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--grid', type=int, default=64)
    ap.add_argument('--nside', type=int, default=16)
    ap.add_argument('--ensemble', type=int, default=8)
    args = ap.parse_args()
    loglike = build(args.grid, args.nside, args.ensemble)
    rng = np.random.default_rng(7)
    start = [5.0, 20.0, 1.5, 15.0]
    t0 = time.perf_counter()
    chain, acc = metropolis(loglike, start, args.steps, np.array([0.05, 0.3, 0.05, 0.5]), rng)
    dt = time.perf_counter() - t0
    burn = chain[len(chain) // 2:]
    print('%d likelihood evaluations in %.2f s (%.1f evals/s), acceptance %.2f'
          % (args.steps + 1, dt, (args.steps + 1) / dt, acc))
    for name, mean, std, true in zip(PRIOR, burn.mean(0), burn.std(0), TRUE.values()):
        print('  %-5s %8.3f +- %.3f   (true %g)' % (name, mean, std, true))


if __name__ == '__main__':
    main()
