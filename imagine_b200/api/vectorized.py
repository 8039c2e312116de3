"""
`VectorizedLikelihood`: the log-likelihood of SEVERAL parameter points per call, for vectorised sampler
callbacks (UltraNest `vectorized=True` -- the reference passes `vectorized=False`,
pipelines/ultranest_pipeline.py:160 --, emcee walkers, the loop over cubes of
`Pipeline._mpi_likelihood`, pipelines/pipeline.py:976-979).

It is built FROM an ordinary reference `Pipeline` whose simulator is a `B200Simulator` and whose likelihood
is a `B200EnsembleLikelihood`: field factories, priors, ensemble seeds and the likelihood rescaler stay the
pipeline's own (the field objects of every point are produced by `Pipeline._get_observables` itself, with a
recorder standing in for the simulator), the points then go through `EnsembleEvaluator.evaluate_batch`:
one ray march over (point, realisation) lanes, one exchange, one chain of Gram phases (hotpath.py).

Scope: HEALPix 'fd' / 'dm' / 'sync' I, Q, U observables of one Nside and one frequency, GPU field classes
(no host-only plugins, no dependencies), diagonal ('variance') measurement covariances, no masks.  Anything
else raises NotImplementedError -- there is no silent fallback to a different code path.
"""
import numpy as np

from .. import engine
from ..hotpath import EnsembleEvaluator
from . import units as u
from .fields import descriptor_of
from .likelihoods import B200EnsembleLikelihood
from .simulators import B200Simulator


class _Recorder:
    """Stands in for the simulator while `Pipeline._get_observables` builds the fields of one point."""

    def __init__(self):
        self.field_list = None

    def __call__(self, field_list):
        self.field_list = list(field_list)
        return None


class VectorizedLikelihood:
    def __init__(self, pipeline):
        sim, lik = pipeline._simulator, pipeline.likelihood
        if not isinstance(sim, B200Simulator) or not isinstance(lik, B200EnsembleLikelihood):
            raise NotImplementedError('needs a B200Simulator and a B200EnsembleLikelihood in the pipeline')
        if sim.masks is not None or lik.mask_dict is not None:
            raise NotImplementedError('masks are not supported by the vectorised callback')
        if lik.compute_dispersion:
            raise NotImplementedError('bootstrap dispersion is not available per batch')
        if lik.estimator not in ('oas', 'empirical') or getattr(lik, 'use_trace_approximation', False):
            raise NotImplementedError('the vectorised callback covers the OAS and empirical estimators')
        self.pipeline, self.sim, self.lik = pipeline, sim, lik
        tags = {'fd': 'fd', 'dm': 'dm'}
        quantities, freqs, nsides = [], set(), set()
        for key in sim.observables:
            name, freq, nside, tag = key
            if name == 'sync':
                if tag not in ('I', 'Q', 'U'):
                    raise NotImplementedError('derived polarisation observables (%r)' % (tag,))
                freqs.add(float(freq))
                quantities.append(tag)
            else:
                quantities.append(tags[name])
            nsides.add(int(nside))
        if len(nsides) != 1 or len(freqs) > 1:
            raise NotImplementedError('one Nside and one frequency per vectorised callback')
        self.keys = list(sim.observables)
        self.quantities = tuple(quantities)
        self.nside = nsides.pop()
        kw = dict(sim.los_kwargs)
        if freqs:
            kw['nu_ghz'] = freqs.pop()
        self.settings = engine.LosSettings(**kw)
        self._ev = None
        self._meas = None

    # -- measurements in the kernels' canonical units ------------------------------------------------
    def _measurements(self):
        """({q: meas}, {q: var}, offset): data / s and variance / s^2 for the unit scale s of every key; the
        log-likelihood in the measurement's units is the canonical one minus sum_k m log s_k."""
        if self._meas is None:
            meas, var, offset = {}, {}, 0.0
            for key, q in zip(self.keys, self.quantities):
                cov = self.lik._covariance(key)
                if cov.dim() != 1:
                    raise NotImplementedError('dense measurement covariance for %r' % (key,))
                s = float(self.sim._scale(key))
                meas[q] = self.lik._measured(key) / s
                var[q] = cov / (s * s)
                offset += meas[q].numel() * np.log(abs(s))
            self._meas = (meas, var, offset)
        return self._meas

    def _evaluator(self, grid, n_real):
        if self._ev is None or self._ev.n_real != n_real:
            box = [[float(u.value_in(b[0], u.kpc)), float(u.value_in(b[1], u.kpc))] for b in grid.box]
            sim = self.sim
            sharded = (12 * self.nside ** 2) % sim.world == 0 and n_real // sim.world <= 16
            self._ev = EnsembleEvaluator(box, [int(r) for r in grid.resolution], self.nside, self.settings,
                                         self.quantities, n_real=n_real, rank=sim.rank, world=sim.world,
                                         group=sim.group, exchange='alltoall' if sharded else 'auto')
        return self._ev

    def _point(self, values):
        """(descs, seeds) of one parameter point, fields built by the pipeline's own code."""
        rec = _Recorder()
        saved = self.pipeline._simulator
        self.pipeline._simulator = rec
        try:
            self.pipeline._get_observables(values)
        finally:
            self.pipeline._simulator = saved
        descs, seed_rows, grid, sizes = [], [], None, set()
        for f in rec.field_list:
            d = descriptor_of(f)
            if d is None or getattr(f, 'dependencies_list', None):
                raise NotImplementedError('field %r has no device descriptor (host plugin or dependencies)'
                                          % (getattr(f, 'name', f),))
            if getattr(f, 'grid', None) is not None:
                grid = f.grid
            sizes.add(int(f.ensemble_size))
            row = 0
            if f.stochastic_field:
                seed_rows.append(np.asarray(f.ensemble_seeds, dtype=np.uint64))
                row = len(seed_rows) - 1
            descs.append((d[0], d[1], d[2], row))
        if len(sizes) != 1:
            raise AssertionError('All fields should have the same ensemble size')
        n_real = sizes.pop()
        seeds = np.stack(seed_rows) if seed_rows else np.zeros((1, n_real), dtype=np.uint64)
        return descs, seeds, grid, n_real

    def __call__(self, cubes):
        """cubes [P, n_active]: PHYSICAL parameter values per point (after `prior_transform`, what
        `Pipeline._likelihood_function` takes).  Returns the P log-likelihoods as an array."""
        cubes = np.atleast_2d(np.asarray(cubes, dtype=float))
        points, grid, n_real = [], None, None
        for values in cubes:
            descs, seeds, g, n = self._point(values)
            grid, n_real = g, n
            points.append((descs, seeds))
        ev = self._evaluator(grid, n_real)
        meas, var, offset = self._measurements()
        vals = ev.evaluate_batch(points, meas, var, estimator=self.lik.estimator)
        self.pipeline._likelihood_evaluations_counter = \
            getattr(self.pipeline, '_likelihood_evaluations_counter', 0) + len(points)
        return (np.asarray(vals) - offset) * self.pipeline.likelihood_rescaler
