"""
GPU Likelihood drop-ins: same constructors, same `call(simulations_dict) -> float`, same
assertions as the reference classes

  EnsembleLikelihood           likelihoods/ensemble_likelihood.py:20-107
  EnsembleLikelihoodDiagonal   likelihoods/ensemble_likelihood.py:110-173
  SimpleLikelihood             likelihoods/simple_likelihood.py:19-84

Measurements and (co)variances are constant across evaluations, so they are uploaded once
(after masking, like `Likelihood.measurement_dict`'s setter, likelihood.py:114-148).  Masks
are applied to the simulated ensemble on the device; the masked key is renamed
`(name, freq, n_unmasked, tag)` exactly like `Masks.__call__` (observable_dict.py:203-240).
The simulated block may be a `DeviceObservable` (no copy) or a host `Observable` (uploaded).
"""
import numpy as np

from ._bases import Likelihood, Simulations, Masks
from .device_observable import DeviceObservable, ShardedDeviceObservable
from .. import engine


def _torch():
    import torch
    return torch


_ESTIMATOR_NAMES = {'oas_mcov': 'oas', 'empirical_mcov': 'empirical', 'diagonal_mcov': 'diagonal',
                    'oas': 'oas', 'empirical': 'empirical', 'diagonal': 'diagonal'}


def _estimator_name(cov_func):
    """The reference passes a function object (covariance_estimator.py); map it by name."""
    if cov_func is None:
        return 'oas'
    name = cov_func if isinstance(cov_func, str) else getattr(cov_func, '__name__', None)
    if name not in _ESTIMATOR_NAMES:
        raise NotImplementedError(
            'cov_func %r has no CUDA implementation (supported: oas_mcov, empirical_mcov, '
            'diagonal_mcov); there is no CPU fallback' % (cov_func,))
    return _ESTIMATOR_NAMES[name]


class _DeviceLikelihoodMixin:
    """Device copies of measurements / covariances and on-device masking of simulations."""

    def _reset_device_state(self):
        self._dev_meas, self._dev_cov, self._dev_keep = {}, {}, {}
        self._pinned = {}
        self._shard_bufs = {}
        self._copy_stream, self._upload_event, self._last_use = None, None, None

    def _host_array(self, kind, key):
        if kind == 'meas':
            return np.ascontiguousarray(self._measurement_dict[key].data[0], dtype=np.float64)
        obs = self._covariance_dict[key]
        return np.ascontiguousarray(obs.var if obs.dtype == 'variance' else obs.data, dtype=np.float64)

    def upload_measurements(self):
        """
        (Re-)upload measurements and variances from pinned host memory, asynchronously on a copy
        stream, so the transfer overlaps whatever the caller launches next (normally the
        simulator).  The next `call` waits for it on the device only.  Dense covariances keep
        the lazy, blocking upload (they are too large to pin per Likelihood).
        """
        t = _torch()
        if self._copy_stream is None:
            self._copy_stream = t.cuda.Stream()
        jobs = []
        for kind, store in (('meas', self._dev_meas), ('cov', self._dev_cov)):
            if kind == 'cov' and self._covariance_dict is None:
                continue
            source = self._measurement_dict if kind == 'meas' else self._covariance_dict
            for key in source.keys():
                if kind == 'cov' and source[key].dtype != 'variance':
                    continue
                pin = self._pinned.get((kind, key))
                if pin is None:
                    pin = t.from_numpy(self._host_array(kind, key)).pin_memory()
                    self._pinned[(kind, key)] = pin
                if key not in store:
                    store[key] = t.empty(pin.shape, dtype=t.float64, device='cuda')
                jobs.append((store[key], pin))
        # the buffers may still be read by the previous evaluation (not by work queued since, e.g.
        # the simulator kernels of this one: waiting for those would serialise the copy behind them)
        if self._last_use is not None:
            self._copy_stream.wait_event(self._last_use)
        with t.cuda.stream(self._copy_stream):
            for dev, pin in jobs:
                dev.copy_(pin, non_blocking=True)
            self._upload_event = self._copy_stream.record_event()

    def _await_upload(self):
        if self._upload_event is not None:
            _torch().cuda.current_stream().wait_event(self._upload_event)
            self._upload_event = None

    def _measured(self, key):
        self._await_upload()
        if key not in self._dev_meas:
            self._dev_meas[key] = _torch().from_numpy(self._host_array('meas', key)).cuda()
        return self._dev_meas[key]

    def _covariance(self, key):
        """[m] for the 'variance' dtype (never expanded with np.diag), [m,m] otherwise."""
        self._await_upload()
        if key not in self._dev_cov:
            self._dev_cov[key] = _torch().from_numpy(self._host_array('cov', key)).cuda()
        return self._dev_cov[key]

    def _variance(self, key):
        c = self._covariance(key)
        return c if c.dim() == 1 else c.diagonal().contiguous()

    def _block(self, obs):
        if isinstance(obs, (DeviceObservable, ShardedDeviceObservable)):
            return obs.tensor                       # sharded: all-gather (collective)
        return _torch().from_numpy(np.ascontiguousarray(obs.data, dtype=np.float64)).cuda()

    def _masked_simulations(self, observable_dict):
        """Masks.__call__ for the simulated ensemble, on the device."""
        if self._mask_dict is None:
            return observable_dict
        t = _torch()
        out = Simulations()
        for name in observable_dict.keys():
            obs = observable_dict[name]
            if name not in self._mask_dict.keys():
                out.append(name=name, data=obs)
                continue
            if name not in self._dev_keep:
                self._dev_keep[name] = t.from_numpy(self._mask_dict.kept_indices(name)).cuda() \
                    if hasattr(self._mask_dict, 'kept_indices') else \
                    t.from_numpy(np.flatnonzero(self._mask_dict[name].data[0].astype(bool)).astype(np.int32)).cuda()
            keep = self._dev_keep[name]
            block = engine.gather_columns(self._block(obs), keep)
            new_name = (name[0], name[1], int(keep.numel()), name[3])
            out.append(name=new_name, data=DeviceObservable(block, obs.unit, obs.coords, 'plain'))
        return out

    def __call__(self, observable_dict, **kwargs):
        # same flow as Likelihood.__call__ (likelihood.py:79-92) with the mask step on device
        observable_dict = self._masked_simulations(observable_dict)
        value = self.call(observable_dict, **kwargs)
        if not self.compute_dispersion:
            self._mark_last_use()
            return value
        sample = [self._bootstrapped_likelihood(observable_dict, **kwargs)
                  for _ in range(self.n_bootstrap)]
        self._mark_last_use()
        return value, np.std(sample)

    def _mark_last_use(self):
        if self._copy_stream is not None:
            self._last_use = _torch().cuda.current_stream().record_event()

    def _bootstrapped_likelihood(self, simulations, **kwargs):
        # likelihood.py:94-100: rows resampled with replacement (on the device)
        size = simulations[next(iter(simulations.keys()))].shape[0]
        idx = np.random.randint(0, size, size)
        sub = Simulations()
        for k in simulations.keys():
            obs = simulations[k]
            block = self._block(obs).index_select(0, _torch().from_numpy(idx).cuda())
            sub.append(name=k, data=DeviceObservable(block, obs.unit, obs.coords, obs.otype))
        return self.call(sub, **kwargs)

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ('_dev_meas', '_dev_cov', '_dev_keep', '_shard_bufs', '_pinned'):
            state[k] = {}
        state['_copy_stream'] = state['_upload_event'] = state['_last_use'] = None
        return state


class B200EnsembleLikelihood(_DeviceLikelihoodMixin, Likelihood):
    """
    Gaussian log-likelihood with the covariance of the stochastic fields estimated from the
    ensemble.  `cov_func` accepts the reference's estimator functions (or their names);
    `route` = 'auto' | 'gram' | 'dense' selects the O(m N^2) Gram route (diagonal measurement
    covariance) or the dense DMMA + Cholesky/LU route.
    """

    def __init__(self, measurement_dict, covariance_dict=None, mask_dict=None, cov_func=None,
                 use_trace_approximation=False, route='auto', **kwargs):
        self._reset_device_state()
        super().__init__(measurement_dict, covariance_dict=covariance_dict, mask_dict=mask_dict, **kwargs)
        assert self._covariance_dict is not None
        self.cov_func = cov_func
        self.estimator = _estimator_name(cov_func)
        self.use_trace_approximation = use_trace_approximation
        self.route = route

    def call(self, simulations_dict):
        assert isinstance(simulations_dict, Simulations)
        assert set(simulations_dict.keys()).issubset(self._measurement_dict.keys())
        assert set(simulations_dict.keys()).issubset(self._covariance_dict.keys())
        gram_ok = (self.estimator in ('oas', 'empirical') and not self.use_trace_approximation
                   and self.route in ('auto', 'gram'))
        if gram_ok and self._all_pixel_sharded(simulations_dict):
            return self._call_pixel_sharded(simulations_dict)
        # keys that may take the Gram route are evaluated in batches of identical [N, m]
        # (one launch sequence + one read-back per batch); the others go one by one
        names = list(simulations_dict.keys())
        values, batches = {}, {}
        gram_ok = (self.estimator in ('oas', 'empirical') and not self.use_trace_approximation
                   and self.route in ('auto', 'gram'))
        for name in names:
            block = self._block(simulations_dict[name])
            cov = self._covariance(name)
            if gram_ok and cov.dim() == 1:
                batches.setdefault(tuple(block.shape), []).append((name, block, cov))
            else:
                values[name] = engine.loglike_ensemble(
                    block, self._measured(name), cov, estimator=self.estimator, route=self.route,
                    use_trace_approximation=self.use_trace_approximation)
        for items in batches.values():
            vals = engine.loglike_ensemble_batch([b for _, b, _ in items],
                                                 [self._measured(n) for n, _, _ in items],
                                                 [c for _, _, c in items], estimator=self.estimator)
            values.update({n: v for (n, _, _), v in zip(items, vals)})
        total = 0.
        for name in names:                       # summed in dictionary order, like the reference
            total += values[name]
        return total


    # -- pixel-sharded ensembles (multi-rank): phased Gram route, two small all-reduces ----------------
    def _all_pixel_sharded(self, simulations_dict):
        obs = [simulations_dict[n] for n in simulations_dict]
        if not obs or not all(isinstance(o, ShardedDeviceObservable) for o in obs):
            return False
        same = all(o.recv is obs[0].recv for o in obs)
        variance = all(self._covariance_dict[n].dtype == 'variance' for n in simulations_dict)
        return same and variance and len(obs) <= 16

    def _call_pixel_sharded(self, simulations_dict):
        import torch.distributed as dist
        torch = _torch()
        names = list(simulations_dict.keys())
        obs = [simulations_dict[n] for n in names]
        o0 = obs[0]
        K, N, mp = len(obs), o0.n_real, o0.m_local
        plo, phi = o0.pixel_bounds
        meas = [self._measured(n)[plo:phi] for n in names]
        var = [self._covariance(n)[plo:phi] for n in names]
        sims = [o.base for o in obs]
        key = (K, N)
        if self._shard_bufs.get('key') != key:
            self._shard_bufs = {'key': key,
                                'G': torch.empty((K, N, N), dtype=torch.float64, device='cuda'),
                                'P2': torch.empty((K, N * N + N + 4), dtype=torch.float64, device='cuda')}
        G, P2 = self._shard_bufs['G'], self._shard_bufs['P2']
        from ..hotpath import ensure_library_comm
        if ensure_library_comm(o0.rank, o0.world, o0.group):
            # phases and both all-reduces in ONE library call on the kernels' stream (csrc/comm.cu)
            vals = engine.loglike_ensemble_sharded(sims, N, mp, o0.row_layout, meas, var, o0.m, G, P2,
                                                   estimator=self.estimator)
        else:
            engine.gram_phase1(sims, N, mp, o0.row_layout, meas, G)
            dist.all_reduce(G, op=dist.ReduceOp.SUM, group=o0.group)
            engine.gram_phase2(sims, N, mp, o0.row_layout, meas, var, o0.m, G, P2, estimator=self.estimator)
            dist.all_reduce(P2, op=dist.ReduceOp.SUM, group=o0.group)
            vals = engine.gram_phase3(K, N, mp, G, P2)
        total = 0.
        for v in vals:
            total += v
        return total


class B200EnsembleLikelihoodDiagonal(_DeviceLikelihoodMixin, Likelihood):
    def __init__(self, measurement_dict, covariance_dict=None, mask_dict=None, **kwargs):
        self._reset_device_state()
        super().__init__(measurement_dict, covariance_dict=covariance_dict, mask_dict=mask_dict, **kwargs)
        assert self._covariance_dict is not None

    def call(self, simulations_dict):
        assert isinstance(simulations_dict, Simulations)
        assert set(simulations_dict.keys()).issubset(self._measurement_dict.keys())
        assert set(simulations_dict.keys()).issubset(self._covariance_dict.keys())
        total = 0.
        for name in simulations_dict.keys():
            total += engine.loglike_ensemble_diagonal(self._block(simulations_dict[name]),
                                                      self._measured(name), self._variance(name))
        return total


class B200SimpleLikelihood(_DeviceLikelihoodMixin, Likelihood):
    def __init__(self, measurement_dict, covariance_dict=None, mask_dict=None, **kwargs):
        self._reset_device_state()
        super().__init__(measurement_dict, covariance_dict=covariance_dict, mask_dict=mask_dict, **kwargs)

    def call(self, simulations_dict):
        assert isinstance(simulations_dict, Simulations)
        assert set(simulations_dict.keys()).issubset(self._measurement_dict.keys())
        cov_keys = self._covariance_dict.keys() if self._covariance_dict is not None else ()
        total = 0
        for name in self._measurement_dict:          # KeyError if a measurement was not simulated
            sims = self._block(simulations_dict[name])
            cov = self._covariance(name) if name in cov_keys else None
            total += engine.loglike_simple(sims, self._measured(name), cov)
        return total
