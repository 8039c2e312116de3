"""
`DeviceObservable`: a 'simulated' Observable whose [N, m] block lives in HBM.

`Simulations.append(name=key, data=<Observable instance>)` stores the object as given
(observables/observable_dict.py:306-307), so the GPU simulator hands the likelihood a
device-resident ensemble without the reference's per-realisation `np.vstack`
(observables/observable.py:205).  `.data` is a lazy device->host copy for any host consumer
(plots, the reference's own likelihoods).
"""
import numpy as np

from ._bases import Observable


class _DeviceRows:
    """
    What `global_data` returns for a device-resident ensemble.  `Simulations.sub_sim` takes a row subset
    with `self[k].global_data[indices, :]` and appends the result to a new `Simulations`
    (observables/observable_dict.py:342-364; the likelihood convergence report, pipelines/pipeline.py:
    1035-1124, does that for every ensemble size); `append` accepts an `Observable` as it is
    (observable_dict.py:306-307), so indexing rows here returns a `DeviceObservable` of the subset and the
    whole report stays on the device.  Every other use sees the host array.
    """

    def __init__(self, obs):
        self._obs = obs

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 2 and isinstance(idx[1], slice) and idx[1] == slice(None):
            rows = np.arange(self._obs.shape[0])[idx[0]]
            return self._obs.take_rows(np.atleast_1d(rows))
        return self._obs.data[idx]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._obs.data, dtype=dtype)

    def __len__(self):
        return self._obs.shape[0]

    @property
    def shape(self):
        return self._obs.shape

    @property
    def dtype(self):
        return np.dtype(np.float64)

    def __getattr__(self, name):              # anything else NumPy-like: the host array
        return getattr(self._obs.data, name)


class DeviceObservable(Observable):
    def __init__(self, tensor, unit=None, coords=None, otype=None):
        assert tensor.dim() == 2, 'simulated observables are [ensemble, size]'
        self._dtype = 'simulated'
        self.otype = otype
        self.unit = unit
        self.coords = coords
        self._rw_flag = False
        self._tensor = tensor
        self._host = None

    # device face ----------------------------------------------------------------------------
    @property
    def tensor(self):
        return self._tensor

    def take_rows(self, indices):
        """Row subset (bootstrap resampling, Simulations.sub_sim) without leaving the device."""
        import torch
        idx = torch.as_tensor(np.asarray(indices), device=self._tensor.device)
        if idx.dtype == torch.bool:
            idx = idx.nonzero().reshape(-1)
        return DeviceObservable(self._tensor.index_select(0, idx.long()), self.unit, self.coords, self.otype)

    def take_columns(self, keep):
        """mask_obs on device (tools/masker.py:25-47)."""
        import torch
        from .. import engine
        idx = torch.as_tensor(np.asarray(keep, dtype=np.int32), device=self._tensor.device)
        return DeviceObservable(engine.gather_columns(self._tensor, idx), self.unit, self.coords, 'plain')

    # host face (reference contract) -------------------------------------------------------------
    @property
    def data(self):
        if self._host is None:
            self._host = self._tensor.cpu().numpy()
        return self._host

    @data.setter
    def data(self, value):
        raise AttributeError('DeviceObservable is read-only; build a new one')

    @property
    def global_data(self):
        return _DeviceRows(self)

    @property
    def shape(self):
        return tuple(self._tensor.shape)

    @property
    def size(self):
        return int(self._tensor.shape[1])

    @property
    def ensemble_mean(self):
        from .. import engine
        return engine.ensemble_mean(self._tensor).cpu().numpy()

    def append(self, new_data):
        raise TypeError('DeviceObservable holds the whole ensemble; append is not supported')

    def __getstate__(self):
        # checkpoints (Pipeline.save, tools/io.py:25-75) must not hold device pointers
        state = dict(self.__dict__)
        state['_host'] = self.data
        state['_tensor'] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self._tensor is None and self._host is not None:
            import torch
            if torch.cuda.is_available():
                self._tensor = torch.from_numpy(self._host).cuda()


class ShardedDeviceObservable(Observable):
    """
    A simulated observable whose PIXELS are sharded over ranks: after the all-to-all of the
    simulator's destination-major output this rank holds all N realisations of its m/P pixels,
    as key `index` of the buffer `recv` [P(source), K, n_local, m/P].  `B200EnsembleLikelihood`
    consumes it in place (phased Gram route, two small all-reduces).  Anything that needs the
    full [N, m] block (`.tensor`, `.data`, masks, other likelihoods) triggers an all-gather --
    a collective, so every rank must ask for it.
    """

    def __init__(self, recv, index, n_real, m, rank, world, group=None, unit=None, coords=None, otype=None,
                 still_valid=None):
        self._dtype = 'simulated'
        self.otype, self.unit, self.coords = otype, unit, coords
        self._rw_flag = False
        self.recv, self.index = recv, int(index)
        self.n_real, self.m = int(n_real), int(m)
        self.rank, self.world, self.group = int(rank), int(world), group
        self._full = None
        self._host = None
        # exchange='peer': `recv` is the simulator's persistent NVLink buffer, which the NEXT simulator call
        # (and the other ranks' kernels) overwrite; `still_valid()` turns a late read into an error
        self._still_valid = still_valid

    def _check_fresh(self):
        if self._still_valid is not None and not self._still_valid():
            raise RuntimeError('this Simulations object belongs to an earlier evaluation: with exchange="peer" '
                               'the maps live in the simulator\'s NVLink receive buffer, which the next '
                               'simulator call overwrites; read `.tensor` / `.data` (or evaluate the '
                               'likelihood) before simulating again, or use exchange="alltoall"')

    # layout of the rows of this key inside `recv` (imb_gram_phase*): rows_per_block, strides
    @property
    def n_local(self):
        return int(self.recv.shape[2])

    @property
    def m_local(self):
        return int(self.recv.shape[3])

    @property
    def row_layout(self):
        return (self.n_local, int(self.recv.shape[1]) * self.n_local * self.m_local, self.m_local)

    @property
    def base(self):
        self._check_fresh()
        return self.recv[0, self.index]

    @property
    def pixel_bounds(self):
        return self.rank * self.m_local, (self.rank + 1) * self.m_local

    @property
    def local_block(self):
        """[N, m_local]: every realisation of this rank's pixels (copy)."""
        self._check_fresh()
        return self.recv[:, self.index].reshape(self.n_real, self.m_local)

    @property
    def tensor(self):
        if self._full is None:
            import torch
            import torch.distributed as dist
            local = self.local_block.contiguous()
            parts = torch.empty((self.world, self.n_real, self.m_local), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(parts, local, group=self.group)
            self._full = parts.permute(1, 0, 2).reshape(self.n_real, self.m).contiguous()
        return self._full

    @property
    def data(self):
        if self._host is None:
            self._host = self.tensor.cpu().numpy()
        return self._host

    @data.setter
    def data(self, value):
        raise AttributeError('ShardedDeviceObservable is read-only')

    global_data = property(lambda self: self.data)
    shape = property(lambda self: (self.n_real, self.m))
    size = property(lambda self: self.m)

    @property
    def ensemble_mean(self):
        from .. import engine
        return engine.ensemble_mean(self.tensor).cpu().numpy()

    def append(self, new_data):
        raise TypeError('ShardedDeviceObservable holds the whole ensemble; append is not supported')

    def __getstate__(self):
        raise TypeError('pixel-sharded observables are transient (one evaluation) and not picklable; '
                        'pickle `.data` instead')
