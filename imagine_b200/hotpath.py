"""
One ensemble log-likelihood evaluation on device: fields -> LOS maps -> log-likelihood.

`EnsembleEvaluator` owns the device buffers of one parameter-point evaluation and shards the
ensemble over ranks the way the reference's `distribute_ensemble` does
(pipelines/pipeline.py:737-744: contiguous blocks of N/P realisations, N a multiple of P).
Randomness is keyed on the GLOBAL realisation seed, so results do not depend on P.  The only
exchange is one NCCL all-gather of the simulated maps per observable; every rank then forms
the N x N Gram matrix redundantly and returns the identical scalar ("copied to all nodes",
likelihoods/ensemble_likelihood.py:79-80).  The GPU Simulator / Likelihood drop-ins
(api/gpu_simulator.py, api/gpu_likelihood.py) are thin shells over this class.
"""
import numpy as np

from . import engine

# observable name -> (LOS output, needs)
LOS_OUTPUT = {'fd': 'fd', 'dm': 'dm', 'I': 'I', 'Q': 'Q', 'U': 'U'}


def shard_bounds(n_real, rank, world):
    """Contiguous realisation block of `rank` (pipeline.py:737-744: N must be a multiple of P)."""
    if n_real % world != 0:
        raise ValueError('ensemble size %d is not a multiple of the number of ranks %d' % (n_real, world))
    per = n_real // world
    return rank * per, (rank + 1) * per


def gather_rows(local, n_real, group=None):
    """
    All-gather of per-rank row blocks [n_local, m] into the global ensemble [n_real, m], rows in
    global realisation order.  Backend-agnostic (NCCL on device tensors, gloo on CPU tensors in
    the host-logic tests); a single-rank call returns `local` itself.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    full = torch.empty((n_real, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full


class EnsembleEvaluator:
    def __init__(self, box, resolution, nside, los_settings, quantities=('fd', 'dm', 'I', 'Q', 'U'),
                 n_real=1, rank=0, world=1, b_stride=16, group=None):
        import torch
        self.torch = torch
        self.grid = engine.DeviceGrid(box, resolution)
        self.nside = int(nside)
        self.m = 12 * self.nside ** 2
        self.settings = los_settings
        self.quantities = tuple(quantities)
        self.n_real = int(n_real)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.lo, self.hi = shard_bounds(self.n_real, self.rank, self.world)
        self.n_local = self.hi - self.lo
        self.b_stride = int(b_stride)
        self.dirs = engine.healpix_ring_dirs(self.nside)
        self._field_bufs = {}
        dev = self.dirs.device
        self.local_maps = {q: torch.empty((self.n_local, self.m), dtype=torch.float64, device=dev)
                           for q in self.quantities}
        if self.world > 1:
            self.full_maps = {q: torch.empty((self.n_real, self.m), dtype=torch.float64, device=dev)
                              for q in self.quantities}
        else:
            self.full_maps = self.local_maps

    # -- stage 1 + 2 -----------------------------------------------------------------------
    def fields(self, descs, seeds):
        """Evaluate this rank's realisations; buffers are reused across evaluations."""
        seeds = np.atleast_2d(np.asarray(seeds, dtype=np.uint64))
        if seeds.shape[1] != self.n_real:
            raise ValueError('need one seed per (global) realisation')
        shapes, _ = engine.field_shapes(self.grid, descs, self.n_local, self.b_stride)
        out = {}
        for name, shape in shapes.items():
            buf = self._field_bufs.get(name)
            if buf is None or tuple(buf.shape) != shape:
                buf = self.torch.empty(shape, dtype=self.torch.float64, device=self.dirs.device)
                self._field_bufs[name] = buf
            out[name] = buf
        return engine.field_eval(self.grid, descs, seeds[:, self.lo:self.hi], n_real=self.n_local,
                                 b_stride=self.b_stride, out=out)

    def los(self, fields):
        want = tuple(LOS_OUTPUT[q] for q in self.quantities)
        return engine.los_integrate(self.grid, self.settings, self.dirs, fields['magnetic_field'],
                                    fields.get('thermal_electron_density'),
                                    fields.get('cosmic_ray_electron_density'), n_real=self.n_local,
                                    want=want, out={LOS_OUTPUT[q]: self.local_maps[q] for q in self.quantities})

    def simulate(self, descs, seeds):
        """Local maps {quantity: [n_local, m]} for this rank's block of realisations."""
        self.los(self.fields(descs, seeds))
        return self.local_maps

    # -- exchange ----------------------------------------------------------------------------
    def gather(self):
        """All ranks end with every realisation's maps, rows in global order."""
        if self.world == 1:
            return self.local_maps
        import torch.distributed as dist
        for q in self.quantities:
            dist.all_gather_into_tensor(self.full_maps[q], self.local_maps[q], group=self.group)
        return self.full_maps

    # -- stage 3 -------------------------------------------------------------------------------
    def loglike(self, maps, meas, var, estimator='oas'):
        """Sum over observables of EnsembleLikelihood's per-key term (diagonal measurement cov)."""
        qs = self.quantities
        vals = engine.loglike_ensemble_batch([maps[q] for q in qs], [meas[q] for q in qs],
                                             [var[q] for q in qs], estimator=estimator)
        total = 0.0
        for v in vals:                      # same left-to-right sum as the reference's loop over keys
            total += v
        return total

    def evaluate(self, descs, seeds, meas, var, estimator='oas'):
        self.simulate(descs, seeds)
        return self.loglike(self.gather(), meas, var, estimator)
