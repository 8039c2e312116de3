"""
One ensemble log-likelihood evaluation on device: fields -> LOS maps -> log-likelihood.

`EnsembleEvaluator` owns the device buffers of one parameter-point evaluation and shards the
ensemble over ranks the way the reference's `distribute_ensemble` does
(pipelines/pipeline.py:737-744: contiguous blocks of N/P realisations, N a multiple of P).
Randomness is keyed on the GLOBAL realisation seed, so results do not depend on P.

Two exchange modes between the simulation and the likelihood:

  'alltoall'  (default when P divides the pixel count)  The LOS kernel writes its maps directly in
              destination-major order [P][K][n_local][m/P]; ONE all-to-all then gives every rank
              all N realisations of its m/P pixels of every observable (1/P of the all-gather
              volume).  The Gram route runs in three phases on the local pixels with two small
              SUM all-reduces in between (G: K*N*N doubles, P2: K*(N*N+N+4) doubles) -- the
              exchange the reference does with mpi4py Allreduce / ring matmul
              (tools/mpi_helper.py:132, 234-246).  The likelihood work is sharded P ways.
  'peer'      Same data flow as 'alltoall', but the collective is FUSED into the LOS kernel: the
              receive buffer is NVLink symmetric memory (torch.distributed._symmetric_memory) and
              the kernel's epilogue stores the block of destination rank d straight into rank d's
              buffer (peer stores, `imb_los_cfg.peer_delta`), overlapping the exchange with the
              ray march; a signal-pad barrier replaces the all-to-all.  On request only: with row-major
              maps the remote stores of a warp are 32 x 8 bytes and the NCCL all-to-all below is faster
              (8 B200s, C3: 835 vs 760 evals/s; C4: 66 vs 49).
  'allgather' Every rank receives every realisation's full maps (one all-gather per observable)
              and evaluates the likelihood redundantly; what the plugin API needs, since
              `Simulations` must expose the full [N, m] block.

Either way every rank returns the identical scalar ("copied to all nodes",
likelihoods/ensemble_likelihood.py:79-80).

Several parameter points at once: `evaluate_many` (asynchronous pipelining), `evaluate_batch` (up to four
points per LAUNCH: the lanes of the ray march are (point, realisation) pairs, one exchange and one chain of
Gram phases for points x observables keys), `evaluate_points` (one point per GPU).
"""
import logging

import numpy as np

from . import engine

log = logging.getLogger(__name__)

LOS_OUTPUT = {'fd': 'fd', 'dm': 'dm', 'I': 'I', 'Q': 'Q', 'U': 'U'}


def shard_bounds(n_real, rank, world):
    """Contiguous realisation block of `rank` (pipeline.py:737-744: N must be a multiple of P)."""
    if n_real % world != 0:
        raise ValueError('ensemble size %d is not a multiple of the number of ranks %d' % (n_real, world))
    per = n_real // world
    return rank * per, (rank + 1) * per


def pixel_bounds(m, rank, world):
    """Contiguous pixel block of `rank` in the pixel-sharded likelihood (needs P | m)."""
    if m % world != 0:
        raise ValueError('pixel count %d is not a multiple of the number of ranks %d' % (m, world))
    per = m // world
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


_LIBRARY_COMM = {'tried': False, 'ok': False}


def ensure_library_comm(rank, world, group=None):
    """
    Initialise the library's OWN NCCL communicator (csrc/comm.cu) next to torch.distributed's, once per
    process: the all-to-all and the two all-reduces of a sharded evaluation then run as plain library calls
    on the kernels' stream -- no cross-stream events, a fraction of torch.distributed's host cost per
    collective (at 8 GPUs one evaluation is 1.1 ms, the host side of the plugin path about as much).
    The 128-byte id travels through a torch.distributed broadcast.  Every rank first reports whether it
    can load NCCL at all, so a rank that cannot never leaves the others waiting in the init.  Returns
    True when the library communicator spans exactly the default group's ranks; IMB_LIBRARY_COLLECTIVES=0
    switches it off.
    """
    import os
    import torch
    import torch.distributed as dist
    if world <= 1 or group is not None or os.environ.get('IMB_LIBRARY_COLLECTIVES', '1') == '0':
        return False
    if _LIBRARY_COMM['tried']:
        return _LIBRARY_COMM['ok'] and engine.comm_size() == world
    _LIBRARY_COMM['tried'] = True
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != world:
        return False
    dev = torch.device('cuda', torch.cuda.current_device())
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    can = torch.ones(1, dtype=torch.int32, device=dev)
    try:
        uid = engine.comm_unique_id()                    # loads NCCL; only rank 0's id is used
        if rank == 0:
            idt = torch.frombuffer(bytearray(uid), dtype=torch.uint8).to(dev)
    except Exception as exc:
        can.zero_()
        log.warning('library collectives unavailable on rank %d: %s', rank, exc)
    dist.all_reduce(can, op=dist.ReduceOp.MIN)
    if int(can.item()) == 0:
        return False
    dist.broadcast(idt, 0)
    engine.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    _LIBRARY_COMM['ok'] = True
    return True


def exchange_pixel_shards(send, group=None):
    """
    All-to-all of a destination-major buffer [P, K, n_local, m/P]: block p goes to rank p; the
    result [P, K, n_local, m/P] holds, for source rank q, its n_local realisations of MY pixels.
    Backend-agnostic (NCCL / gloo).  Single rank: identity.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return send
    recv = torch.empty_like(send)
    if send.is_cuda and group is None and _LIBRARY_COMM['ok'] and engine.comm_size() == send.shape[0]:
        return engine.alltoall(send, recv)
    dist.all_to_all_single(recv, send, group=group)
    return recv


class PeerExchange:
    """
    Receive buffer [P(source), K, n_local, m/P] in NVLink symmetric memory, the offsets of every
    peer's copy relative to mine (what `imb_los_cfg.peer_delta` wants) and the barrier that makes
    the peer stores of all ranks' LOS kernels visible.  Raises if symmetric memory is unavailable.
    """

    def __init__(self, shape, rank, group=None, device='cuda'):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.recv = symm_mem.empty(tuple(shape), dtype=torch.float64, device=device)
        group = group if group is not None else dist.group.WORLD
        self._handle = symm_mem.rendezvous(self.recv, group.group_name)
        ptrs = [int(p) for p in self._handle.buffer_ptrs]
        mine = ptrs[rank]
        if mine != self.recv.data_ptr() or any((p - mine) % 8 for p in ptrs):
            raise RuntimeError('unexpected symmetric-memory mapping')
        self.peer_delta = [(p - mine) // 8 for p in ptrs]

    def barrier(self):
        self._handle.barrier()


class EnsembleEvaluator:
    def __init__(self, box, resolution, nside, los_settings, quantities=('fd', 'dm', 'I', 'Q', 'U'),
                 n_real=1, rank=0, world=1, b_stride='auto', group=None, exchange='auto'):
        import torch
        self.torch = torch
        self.grid = engine.DeviceGrid(box, resolution)
        self.nside = int(nside)
        self.m = 12 * self.nside ** 2
        self.settings = los_settings
        self.quantities = tuple(quantities)
        self.K = len(self.quantities)
        self.n_real = int(n_real)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.lo, self.hi = shard_bounds(self.n_real, self.rank, self.world)
        self.n_local = self.hi - self.lo
        self.b_stride = engine.pick_b_stride(self.n_local) if b_stride == 'auto' else int(b_stride)
        self.dirs = engine.healpix_ring_dirs(self.nside)
        self._field_bufs = {}
        self._local = None                          # world = 1 evaluator of evaluate_points
        # collectives through the library's own communicator (imb_comm_*) instead of torch.distributed
        self.library_collectives = ensure_library_comm(self.rank, self.world, group)
        dev = self.dirs.device
        shardable = self.world > 1 and self.m % self.world == 0 and self.world <= 16
        if exchange == 'auto':
            # measured on 8 B200s (profiles/README.md): local stores + one NCCL all-to-all beat the fused
            # peer-store exchange (C3: 835 vs 760 evals/s, C4: 66 vs 49) -- with row-major maps a warp's 32
            # remote stores are 8 bytes each; 'peer' stays available on request
            exchange = 'alltoall' if shardable else 'allgather'
            auto = True
        else:
            auto = False
        f64 = torch.float64
        self._symm = None
        self.exchange_note = None                   # why 'auto' did not end up on the fused exchange
        if auto and not shardable and self.world > 1:
            self.exchange_note = ('all-gather: %d pixels do not split over %d ranks' % (self.m, self.world)
                                  if self.m % self.world else 'all-gather: more than 16 ranks')
        if exchange == 'peer':
            try:
                self._init_peer_buffers(dev)
            except Exception as exc:
                if not auto:
                    raise
                exchange = 'alltoall'               # symmetric memory unavailable: NCCL all-to-all
                self.exchange_note = 'all-to-all: symmetric memory unavailable (%s: %s)' % (type(exc).__name__, exc)
        if self.exchange_note and self.rank == 0:
            log.warning('EnsembleEvaluator exchange fallback -- %s', self.exchange_note)
        self.exchange = exchange
        if self.exchange in ('alltoall', 'peer'):
            self.mp = self.m // self.world
            self.plo, self.phi = pixel_bounds(self.m, self.rank, self.world)
            if self.exchange == 'alltoall':
                # destination-major send buffer the LOS kernel writes into; recv after the all-to-all
                self.send = torch.empty((self.world, self.K, self.n_local, self.mp), dtype=f64, device=dev)
                self.recv = self.send if self.world == 1 else torch.empty_like(self.send)
            self.G = torch.empty((self.K, self.n_real, self.n_real), dtype=f64, device=dev)
            self.P2 = torch.empty((self.K, self.n_real * self.n_real + self.n_real + 4), dtype=f64, device=dev)
            self.local_maps = None
        else:
            self.local_maps = {q: torch.empty((self.n_local, self.m), dtype=f64, device=dev)
                               for q in self.quantities}
            if self.world > 1:
                self.full_maps = {q: torch.empty((self.n_real, self.m), dtype=f64, device=dev)
                                  for q in self.quantities}
            else:
                self.full_maps = self.local_maps
                self.mp, self.plo, self.phi = self.m, 0, self.m      # evaluate_batch: one "shard" = everything

    def _init_peer_buffers(self, dev):
        """Receive buffer [P(source), K, n_local, m/P] in NVLink symmetric memory + peer offsets."""
        mp = self.m // self.world
        self._symm = PeerExchange((self.world, self.K, self.n_local, mp), self.rank, self.group, dev)
        self.recv, self.peer_delta = self._symm.recv, self._symm.peer_delta
        self.send = None

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
        if self.exchange == 'alltoall':
            out = {LOS_OUTPUT[q]: self.send[0, k] for k, q in enumerate(self.quantities)}
            layout = (self.mp, self.K * self.n_local * self.mp)
        elif self.exchange == 'peer':
            # my source slot inside MY receive buffer; the kernel redirects each pixel block to its owner
            out = {LOS_OUTPUT[q]: self.recv[self.rank, k] for k, q in enumerate(self.quantities)}
            layout = (self.mp, 0, self.peer_delta)
        else:
            out = {LOS_OUTPUT[q]: self.local_maps[q] for q in self.quantities}
            layout = None
        return engine.los_integrate(self.grid, self.settings, self.dirs, fields['magnetic_field'],
                                    fields.get('thermal_electron_density'),
                                    fields.get('cosmic_ray_electron_density'), n_real=self.n_local,
                                    want=want, out=out, out_layout=layout)

    def simulate(self, descs, seeds):
        """Run fields + LOS for this rank's block of realisations (outputs stay in device buffers)."""
        self.los(self.fields(descs, seeds))
        return self.local_maps if self.exchange == 'allgather' else (self.send if self.send is not None else self.recv)

    # -- exchange ----------------------------------------------------------------------------
    def gather(self):
        """Move the simulated maps to where the likelihood needs them."""
        if self.exchange == 'peer':
            self._symm.barrier()                    # every rank's LOS kernel (and its peer stores) is done
            return self.recv
        if self.exchange == 'alltoall':
            if self.world > 1:
                self._alltoall(self.send, self.recv)
            return self.recv
        if self.world == 1:
            return self.local_maps
        import torch.distributed as dist
        for q in self.quantities:
            dist.all_gather_into_tensor(self.full_maps[q], self.local_maps[q], group=self.group)
        return self.full_maps

    def full_maps_from_shards(self, recv=None):
        """[K][N, m_local] views of the received pixel shards in global realisation order (tests)."""
        recv = self.recv if recv is None else recv
        return {q: recv[:, k].reshape(self.n_real, self.mp) for k, q in enumerate(self.quantities)}

    # -- stage 3 -------------------------------------------------------------------------------
    def loglike(self, maps, meas, var, estimator='oas'):
        """Sum over observables of EnsembleLikelihood's per-key term (diagonal measurement cov)."""
        qs = self.quantities
        if self.exchange in ('alltoall', 'peer'):
            sims = [maps[0, k] for k in range(self.K)]        # base pointer of key k, blocked rows
            layout = (self.n_local, self.K * self.n_local * self.mp, self.mp)
            meas_l = [meas[q][self.plo:self.phi] for q in qs]
            var_l = [var[q][self.plo:self.phi] for q in qs]
            if self.library_collectives and engine.comm_size() == self.world:   # phases + all-reduces inside the library (csrc/comm.cu)
                vals = engine.loglike_ensemble_sharded(sims, self.n_real, self.mp, layout, meas_l, var_l, self.m,
                                                       self.G, self.P2, estimator=estimator)
            else:
                engine.gram_phase1(sims, self.n_real, self.mp, layout, meas_l, self.G)
                self._allreduce(self.G)
                engine.gram_phase2(sims, self.n_real, self.mp, layout, meas_l, var_l, self.m, self.G, self.P2,
                                   estimator=estimator)
                self._allreduce(self.P2)
                vals = engine.gram_phase3(self.K, self.n_real, self.mp, self.G, self.P2)
        else:
            vals = engine.loglike_ensemble_batch([maps[q] for q in qs], [meas[q] for q in qs],
                                                 [var[q] for q in qs], estimator=estimator)
        total = 0.0
        for v in vals:                      # same left-to-right sum as the reference's loop over keys
            total += v
        return total

    def loglike_async(self, maps, meas, var, out, flags, estimator='oas'):
        """`loglike` without the host read-back: per-key values into `out` [K] (device)."""
        qs = self.quantities
        if self.exchange in ('alltoall', 'peer'):
            sims = [maps[0, k] for k in range(self.K)]
            layout = (self.n_local, self.K * self.n_local * self.mp, self.mp)
            meas_l = [meas[q][self.plo:self.phi] for q in qs]
            var_l = [var[q][self.plo:self.phi] for q in qs]
            engine.gram_phase1(sims, self.n_real, self.mp, layout, meas_l, self.G)
            self._allreduce(self.G)
            engine.gram_phase2(sims, self.n_real, self.mp, layout, meas_l, var_l, self.m, self.G, self.P2,
                               estimator=estimator)
            self._allreduce(self.P2)
            engine.gram_phase3_async(self.K, self.n_real, self.mp, self.G, self.P2, out, flags)
        else:
            engine.loglike_ensemble_batch_async([maps[q] for q in qs], [meas[q] for q in qs],
                                                [var[q] for q in qs], out, flags, estimator=estimator)

    def evaluate_many(self, points, meas, var, estimator='oas'):
        """
        Pipelined evaluation of several parameter points (list of (descs, seeds)): every kernel and
        collective of every point is enqueued back to back and the host reads all values once --
        what a vectorised sampler callback (UltraNest `vectorized=True`, emcee walkers, the loop
        over cubes in pipelines/pipeline.py:976-979) needs.  Returns one log-likelihood per point.
        """
        torch = self.torch
        n = len(points)
        out = torch.empty((n, self.K), dtype=torch.float64, device=self.dirs.device)
        flags = torch.zeros((n, self.K), dtype=torch.int32, device=self.dirs.device)
        for i, (descs, seeds) in enumerate(points):
            self.simulate(descs, seeds)
            self.loglike_async(self.gather(), meas, var, out[i], flags[i], estimator)
        host, bad = out.cpu(), flags.cpu()                      # the one synchronisation
        if bool(bad.any()):
            raise np.linalg.LinAlgError('Gram route not applicable (non-positive shrunk diagonal) for '
                                        'points %s; evaluate them with `evaluate`' % bad.any(1).nonzero().flatten().tolist())
        totals = []
        for row in host.tolist():
            t = 0.0
            for v in row:
                t += v
            totals.append(t)
        return totals

    # -- several parameter points per launch ----------------------------------------------------------
    def batch_capacity(self, points=None):
        """
        How many parameter points one batched pass takes (0: batching does not apply).  The rays, cells
        and trilinear weights of the march do not depend on the parameter point, so the lanes of a warp
        can be (point, realisation) pairs: P points x n_local realisations fill the 32 lanes of the
        fastest ray-march variant, where a sharded evaluation alone uses 8 or 16 of them in ray segments.
        Needs the pixel-sharded exchange (or a single GPU), at most 16 local realisations, a stochastic magnetic field and
        -- because the march shares the (n_e, CRE) interpolants across its lanes -- identical
        non-magnetic fields in every point of the batch (the usual GMF fit).
        """
        if not (self.exchange == 'alltoall' or self.world == 1) or self.n_local > 16:
            return 0
        lanes = 8 if self.n_local <= 8 else 16
        cap = min(32 // lanes, engine.KMAX_BATCH // self.K)
        if points is not None:
            def scalar_part(descs):
                return [(d[0], tuple(float(x) for x in d[1])) for d in descs if engine.field_type(d[0]) != 0]
            first = scalar_part(points[0][0])
            if any(scalar_part(d) != first for d, _ in points[1:]):
                return 0
            if any(engine.field_rows(d, 2)[0] != 2 for d, _ in points):
                return 0                             # no stochastic magnetic field: nothing to interleave
            if max(engine.field_rows(points[0][0], 2)[1:]) > 1:
                return 0                             # per-realisation scalar grids
        return cap if cap >= 2 else 0

    def evaluate_batch(self, points, meas, var, estimator='oas'):
        """
        `evaluate_many` with up to `batch_capacity()` parameter points per LAUNCH: one field kernel per
        point writes its lanes of a shared wide buffer (imb_field_eval_lanes), ONE ray march integrates all
        points of the chunk, ONE all-to-all moves their maps and ONE chain of Gram phases (two all-reduces)
        evaluates all points x observables keys.  Same realisations, same Philox keys as the one-point
        path; the maps differ from it by rounding only (the one-point path splits a ray into segments per
        warp at these shard sizes, this one does not).  Falls back to `evaluate_many` when batching does
        not apply.  Returns one log-likelihood per point.
        """
        cap = self.batch_capacity(points)
        if cap == 0 or len(points) < 2:
            return self.evaluate_many(points, meas, var, estimator)
        torch = self.torch
        dev = self.dirs.device
        f64 = torch.float64
        lanes = 8 if self.n_local <= 8 else 16
        qs = self.quantities
        totals = []
        pending = []
        for c0 in range(0, len(points), cap):
            chunk = points[c0:c0 + cap]
            P = len(chunk)
            rows = P * lanes                                  # rows of the shared ensemble (padded lanes incl.)
            key = ('batch', rows)
            bufs = self._field_bufs.get(key)
            if bufs is None:
                bufs = dict(B=torch.zeros((1, self.grid.ncell, 3, 32), dtype=f64, device=dev),
                            send=torch.empty((self.world, self.K, rows, self.mp), dtype=f64, device=dev),
                            G=torch.empty((P * self.K, self.n_real, self.n_real), dtype=f64, device=dev),
                            P2=torch.empty((P * self.K, self.n_real * self.n_real + self.n_real + 4), dtype=f64,
                                           device=dev))
                bufs['recv'] = bufs['send'] if self.world == 1 else torch.empty_like(bufs['send'])
                self._field_bufs[key] = bufs
            # stage 1: the deterministic scalar grids once (shared by the chunk), B per point into its lanes
            descs0, seeds0 = chunk[0]
            scalar_descs = [d for d in descs0 if engine.field_type(d[0]) != 0]
            shared = engine.field_eval(self.grid, scalar_descs, np.zeros((1, 1), dtype=np.uint64), n_real=1) \
                if scalar_descs else {}
            for p, (descs, seeds) in enumerate(chunk):
                seeds = np.atleast_2d(np.asarray(seeds, dtype=np.uint64))
                engine.field_eval_lanes(self.grid, [d for d in descs if engine.field_type(d[0]) == 0],
                                        seeds[:, self.lo:self.hi], self.n_local, 3 * lanes, bufs['B'], p * lanes)
            # stage 2: one march over (point, realisation) lanes, destination-major output
            out = {LOS_OUTPUT[q]: bufs['send'][0, k] for k, q in enumerate(qs)}
            engine.los_integrate(self.grid, self.settings, self.dirs, bufs['B'],
                                 shared.get('thermal_electron_density'), shared.get('cosmic_ray_electron_density'),
                                 n_real=rows, want=tuple(LOS_OUTPUT[q] for q in qs), out=out,
                                 out_layout=(self.mp, self.K * rows * self.mp))
            if self.world > 1:
                self._alltoall(bufs['send'], bufs['recv'])
            recv = bufs['recv']
            # stage 3: P x K keys; key (p, k) = rows [p lanes, p lanes + n_local) of every source block
            sims = [recv[0, k, p * lanes] for p in range(P) for k in range(self.K)]
            layout = (self.n_local, self.K * rows * self.mp, self.mp)
            meas_l = [meas[q][self.plo:self.phi] for q in qs] * P
            var_l = [var[q][self.plo:self.phi] for q in qs] * P
            vals = torch.empty((P * self.K,), dtype=f64, device=dev)
            flags = torch.zeros((P * self.K,), dtype=torch.int32, device=dev)
            engine.gram_phase1(sims, self.n_real, self.mp, layout, meas_l, bufs['G'])
            self._allreduce(bufs['G'])
            engine.gram_phase2(sims, self.n_real, self.mp, layout, meas_l, var_l, self.m, bufs['G'], bufs['P2'],
                               estimator=estimator)
            self._allreduce(bufs['P2'])
            engine.gram_phase3_async(P * self.K, self.n_real, self.mp, bufs['G'], bufs['P2'], vals, flags)
            pending.append((P, vals, flags))
        for P, vals, flags in pending:                         # the host reads once per chunk, at the end
            if bool(flags.cpu().any()):
                raise np.linalg.LinAlgError('Gram route not applicable (non-positive shrunk diagonal); '
                                            'evaluate these points with `evaluate`')
            host = vals.cpu().reshape(P, self.K).tolist()
            for row in host:
                t = 0.0
                for v in row:
                    t += v
                totals.append(t)
        return totals

    def evaluate_points(self, points, meas, var, estimator='oas'):
        """
        Point-parallel evaluation: parameter point i is evaluated ENTIRELY on rank i % world (all N
        realisations, no data-path collective), every rank then learns every value through one small
        all-gather -- the decomposition of the reference's MPI mode (`Pipeline._mpi_likelihood`,
        pipelines/pipeline.py:945-996: the cubes of a batch are spread over the ranks and the results
        broadcast), which is what fills several GPUs when one evaluation does not.  `points` is the same
        list on every rank.  Values are bit-identical to `world = 1` evaluations (same kernels, same
        order); ranks with fewer points than the others just wait in the all-gather.
        """
        torch = self.torch
        if self._local is None:
            self._local = EnsembleEvaluator(self.grid.box, self.grid.resolution, self.nside, self.settings,
                                            self.quantities, n_real=self.n_real, rank=0, world=1,
                                            exchange='allgather')
        mine = list(range(self.rank, len(points), self.world))
        per_rank = (len(points) + self.world - 1) // self.world
        vals = torch.full((per_rank,), float('nan'), dtype=torch.float64, device=self.dirs.device)
        if mine:
            vals[:len(mine)] = torch.tensor(self._local.evaluate_many([points[i] for i in mine], meas, var, estimator),
                                            dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            allv = torch.empty((self.world, per_rank), dtype=torch.float64, device=self.dirs.device)
            dist.all_gather_into_tensor(allv, vals, group=self.group)
        else:
            allv = vals.reshape(1, -1)
        host = allv.cpu().tolist()
        return [host[i % self.world][i // self.world] for i in range(len(points))]

    def _allreduce(self, t):
        if self.world > 1:
            if self.library_collectives and engine.comm_size() == self.world:
                engine.allreduce_sum(t)
                return
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _alltoall(self, send, recv):
        if self.library_collectives and engine.comm_size() == self.world:
            engine.alltoall(send, recv)
            return
        import torch.distributed as dist
        dist.all_to_all_single(recv, send, group=self.group)

    def evaluate(self, descs, seeds, meas, var, estimator='oas'):
        self.simulate(descs, seeds)
        return self.loglike(self.gather(), meas, var, estimator)
