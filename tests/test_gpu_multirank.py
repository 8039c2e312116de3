"""
GPU, >= 2 devices (skipped on a 1-GPU box): the N > 1 path over NCCL -- ensemble sharding, the
destination-major LOS output + all-to-all, the phased Gram route with its two all-reduces, and the
plugin drop-ins in pixel-sharded mode -- must reproduce the single-GPU numbers.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOX = [[-20.0, 20.0], [-20.0, 20.0], [-4.0, 4.0]]
QUANT = ('fd', 'dm', 'I', 'Q', 'U')
DESCS = [(3, [0.01, 3.0, 0.5], 0), (0, [0.0, 1.0, 0.0], 1), (1, [0.0, 2.0], 2), (6, [0.0217, 5.0, 1.0, 8.3], 3)]
# 16 realisations: one GPU marches them 16 wide (b_stride 48), two GPUs 8 wide (b_stride 24)
N, NSIDE, RES = 16, 8, [24, 24, 12]


def _setup(rank, world, exchange, N=N):
    import torch
    from imagine_b200 import engine
    from imagine_b200.hotpath import EnsembleEvaluator
    settings = engine.LosSettings(dr=0.25)
    ev = EnsembleEvaluator(BOX, RES, NSIDE, settings, QUANT, n_real=N, rank=rank, world=world, exchange=exchange)
    seeds = np.arange(1001, 1001 + N, dtype=np.uint64)
    m = 12 * NSIDE ** 2
    gen = torch.Generator(device='cpu').manual_seed(5)
    meas = {q: (torch.randn(m, dtype=torch.float64, generator=gen) * s).cuda()
            for q, s in zip(QUANT, (30.0, 100.0, 1e-4, 1e-4, 1e-4))}
    var = {q: (torch.rand(m, dtype=torch.float64, generator=gen) + 0.5).cuda() * sc ** 2
           for q, sc in zip(QUANT, (3.0, 10.0, 1e-5, 1e-5, 1e-5))}
    return ev, seeds, meas, var


def _plugin_value(rank, world, exchange='auto', twice=False):
    import imagine_b200.api as api
    u = api.units
    m = 12 * NSIDE ** 2
    ev, seeds, meas, var = _setup(0, 1, 'allgather')
    grid = api.UniformGrid(box=[[b[0], b[1]] * u.kpc for b in BOX], resolution=RES)
    keys = {'fd': ('fd', None, NSIDE, None), 'dm': ('dm', None, NSIDE, None), 'I': ('sync', 23.0, NSIDE, 'I'),
            'Q': ('sync', 23.0, NSIDE, 'Q'), 'U': ('sync', 23.0, NSIDE, 'U')}
    M = api.Measurements()
    for q in QUANT:
        M.append(name=keys[q], data=meas[q].cpu().numpy().reshape(1, m), cov_data=var[q].cpu().numpy(),
                 otype='HEALPix')
    sim = api.B200Simulator(M, radial_step=0.25, rank=rank, world=world, exchange=exchange)
    lik = api.B200EnsembleLikelihood(M)
    fields = [api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3,
                                                                'scale_radius': 3 * u.kpc,
                                                                'scale_height': 0.5 * u.kpc}, ensemble_size=N),
              api.ConstantMagneticField(grid, parameters={'Bx': 0 * u.microgauss, 'By': 1 * u.microgauss,
                                                          'Bz': 0 * u.microgauss}, ensemble_size=N),
              api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 2 * u.microgauss},
                                             ensemble_size=N, ensemble_seeds=list(range(1001, 1001 + N))),
              api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5 * u.kpc, 'z0': 1 * u.kpc,
                                                               'r_sun': 8.3 * u.kpc}, ensemble_size=N)]
    sims = sim(fields)
    if twice:                     # exchange='peer': the first Simulations aliases a buffer the second call reuses
        first, sims = sims, sim(fields)
        try:
            lik(first)
            return 'no error', sims
        except RuntimeError as exc:
            return ('stale' if 'earlier evaluation' in str(exc) else repr(exc)), sims
    return lik(sims), sims


def _points(seeds):
    """Three parameter points (the regular field's strength varies) for the point-parallel mode."""
    pts = []
    for by in (0.5, 1.0, 2.0):
        d = list(DESCS)
        d[1] = (0, [0.0, by, 0.0], 1)
        pts.append((d, seeds))
    return pts


def _worker(rank, world, port, queue):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from imagine_b200 import _lib
    _lib.init(rank)
    out = {}
    for exchange in ('peer', 'alltoall', 'allgather'):
        ev, seeds, meas, var = _setup(rank, world, exchange)
        out[exchange] = ev.evaluate(DESCS, seeds, meas, var)
    # one realisation per rank (ensemble size == world size, which distribute_ensemble allows):
    # the wide layout pads the single realisation into an 8-lane group
    for exchange in ('peer', 'allgather'):
        ev, seeds, meas, var = _setup(rank, world, exchange, N=world)
        out['tiny_' + exchange] = ev.evaluate(DESCS, seeds, meas, var)
    ev, seeds, meas, var = _setup(rank, world, 'allgather')
    out['points'] = ev.evaluate_points(_points(seeds), meas, var)
    # several points per launch on the sharded path: one march, one all-to-all, one chain of Gram phases
    evb, seedsb, measb, varb = _setup(rank, world, 'alltoall')
    out['batch_capacity'] = evb.batch_capacity(_points(seedsb))
    out['batched'] = evb.evaluate_batch(_points(seedsb), measb, varb)
    # collectives behind the C ABI: the library's own NCCL communicator (the 128-byte id travels through
    # torch.distributed here; any host transport would do) runs the phased Gram route in ONE call
    from imagine_b200 import engine
    idt = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0:
        idt = torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    engine.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    ev, seeds, meas, var = _setup(rank, world, 'alltoall')
    ev.library_collectives = True
    out['c_abi_collectives'] = ev.evaluate(DESCS, seeds, meas, var)
    t = torch.full((5,), float(rank + 1), dtype=torch.float64, device='cuda')
    out['c_abi_allreduce'] = engine.allreduce_sum(t).cpu().tolist()
    engine.comm_finalize()
    value, sims = _plugin_value(rank, world)
    out['plugin'] = value
    key = ('sync', 23.0, NSIDE, 'Q')
    out['plugin_sharded'] = type(sims[key]).__name__
    out['plugin_full_Q'] = sims[key].data           # collective all-gather on every rank
    out['plugin_peer'] = _plugin_value(rank, world, exchange='peer')[0]
    out['plugin_peer_stale'] = _plugin_value(rank, world, exchange='peer', twice=True)[0]
    queue.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_paths_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from imagine_b200 import _lib
    _lib.init(0)
    ev, seeds, meas, var = _setup(0, 1, 'allgather')
    want = ev.evaluate(DESCS, seeds, meas, var)
    want_plugin, sims1 = _plugin_value(0, 1)
    want_points = [ev.evaluate(d, sd, meas, var) for d, sd in _points(seeds)]
    ev2, seeds2, meas2, var2 = _setup(0, 1, 'allgather', N=2)
    want_tiny = ev2.evaluate(DESCS, seeds2, meas2, var2)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results:
        # same realisations, other ray segmentation (layout follows the shard size): rounding only
        assert abs(out['allgather'] - want) <= 1e-9 * max(1.0, abs(want))
        assert abs(out['alltoall'] - want) <= 1e-9 * max(1.0, abs(want))  # different summation order
        assert out['peer'] == out['alltoall']                             # fused NVLink exchange: same data, same kernels
        assert out['plugin_sharded'] == 'ShardedDeviceObservable'
        assert out['points'] == want_points          # point-parallel mode: whole points per rank, bit-identical
        assert out['batch_capacity'] == 4            # 8 local realisations: four points fill the 32 lanes
        for a, b in zip(out['batched'], want_points):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (a, b)
        assert out['c_abi_collectives'] == out['alltoall']      # same phases, NCCL called by the library itself
        assert out['c_abi_allreduce'] == [3.0] * 5
        for mode in ('tiny_peer', 'tiny_allgather'):
            assert abs(out[mode] - want_tiny) <= 1e-9 * max(1.0, abs(want_tiny)), mode
        assert abs(out['plugin'] - want_plugin) <= 1e-9 * max(1.0, abs(want_plugin))
        assert out['plugin_peer'] == out['plugin']          # same kernels, other transport
        assert out['plugin_peer_stale'] == 'stale'          # a Simulations kept past the next call is refused
        q1 = sims1[('sync', 23.0, NSIDE, 'Q')].data
        assert np.allclose(out['plugin_full_Q'], q1, rtol=1e-9, atol=1e-12 * np.abs(q1).max())
    assert results[0][1]['alltoall'] == results[1][1]['alltoall']         # replicated scalar
    assert results[0][1]['plugin'] == results[1][1]['plugin']
    assert results[0][1]['batched'] == results[1][1]['batched']
