"""
not-gpu: the N > 1 host logic on CPU with world_size-2 gloo -- realisation sharding
(pipelines/pipeline.py:737-744), global-seed keying and the all-gather exchange.  The compute
steps are played by the oracle here (this is a test); on the GPU box the same functions move
device tensors over NCCL.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagine_b200.hotpath import gather_rows, shard_bounds
from oracle import fields as OF
from oracle import healpix as OH
from oracle import likelihood as OK
from oracle import los as OL

BOX, RES, NSIDE, N = [[-20, 20], [-20, 20], [-4, 4]], [10, 9, 6], 2, 4
DESCS = [(OF.CONST_B, [0., 1., 0.], 0), (OF.GAUSS_B, [0., 2.], 1), (OF.EXP_NE, [0.01, 3., 0.5], 2)]
SEEDS = [11, 22, 33, 44]


def _maps(seeds):
    og = OF.Grid(BOX, RES)
    f = OF.evaluate_fields(og, DESCS, seeds)
    out = OL.integrate(og, OL.LosConfig(dr=0.5), OH.pix2vec_ring(NSIDE), f['magnetic_field'],
                       f['thermal_electron_density'])
    return out['fd']


def _worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_bounds(N, rank, world)
    local = torch.from_numpy(_maps(SEEDS[lo:hi]))            # this rank's block of realisations
    full = gather_rows(local, N)
    meas = np.linspace(-1, 1, full.shape[1])
    val = OK.ensemble_loglike_gram(full.numpy(), meas, np.full(full.shape[1], 0.3))
    queue.put((rank, full.numpy(), val))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    assert [shard_bounds(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert shard_bounds(5, 0, 1) == (0, 5)
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)          # N must be a multiple of P (pipeline.py:741-744)


def test_two_rank_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([queue.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _maps(SEEDS)
    meas = np.linspace(-1, 1, single.shape[1])
    want = OK.ensemble_loglike_gram(single, meas, np.full(single.shape[1], 0.3))
    for rank, full, val in results:
        assert np.array_equal(full, single)      # global-seed keying: independent of the rank count
        assert val == want                       # replicated scalar, identical on every rank


def test_gather_rows_single_rank_is_identity():
    x = torch.arange(6.0).reshape(2, 3)
    assert gather_rows(x, 2) is x


def _a2a_worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from imagine_b200.hotpath import exchange_pixel_shards, pixel_bounds
    K, n_local, m = 2, 2, 12
    mp = m // world
    lo, hi = shard_bounds(world * n_local, rank, world)
    # maps[k][r, pix] = 1000 k + 100 r_global + pix, written destination-major like the LOS kernel
    send = torch.empty((world, K, n_local, mp), dtype=torch.float64)
    for d in range(world):
        for k in range(K):
            for i in range(n_local):
                send[d, k, i] = 1000 * k + 100 * (lo + i) + torch.arange(d * mp, (d + 1) * mp)
    recv = exchange_pixel_shards(send)
    plo, phi = pixel_bounds(m, rank, world)
    queue.put((rank, recv.numpy(), plo, phi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pixel_shard_exchange():
    """All-to-all of the destination-major LOS output: every rank ends with all realisations of its pixels."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, recv, plo, phi in results:
        K, n_local = 2, 2
        full = recv.transpose(1, 0, 2, 3).reshape(K, 2 * n_local, phi - plo)     # [K][N, m_local]
        for k in range(K):
            for r in range(2 * n_local):
                assert np.array_equal(full[k, r], 1000 * k + 100 * r + np.arange(plo, phi))


def test_library_communicator_is_only_tried_where_it_applies(monkeypatch):
    """hotpath.ensure_library_comm: single rank, a sub-group or IMB_LIBRARY_COLLECTIVES=0 never touch NCCL (and
    never start a collective): torch.distributed stays the transport."""
    from imagine_b200 import hotpath
    assert hotpath.ensure_library_comm(0, 1) is False
    assert hotpath.ensure_library_comm(0, 2, group=object()) is False
    monkeypatch.setenv('IMB_LIBRARY_COLLECTIVES', '0')
    assert hotpath.ensure_library_comm(0, 2) is False
    assert hotpath._LIBRARY_COMM['tried'] is False            # none of the above counted as an attempt
