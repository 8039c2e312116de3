"""
GPU: several parameter points per launch (EnsembleEvaluator.evaluate_batch, imb_field_eval_lanes) -- the
lanes of the ray march are (point, realisation) pairs, because rays, cells and weights do not depend on
the parameter point (batched counterpart of the loop over cubes in pipelines/pipeline.py:976-979).
"""
import numpy as np
import pytest

from tests.test_gpu_multirank import BOX, DESCS, RES, _setup

pytestmark = pytest.mark.gpu


def _points(seeds, strengths):
    pts = []
    for by in strengths:
        d = list(DESCS)
        d[1] = (0, [0.0, by, 0.0], 1)
        d[2] = (1, [0.0, 1.0 + by], 2)
        pts.append((d, seeds))
    return pts


@pytest.fixture(scope='module')
def lib():
    import torch
    assert torch.cuda.is_available()
    from imagine_b200 import _lib
    _lib.init(0)
    return _lib


def test_lanes_of_a_shared_buffer_hold_the_same_field_bit_for_bit(lib):
    import torch
    from imagine_b200 import engine
    grid = engine.DeviceGrid(BOX, RES)
    seeds = np.arange(7, 15, dtype=np.uint64)
    bdescs = [d for d in DESCS if d[0] in (0, 1)]
    alone = engine.field_eval(grid, bdescs, seeds, n_real=8, b_stride=24)['magnetic_field']      # [1, ncell, 3, 8]
    wide = torch.full((1, grid.ncell, 3, 32), 7.0, dtype=torch.float64, device='cuda')
    for off in (0, 8, 24):
        engine.field_eval_lanes(grid, bdescs, seeds, 8, 24, wide, off)
        assert torch.equal(wide[..., off:off + 8], alone)
    assert torch.all(wide[..., 16:24] == 7.0)                   # lanes of other points are not touched
    with pytest.raises(ValueError):
        engine.field_eval_lanes(grid, bdescs, seeds, 8, 24, wide, 28)     # would run past the 32 lanes


@pytest.mark.parametrize('N', [8, 4, 16])
def test_batched_points_match_one_point_at_a_time(lib, N):
    """world = 1 with the pixel-sharded data flow (what every rank of a sharded run executes)."""
    ev, seeds, meas, var = _setup(0, 1, 'alltoall', N=N)
    pts = _points(seeds, (0.5, 1.0, 2.0, 0.25, 1.5))
    cap = ev.batch_capacity(pts)
    assert cap == (2 if N == 16 else 4)
    one = [ev.evaluate(d, s, meas, var) for d, s in pts]
    got = ev.evaluate_batch(pts, meas, var)
    for a, b in zip(got, one):
        # same realisations; the one-point march splits each ray into 4 (2) segments per warp: rounding only
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (a, b)
    # the value of a point does not depend on its lane slot or on its neighbours in the batch: bit-identical
    again = ev.evaluate_batch([pts[2], pts[0], pts[4], pts[1]], meas, var)
    assert again == [got[2], got[0], got[4], got[1]]


def test_batching_steps_aside_when_it_does_not_apply(lib):
    ev, seeds, meas, var = _setup(0, 1, 'alltoall', N=8)
    pts = _points(seeds, (0.5, 1.0))
    other_ne = list(pts[1][0])
    other_ne[0] = (3, [0.02, 3.0, 0.5], 0)                      # a different n_e model in the second point
    mixed = [pts[0], (other_ne, seeds)]
    assert ev.batch_capacity(mixed) == 0
    want = [ev.evaluate(d, s, meas, var) for d, s in mixed]
    assert ev.evaluate_batch(mixed, meas, var) == ev.evaluate_many(mixed, meas, var)
    assert np.allclose(ev.evaluate_batch(mixed, meas, var), want, rtol=1e-12)
    ev32, seeds32, meas32, var32 = _setup(0, 1, 'alltoall', N=32)
    assert ev32.batch_capacity() == 0                            # 32 local realisations already fill the warp


def test_single_gpu_evaluator_batches_whatever_its_exchange_mode(lib):
    ev, seeds, meas, var = _setup(0, 1, 'auto', N=16)            # world = 1: 'auto' is the all-gather layout
    assert ev.exchange == 'allgather' and ev.batch_capacity() == 2
    pts = _points(seeds, (0.5, 1.0, 2.0))
    one = [ev.evaluate(d, s, meas, var) for d, s in pts]
    got = ev.evaluate_batch(pts, meas, var)
    for a, b in zip(got, one):
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (a, b)
