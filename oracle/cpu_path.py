"""
CPU timing of the hot path with the oracle (test infrastructure / bench baseline only).

Used by `bench.py` for the `cpu_baseline` object and for `--impl reference`.  What runs:
  fields      NumPy restatement of the reference compute_field bodies with the reference's own
              MT19937 stream (oracle/fields.py, rng='reference')
  LOS         oracle/los.py  -- OUR oracle, not reference code: the reference delegates LOS
              integration to the external hammurabiX binary (not vendored)
  likelihood  oracle/likelihood.py Gram-route port; the reference's dense LU path needs three
              live m x m matrices (58 GB at Nside 64) and is reported as infeasible there
A bounded SAMPLE of the workload is timed (a few realisations, a pixel subset) on all host
cores (fork pool over realisations x pixel chunks) and scaled linearly to one full evaluation.
"""
import os
import time
from multiprocessing import get_context

import numpy as np

from . import fields as F
from . import healpix as H
from . import likelihood as LK
from . import los as L

_G = {}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _los_task(args):
    r, s, e = args
    out = L.integrate(_G['grid'], _G['cfg'], _G['dirs'][s:e], _G['B'][r:r + 1], _G['ne'], _G['cre'])
    return r, s, e, {k: v[0] for k, v in out.items()}


def _field_task(r):
    f = F.evaluate_fields(_G['grid'], _G['descs'], [_G['seeds'][r]], rng='reference')
    return r, f['magnetic_field'][0]


def time_sample(box, resolution, nside, descs, seeds, cfg_kw, n_real_full, n_real_sample=2,
                pixel_fraction=0.125, cores=None, quantities=('fd', 'dm', 'I', 'Q', 'U')):
    """
    Times a bounded sample and extrapolates to one full ensemble evaluation.
    Returns dict(seconds_per_eval, stages, cores, sample).
    """
    cores = cores or host_cores()
    grid = F.Grid(box, resolution)
    cfg = L.LosConfig(**cfg_kw)
    dirs_all = H.pix2vec_ring(nside)
    m = dirs_all.shape[0]
    m_sub = max(cores * 64, int(m * pixel_fraction))
    m_sub = min(m, m_sub)
    pix = np.linspace(0, m - 1, m_sub).astype(np.int64)      # spread over all latitudes
    dirs = dirs_all[pix]
    _G.update(grid=grid, cfg=cfg, dirs=dirs, descs=descs, seeds=list(seeds))

    ctx = get_context('fork')
    # stage 1: fields (stochastic per realisation in a pool; deterministic once)
    t0 = time.perf_counter()
    det = F.evaluate_fields(grid, [d for d in descs if d[0] not in F.STOCHASTIC_KINDS], [0])
    t_det = time.perf_counter() - t0
    t0 = time.perf_counter()
    with ctx.Pool(min(cores, n_real_sample)) as pool:
        res = pool.map(_field_task, range(n_real_sample))
    t_fields_sample = time.perf_counter() - t0
    B = np.stack([b for _, b in sorted(res, key=lambda x: x[0])])
    _G.update(B=B, ne=det.get('thermal_electron_density'), cre=det.get('cosmic_ray_electron_density'))

    # stage 2: LOS over realisations x pixel chunks
    chunk = max(64, m_sub // (4 * cores))
    tasks = [(r, s, min(m_sub, s + chunk)) for r in range(n_real_sample) for s in range(0, m_sub, chunk)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        parts = pool.map(_los_task, tasks)
    t_los_sample = time.perf_counter() - t0
    maps = {q: np.zeros((n_real_sample, m_sub)) for q in quantities}
    for r, s, e, out in parts:
        for q in quantities:
            maps[q][r, s:e] = out[q]

    # stage 3: likelihood at FULL ensemble size on synthetic maps of the sample's scale
    rng = np.random.default_rng(0)
    t_like = 0.0
    for q in quantities:
        scale = np.abs(maps[q]).mean() + 1e-300
        sims = maps[q].mean(0) + 0.3 * scale * rng.normal(size=(n_real_full, m_sub))
        meas = maps[q][0] + 0.01 * scale * rng.normal(size=m_sub)
        var = np.full(m_sub, (0.01 * scale) ** 2)
        t0 = time.perf_counter()
        LK.ensemble_loglike_gram(sims, meas, var)
        t_like += time.perf_counter() - t0

    # linear extrapolation to the full workload
    waves = int(np.ceil(n_real_full / min(cores, n_real_full)))
    fields_full = t_det + t_fields_sample * (waves / max(1, int(np.ceil(n_real_sample / min(cores, n_real_sample)))))
    los_full = t_los_sample * (m / m_sub) * (n_real_full / n_real_sample)
    like_full = t_like * (m / m_sub)
    total = fields_full + los_full + like_full
    return dict(seconds_per_eval=total, cores=cores,
                stages=dict(fields_s=fields_full, los_s=los_full, likelihood_s=like_full),
                sample='%d of %d realisations x %d of %d pixels timed on %d cores, scaled linearly; '
                       'fields: NumPy+MT19937 restatement; LOS: our oracle (reference delegates to '
                       'hammurabiX); likelihood: Gram-route port (reference dense LU infeasible at this m)'
                       % (n_real_sample, n_real_full, m_sub, m, cores))
