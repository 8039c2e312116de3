"""
CPU timing of the hot path (test infrastructure / bench baseline only; bench.py's `cpu_baseline`
object and `--impl reference`).  What runs on the host cores:

  fields      the REFERENCE's own Field classes, unmodified, from baseline/_ref (imagine_b200.reference
              makes the package importable): ExponentialThermalElectrons, ConstantMagneticField,
              NaiveGaussianMagneticField (its own MT19937 / scipy.stats stream), summed like
              Simulator.prepare_fields does.  The analytic CRE has no reference class (base_fields.py:
              104-113 is a stub) and comes from oracle/fields.py.  When the reference cannot be imported
              the NumPy restatement of oracle/fields.py is timed instead and the leg is labelled "port".
  LOS         oracle/los.py -- OUR oracle, not reference code: the reference delegates LOS integration
              to the external hammurabiX binary (not vendored, not installed)
  likelihood  oracle/likelihood.py Gram-route port for the metric (the reference's dense LU path needs
              three live m x m matrices, 58 GB at Nside 64, and O(m^3) work).  Next to it the
              reference's OWN EnsembleLikelihoodDiagonal is timed at full size, and its own
              EnsembleLikelihood (dense LU) where it fits (m <= 3072), as separate reported figures.

A FIXED, bounded sample of the workload is timed -- the same realisations and the same pixels on every
call, so repeated runs are comparable -- on all host cores (fork pool over realisations x pixel chunks)
and scaled linearly to one full evaluation; the measured seconds and the scale factors are returned.
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

SAMPLE_REALISATIONS = 16        # at most; the first ones of the ensemble
SAMPLE_PIXELS = 6144            # at most; evenly spread over all latitudes


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _los_task(args):
    r, s, e = args
    out = L.integrate(_G['grid'], _G['cfg'], _G['dirs'][s:e], _G['B'][r:r + 1], _G['ne'], _G['cre'])
    return r, s, e, {k: v[0] for k, v in out.items()}


def _field_task_port(r):
    f = F.evaluate_fields(_G['grid'], _G['descs'], [_G['seeds'][r]], rng='reference')
    return r, f['magnetic_field'][0]


def _reference_fields():
    """(imagine.fields module, units) or None."""
    try:
        from imagine_b200 import reference, units
        reference.ensure_imagine()
        import imagine.fields as RF
        return RF, units
    except Exception:
        return None


def _ref_field_objects(RF, u, grid, descs, seed):
    """Reference Field objects for the bench's descriptor list (one realisation, its own seed)."""
    rgrid = RF.UniformGrid(box=[[lo * u.kpc, hi * u.kpc] for lo, hi in grid.box], resolution=list(grid.shape))
    objs = []
    for kind, p, _stream in [(d[0], d[1], d[2]) for d in descs]:
        if kind == F.CONST_B:
            objs.append(RF.ConstantMagneticField(rgrid, parameters={
                'Bx': p[0] * u.microgauss, 'By': p[1] * u.microgauss, 'Bz': p[2] * u.microgauss}))
        elif kind == F.GAUSS_B:
            objs.append(RF.NaiveGaussianMagneticField(rgrid, parameters={
                'a0': p[0] * u.microgauss, 'b0': p[1] * u.microgauss}, ensemble_size=1, ensemble_seeds=[int(seed)]))
        elif kind == F.EXP_NE:
            objs.append(RF.ExponentialThermalElectrons(rgrid, parameters={
                'central_density': p[0] * u.cm ** -3, 'scale_radius': p[1] * u.kpc,
                'scale_height': p[2] * u.kpc}))
        elif kind == F.CONST_NE:
            objs.append(RF.ConstantThermalElectrons(rgrid, parameters={'ne': p[0] * u.cm ** -3}))
        elif kind in (F.CRE_ANALYTIC, F.CRE_CONST):
            continue                                    # no reference class
        else:
            raise KeyError(kind)
    return objs


def _field_task_ref(r):
    RF, u = _G['ref']
    total = None
    for f in _ref_field_objects(RF, u, _G['grid'], _G['descs'], _G['seeds'][r]):
        if f.type != 'magnetic_field':
            continue
        data = f.get_data(0)
        total = data if total is None else total + data          # simulator.py:166-178
    return r, np.asarray(total.to_value(u.microgauss))


def _reference_likelihood_times(maps_full, meas, var, quantities, m_sub):
    """Seconds for the reference's own likelihood classes on the synthetic sample maps, or None."""
    try:
        from imagine_b200 import reference
        reference.ensure_imagine()
        import imagine.likelihoods as RL
        import imagine.observables as RO
    except Exception:
        return None
    out = {}
    M, S = RO.Measurements(), RO.Simulations()
    for q in quantities:
        key = (q, None, m_sub, None)
        M.append(name=key, data=meas[q].reshape(1, -1), cov_data=var[q], otype='plain')
        S.append(name=key, data=maps_full[q], otype='plain')
    t0 = time.perf_counter()
    RL.EnsembleLikelihoodDiagonal(M)(S)
    out['diagonal_sample_s'] = time.perf_counter() - t0
    if m_sub <= 3072:
        t0 = time.perf_counter()
        RL.EnsembleLikelihood(M)(S)
        out['dense_sample_s'] = time.perf_counter() - t0
    return out


def time_sample(box, resolution, nside, descs, seeds, cfg_kw, n_real_full, cores=None,
                quantities=('fd', 'dm', 'I', 'Q', 'U'), n_real_sample=None, n_pixels=None):
    """
    Times the fixed bounded sample and extrapolates to one full ensemble evaluation.
    Returns dict(seconds_per_eval, measured_sample_s, scale, stages, kinds, cores, sample, ...).
    """
    cores = cores or host_cores()
    grid = F.Grid(box, resolution)
    cfg = L.LosConfig(**cfg_kw)
    dirs_all = H.pix2vec_ring(nside)
    m = dirs_all.shape[0]
    n_s = min(n_real_full, n_real_sample or SAMPLE_REALISATIONS)
    m_sub = min(m, n_pixels or SAMPLE_PIXELS)
    pix = np.linspace(0, m - 1, m_sub).astype(np.int64)
    dirs = dirs_all[pix]
    ref = _reference_fields()
    _G.update(grid=grid, cfg=cfg, dirs=dirs, descs=descs, seeds=list(seeds), ref=ref)
    ctx = get_context('fork')
    t_all = time.perf_counter()

    # stage 1: fields.  Deterministic scalar grids once; the stochastic magnetic field per realisation.
    t0 = time.perf_counter()
    det = F.evaluate_fields(grid, [d for d in descs if d[0] not in F.STOCHASTIC_KINDS], [0])
    if ref is not None:                                  # the n_e grid by the reference's own class
        RF, u = ref
        for f in _ref_field_objects(RF, u, grid, descs, 0):
            if f.type == 'thermal_electron_density':
                det['thermal_electron_density'] = np.asarray(f.get_data(0).to_value(u.cm ** -3))[None]
    t_det = time.perf_counter() - t0
    t0 = time.perf_counter()
    with ctx.Pool(min(cores, n_s)) as pool:
        res = pool.map(_field_task_ref if ref is not None else _field_task_port, range(n_s))
    t_fields_sample = time.perf_counter() - t0
    B = np.stack([b for _, b in sorted(res, key=lambda x: x[0])])
    _G.update(B=B, ne=det.get('thermal_electron_density'), cre=det.get('cosmic_ray_electron_density'))

    # stage 2: LOS over realisations x pixel chunks
    chunk = max(64, m_sub // (4 * cores))
    tasks = [(r, s, min(m_sub, s + chunk)) for r in range(n_s) for s in range(0, m_sub, chunk)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        parts = pool.map(_los_task, tasks)
    t_los_sample = time.perf_counter() - t0
    maps = {q: np.zeros((n_s, m_sub)) for q in quantities}
    for r, s, e, out in parts:
        for q in quantities:
            maps[q][r, s:e] = out[q]

    # stage 3: likelihood at FULL ensemble size on synthetic maps of the sample's scale
    rng = np.random.default_rng(0)
    t_like = 0.0
    sims_full, meas, var = {}, {}, {}
    for q in quantities:
        scale = np.abs(maps[q]).mean() + 1e-300
        sims_full[q] = maps[q].mean(0) + 0.3 * scale * rng.normal(size=(n_real_full, m_sub))
        meas[q] = maps[q][0] + 0.01 * scale * rng.normal(size=m_sub)
        var[q] = np.full(m_sub, (0.01 * scale) ** 2)
        t0 = time.perf_counter()
        LK.ensemble_loglike_gram(sims_full[q], meas[q], var[q])
        t_like += time.perf_counter() - t0
    ref_like = _reference_likelihood_times(sims_full, meas, var, quantities, m_sub)
    measured = time.perf_counter() - t_all

    # linear extrapolation to the full workload
    waves_full = int(np.ceil(n_real_full / min(cores, n_real_full)))
    waves_sample = max(1, int(np.ceil(n_s / min(cores, n_s))))
    scale = dict(fields=waves_full / waves_sample, los=(m / m_sub) * (n_real_full / n_s), likelihood=m / m_sub)
    fields_full = t_det + t_fields_sample * scale['fields']
    los_full = t_los_sample * scale['los']
    like_full = t_like * scale['likelihood']
    total = fields_full + los_full + like_full
    stages = dict(fields_s=fields_full, los_s=los_full, likelihood_s=like_full)
    if ref_like:
        stages['reference_EnsembleLikelihoodDiagonal_s'] = ref_like['diagonal_sample_s'] * scale['likelihood']
        if 'dense_sample_s' in ref_like:
            stages['reference_EnsembleLikelihood_dense_sample_s'] = ref_like['dense_sample_s']
    kinds = dict(fields='reference (baseline/_ref classes, MT19937 stream)' if ref is not None
                 else 'port (oracle/fields.py restatement)',
                 los='port (oracle/los.py: our oracle; the reference delegates to hammurabiX)',
                 likelihood='port (Gram route; the reference dense LU is infeasible at this m)')
    return dict(seconds_per_eval=total, cores=cores, stages=stages, kinds=kinds,
                measured_sample_s=measured,
                measured=dict(fields_det_s=t_det, fields_sample_s=t_fields_sample, los_sample_s=t_los_sample,
                              likelihood_sample_s=t_like),
                scale_factor=scale, extrapolated=True,
                sample='%d of %d realisations x %d of %d pixels (fixed subset) timed on %d cores and scaled '
                       'linearly; fields: %s; LOS: our oracle (the reference delegates to hammurabiX); '
                       'likelihood: Gram-route port (reference dense LU infeasible at this m)'
                       % (n_s, n_real_full, m_sub, m, cores,
                          "reference's own classes" if ref is not None else 'NumPy restatement'))
