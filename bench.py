#!/usr/bin/env python
"""
bench.py -- ensemble log-likelihood evaluations per second (BASELINE.json metric).

A "step" is ONE full ensemble log-likelihood evaluation at one parameter point:
  fields (N realisations on the grid) -> LOS maps (K observables x N x 12 Nside^2)
  -> sum over observables of the EnsembleLikelihood term (OAS covariance).
Default workload = BASELINE.json configs[2] ("c3"): 128^3 grid, Nside 64, 64 realisations,
RM + DM + synchrotron I/Q/U at 23 GHz, OAS covariance, diagonal measurement covariance.
With N GPUs the ensemble is sharded (N/P realisations per rank, strong scaling) and the maps are
exchanged with one NCCL all-gather per observable.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same metric
through the plugin API with host buffers (H2D of seeds/measurements, D2H of the result inside
the timed region).  `--impl reference` times the CPU path (oracle/cpu_path.py) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: grid G, nside, realisations, observables
    'c2': dict(G=64, nside=32, n_real=16, quantities=('fd',)),
    'c3': dict(G=128, nside=64, n_real=64, quantities=('fd', 'dm', 'I', 'Q', 'U')),
    'c4': dict(G=256, nside=128, n_real=256, quantities=('fd', 'dm', 'I', 'Q', 'U')),
}
BOX = [[-20.0, 20.0], [-20.0, 20.0], [-4.0, 4.0]]          # tests/data/test_ham.xml:38-43
LOS_KW = dict(observer=(-8.3, 0.0, 0.006), dr=0.1, r_max=30.0, nu_ghz=23.0, alpha=3.0,
              beta=0.0, theta=0.0, e0_gev=20.6)
# field kinds (include/imagine_b200.h)
CONST_B, GAUSS_B, CONST_NE, EXP_NE, COS_NE, RANDOM_NE, CRE_ANALYTIC, CRE_CONST = range(8)


def field_descs(quantities):
    descs = [(EXP_NE, [0.01, 3.0, 0.1], 0), (CONST_B, [0.0, 1.0, 0.0], 1), (GAUSS_B, [0.0, 2.0], 2)]
    if any(q in quantities for q in 'IQU'):
        descs.append((CRE_ANALYTIC, [0.0217, 5.0, 1.0, 8.3], 3))
    return descs


def ensemble_seeds(n):
    # pipelines/pipeline.py:805-813 with master_seed = 1 (tools/config.py:54)
    np.random.seed(1)
    return np.random.randint(1, 1431655765, n).astype(np.uint64)


def los_profile_summary(workload, world, n_local):
    """dram bytes of ONE launch of the committed LOS kernel from its `ncu --set full` capture
    (profiles/r02_ncu_los_summary.json, written by tools/ncu_summary.py); None when no capture covers
    this configuration."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_ncu_los_summary.json')) as f:
            prof = json.load(f)
    except Exception:
        return None
    for entry in prof.get('captures', []):
        if entry.get('workload') == workload and entry.get('n_gpus') == world and entry.get('n_local') == n_local:
            return entry
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference(workload):
    """The fixed bounded CPU sample (oracle/cpu_path.py): same realisations and pixels on every call."""
    from oracle import cpu_path
    w = WORKLOADS[workload]
    return cpu_path.time_sample(BOX, [w['G']] * 3, w['nside'], field_descs(w['quantities']),
                                ensemble_seeds(w['n_real']), LOS_KW, w['n_real'], quantities=w['quantities'])


def cpu_baseline_object(res):
    return {'value': 1.0 / res['seconds_per_eval'], 'unit': 'evals/s', 'cores': res['cores'], 'kind': 'port',
            'kinds': res['kinds'], 'sample': res['sample'], 'stages_s': res['stages'],
            'measured_sample_s': res['measured_sample_s'], 'measured_s': res['measured'],
            'scale_factor': res['scale_factor'], 'extrapolated': res['extrapolated']}


def run_reference(args):
    """--impl reference: the CPU path on the host cores, same metric/config, rank 0 only.  Every step
    times the same fixed sample (about 10 s on 16 cores); the loop stops after --steps samples or ~150 s
    of wall clock, whichever comes first, and reports how many were timed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    times, last = [], None
    t_run = time.perf_counter()
    for i in range(args.warmup + args.steps):
        last = cpu_reference(args.workload)
        if i >= args.warmup or args.warmup + args.steps <= 1:
            times.append(last['seconds_per_eval'])
        if time.perf_counter() - t_run > 150 and times:          # bounded run, never hang the driver
            break
    sec = float(np.mean(times))
    value = 1.0 / sec
    base = cpu_baseline_object(last)
    base['value'] = value
    line = {
        'impl': 'reference', 'metric': 'ensemble log-likelihood evals/sec', 'value': value,
        'unit': 'evals/s', 'n_gpus': args.gpus, 'steps': len(times), 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.workload, w, 1),
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'timing_note': ('ms_per_step is the EXTRAPOLATED time of one full evaluation; each timed step ran the '
                        'bounded sample (cpu_baseline.measured_sample_s seconds of wall clock)'),
        'run_wall_s': time.perf_counter() - t_run,
    }
    print(json.dumps(line))


def workload_config(name, w, world):
    return {'workload': '%s: %d^3 grid, Nside %d (m=%d), %d realisations, observables %s @23 GHz, '
                        'OAS covariance, diagonal measurement covariance'
                        % (name, w['G'], w['nside'], 12 * w['nside'] ** 2, w['n_real'],
                           '+'.join(w['quantities'])),
            'grid': [w['G']] * 3, 'nside': w['nside'], 'ensemble': w['n_real'],
            'observables': list(w['quantities']), 'los_step_kpc': LOS_KW['dr'],
            'los_steps': int(np.ceil(LOS_KW['r_max'] / LOS_KW['dr'] - 1e-9)),
            'parallelism': 'ensemble-sharded x%d' % world, 'l2': 'inputs larger than L2 (B grids)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--points', type=int, default=8, help='parameter points per pipelined flush (extra key)')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'peer', 'alltoall', 'allgather'])
    ap.add_argument('--profile-e2e', action='store_true', help='cProfile the plugin-API loop (stderr)')
    ap.add_argument('--b-stride', default='auto', help="field layout: 'auto' | 24 | 48 | 96 (A/B runs)")
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from imagine_b200 import _lib, engine
    from imagine_b200.hotpath import EnsembleEvaluator

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; there is no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    _lib.init(local_rank)

    w = WORKLOADS[args.workload]
    quantities = w['quantities']
    N, G, nside = w['n_real'], w['G'], w['nside']
    m = 12 * nside * nside
    descs = field_descs(quantities)
    seeds = ensemble_seeds(N)
    settings = engine.LosSettings(**LOS_KW)
    b_stride = args.b_stride if args.b_stride == 'auto' else int(args.b_stride)
    ev = EnsembleEvaluator(BOX, [G] * 3, nside, settings, quantities, n_real=N, rank=rank, world=world,
                           exchange=args.exchange, b_stride=b_stride)

    # ---- synthetic measurements: one realisation from an independent seed + 1% noise --------
    mock = EnsembleEvaluator(BOX, [G] * 3, nside, settings, quantities, n_real=1, exchange='allgather')
    mock.simulate(descs, np.array([987654321], dtype=np.uint64))
    gen = torch.Generator(device='cuda').manual_seed(2026)
    meas_h, var_h = {}, {}
    for q in quantities:
        row = mock.local_maps[q][0]
        sigma = 0.01 * row.abs().mean()
        meas = row + sigma * torch.randn(m, dtype=torch.float64, device='cuda', generator=gen)
        meas_h[q] = meas.cpu().pin_memory()
        var_h[q] = torch.full((m,), float(sigma) ** 2, dtype=torch.float64).pin_memory()
    del mock
    meas_d = {q: meas_h[q].cuda() for q in quantities}
    var_d = {q: var_h[q].cuda() for q in quantities}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------------------------------
    def step(record=None):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record is not None else None
        if e: e[0].record()
        f = ev.fields(descs, seeds)
        if e: e[1].record()
        ev.los(f)
        if e: e[2].record()
        maps = ev.gather()
        if e: e[3].record()
        val = ev.loglike(maps, meas_d, var_d)
        if e:
            e[4].record()
            record.append(e)
        return val

    for _ in range(args.warmup):
        value_check = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    records = []
    launches0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        loglike = step(records)
    t_end.record()
    barrier()
    launches = _lib.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    stage = np.array([[r[i].elapsed_time(r[i + 1]) for i in range(4)] for r in records]).mean(0)

    # ---- pipelined multi-point evaluation: same step, `--points` of them enqueued back
    # to back with ONE host read-back; reported as an extra key, the headline `value` stays per-call
    pipelined = None
    if args.points > 1:
        pts = [(descs, seeds)] * args.points
        ev.evaluate_many(pts, meas_d, var_d)
        barrier()
        n_flush = max(1, args.steps // args.points)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(n_flush):
            many = ev.evaluate_many(pts, meas_d, var_d)
        p1.record()
        barrier()
        pip_ms = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(pip_ms, op=dist.ReduceOp.MAX)
        pipelined = {'points_per_flush': args.points, 'flushes': n_flush,
                     'value': n_flush * args.points / (float(pip_ms[0]) * 1e-3), 'unit': 'evals/s',
                     'loglike': many[-1]}

    # ---- several parameter points per LAUNCH: the lanes of one ray march are (point,
    # realisation) pairs; applies on the pixel-sharded path with <= 16 local realisations
    batched = None
    if args.points > 1 and ev.batch_capacity([(descs, seeds)] * 2) >= 2:
        pts = [(descs, seeds)] * args.points
        ev.evaluate_batch(pts, meas_d, var_d)
        barrier()
        n_flush = max(1, args.steps // args.points)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(n_flush):
            bvals = ev.evaluate_batch(pts, meas_d, var_d)
        b1.record()
        barrier()
        bat_ms = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(bat_ms, op=dist.ReduceOp.MAX)
        batched = {'points_per_launch': ev.batch_capacity(pts), 'points_per_flush': args.points, 'flushes': n_flush,
                   'value': n_flush * args.points / (float(bat_ms[0]) * 1e-3), 'unit': 'evals/s',
                   'loglike': bvals[-1],
                   'mode': 'lanes of one ray march = (parameter point, realisation); one all-to-all and one '
                           'chain of Gram phases per launch'}

    # ---- point-parallel mode (the reference's `_mpi_likelihood` decomposition): every rank
    # evaluates whole parameter points on its own, one small all-gather of the scalars per batch
    replicated = None
    if world > 1 and args.points > 0:
        pts = [(descs, seeds)] * (world * 2)
        ev.evaluate_points(pts, meas_d, var_d)
        barrier()
        n_batch = max(1, args.steps // 4)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(n_batch):
            rvals = ev.evaluate_points(pts, meas_d, var_d)
        r1.record()
        barrier()
        rep_ms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(rep_ms, op=dist.ReduceOp.MAX)
        replicated = {'points_per_batch': len(pts), 'batches': n_batch,
                      'value': n_batch * len(pts) / (float(rep_ms[0]) * 1e-3), 'unit': 'evals/s',
                      'loglike': rvals[-1], 'mode': 'one parameter point per GPU, no data-path collective'}

    # ---- end-to-end steps: the plugin API a Pipeline calls (pipelines/pipeline.py:827-852, 905-943):
    # new Field objects per evaluation, Simulator.__call__, Likelihood.__call__; measurements and
    # variances start in HOST memory every step (device caches dropped), result is a host float
    import imagine_b200.api as api
    u = api.units
    api_grid = api.UniformGrid(box=[[b[0], b[1]] * u.kpc for b in BOX], resolution=[G] * 3)
    key_of = {'fd': ('fd', None, nside, None), 'dm': ('dm', None, nside, None),
              'I': ('sync', 23.0, nside, 'I'), 'Q': ('sync', 23.0, nside, 'Q'), 'U': ('sync', 23.0, nside, 'U')}
    unit_of = {'fd': u.rad / u.m ** 2, 'dm': u.pc / u.cm ** 3, 'I': u.K, 'Q': u.K, 'U': u.K}
    api_meas = api.Measurements()
    for q in quantities:
        api_meas.append(name=key_of[q], data=meas_h[q].numpy().reshape(1, -1) * unit_of[q],
                        cov_data=var_h[q].numpy() * unit_of[q] ** 2, otype='HEALPix')
    api_sim = api.B200Simulator(api_meas, observer=LOS_KW['observer'], radial_step=LOS_KW['dr'],
                                radial_max=LOS_KW['r_max'], cre_alpha=LOS_KW['alpha'], cre_E0=LOS_KW['e0_gev'],
                                rank=rank, world=world, b_stride=b_stride)
    api_lik = api.B200EnsembleLikelihood(api_meas)
    seed_list = [int(v) for v in seeds]
    h2d = sum(meas_h[q].numel() * 8 + var_h[q].numel() * 8 for q in quantities) + len(seed_list) * 8

    def e2e_step():
        fields = [api.ExponentialThermalElectrons(api_grid, parameters={
                      'central_density': 0.01 * u.cm ** -3, 'scale_radius': 3.0 * u.kpc,
                      'scale_height': 0.1 * u.kpc}, ensemble_size=N),
                  api.ConstantMagneticField(api_grid, parameters={
                      'Bx': 0 * u.microgauss, 'By': 1 * u.microgauss, 'Bz': 0 * u.microgauss}, ensemble_size=N),
                  api.NaiveGaussianMagneticField(api_grid, parameters={
                      'a0': 0 * u.microgauss, 'b0': 2 * u.microgauss}, ensemble_size=N, ensemble_seeds=seed_list)]
        if any(q in quantities for q in 'IQU'):
            fields.append(api.AnalyticCosmicRayElectrons(api_grid, parameters={
                'j0': 0.0217, 'r0': 5.0 * u.kpc, 'z0': 1.0 * u.kpc, 'r_sun': 8.3 * u.kpc}, ensemble_size=N))
        sims = api_sim(fields)
        api_lik.upload_measurements()     # measurements/variances re-uploaded from pinned host memory
        return float(api_lik(sims))

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    prof = None
    if args.profile_e2e and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    for _ in range(args.steps):
        e2e_val = e2e_step()
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(45)
        pstats.Stats(prof, stream=sys.stderr).sort_stats('tottime').print_stats(25)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    e2e_wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks -------------------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms, e2e_wall_ms, stage[0], stage[1], stage[2], stage[3]],
                     dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_wall_ms = float(t[0]), float(t[1]), float(t[2])
    stage = [float(v) for v in t[3:7]]

    if rank == 0:
        n_loc = N // world
        s_bytes = G ** 3 * 8
        n_det = 2 if any(q in quantities for q in 'IQU') else 1       # n_e (+ CRE) scalar grids
        los_bytes = n_loc * 3 * s_bytes + n_det * s_bytes + n_loc * len(quantities) * m * 8
        field_bytes = n_loc * 3 * s_bytes + n_det * s_bytes
        peak, peak_kind = measured_peak_gbs()
        prof = los_profile_summary(args.workload, world, n_loc)
        los_gbs = los_bytes / (stage[1] * 1e-3) / 1e9
        n_steps = settings.n_steps
        # exact in-box sample count (geometry only)
        d = ev.dirs
        r = (torch.arange(n_steps, dtype=torch.float64, device='cuda') + 0.5) * settings.dr
        obs = torch.tensor(settings.observer, dtype=torch.float64, device='cuda')
        inbox = 0
        for s in range(0, m, 8192):
            pos = obs + r[None, :, None] * d[s:s + 8192, None, :]
            lo = torch.tensor([b[0] for b in BOX], dtype=torch.float64, device='cuda')
            hi = torch.tensor([b[1] for b in BOX], dtype=torch.float64, device='cuda')
            inbox += int(((pos >= lo) & (pos <= hi)).all(-1).sum())
        line = {
            'metric': 'ensemble log-likelihood evals/sec', 'value': args.steps / (total_ms * 1e-3),
            'unit': 'evals/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.workload, w, world),
            'e2e': {'value': args.steps / (max(e2e_ms, e2e_wall_ms) * 1e-3), 'unit': 'evals/s',
                    'h2d_bytes_per_step': int(h2d),
                    # the one read-back of an evaluation: K LikeScalars records (11 doubles + 2 ints each)
                    'd2h_bytes_per_step': 96 * len(quantities),
                    'ms_per_step': max(e2e_ms, e2e_wall_ms) / args.steps},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'kernel': 'los_rl_kernel', 'bound': 'hbm', 'achieved': los_gbs, 'peak': peak,
                         'unit': 'GB/s', 'frac': los_gbs / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the committed kernel
                         'traffic': prof['dram_bytes'] if prof else None,
                         'traffic_source': (prof['source'] if prof else
                                            'no ncu capture for this configuration under profiles/'),
                         'limiter': ('not HBM at this size: the B grid leaves HBM about once (see traffic); the '
                                     'march is bound by FP64 issue (60 FP64 instructions per sample) and the '
                                     'latency of the vertex loads at 20 warps / SM; L1 global-load sectors '
                                     '0.86e9 per launch after the cell-resident rewrite (2.04e9 before)'),
                         'peak_source': peak_kind + ' (MEASURED_PEAKS.json hbm_gbs)',
                         'algorithmic_bytes_per_launch': int(los_bytes),
                         'ms_per_launch': stage[1]},
            'stages_ms': {'fields': stage[0], 'los': stage[1], 'gather': stage[2], 'likelihood': stage[3]},
            'field_kernel': {'achieved': field_bytes / (stage[0] * 1e-3) / 1e9, 'unit': 'GB/s',
                             'frac': field_bytes / (stage[0] * 1e-3) / 1e9 / peak,
                             'algorithmic_bytes': int(field_bytes)},
            'los_samples_per_s': {'nominal': n_loc * m * n_steps / (stage[1] * 1e-3),
                                  'in_box': n_loc * inbox / (stage[1] * 1e-3),
                                  'in_box_per_realisation': inbox},
            'loglike': float(loglike),
            'exchange': {'mode': ev.exchange, 'note': ev.exchange_note},
            'pipelined': pipelined,
            'batched': batched,
            'point_parallel': replicated,
        }
        if not args.no_cpu_baseline and world == 1:          # the CPU leg is timed at N = 1 only
            try:
                line['cpu_baseline'] = cpu_baseline_object(cpu_reference(args.workload))
            except Exception as exc:                                      # never lose the GPU line
                line['cpu_baseline'] = {'value': None, 'unit': 'evals/s', 'cores': 0, 'kind': 'port',
                                        'sample': 'failed: %r' % (exc,)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
