#!/usr/bin/env python
"""LOS-stage probe: fields once, then the ray-march kernel alone (timing loop or one launch for ncu).

  python tools/los_probe.py [--workload c3] [--n-real N] [--reps 20] [--once]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c3')
    ap.add_argument('--n-real', type=int, default=None)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--once', action='store_true', help='one warm-up + one launch (ncu captures)')
    ap.add_argument('--b-stride', default='auto')
    args = ap.parse_args()
    import torch
    from imagine_b200 import _lib, engine
    from imagine_b200.hotpath import EnsembleEvaluator
    torch.cuda.set_device(0)
    _lib.init(0)
    w = bench.WORKLOADS[args.workload]
    N = args.n_real or w['n_real']
    settings = engine.LosSettings(**bench.LOS_KW)
    bs = args.b_stride if args.b_stride == 'auto' else int(args.b_stride)
    ev = EnsembleEvaluator(bench.BOX, [w['G']] * 3, w['nside'], settings, w['quantities'], n_real=N,
                           exchange='allgather', b_stride=bs)
    descs = bench.field_descs(w['quantities'])
    f = ev.fields(descs, bench.ensemble_seeds(N))
    ev.los(f)
    torch.cuda.synchronize()
    reps = 1 if args.once else args.reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ev.los(f)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if not args.once:
        seeds = bench.ensemble_seeds(N)
        e0.record()
        for _ in range(reps):
            ev.fields(descs, seeds)
        e1.record()
        torch.cuda.synchronize()
        print('fields_ms %.4f' % (e0.elapsed_time(e1) / reps))
    chk = {q: float(ev.local_maps[q].double().abs().sum()) for q in w['quantities']}
    print('los_ms %.4f  b_stride %s  n_real %d  checksum %s' % (ms, ev.b_stride, N, chk))


if __name__ == '__main__':
    main()
