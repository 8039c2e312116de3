#!/usr/bin/env python
"""Likelihood-stage probe: K random [N, m] map blocks on the device, the batched Gram route alone.

  python tools/like_probe.py [--K 5] [--N 64] [--m 49152] [--reps 50] [--flush]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--K', type=int, default=5)
    ap.add_argument('--N', type=int, default=64)
    ap.add_argument('--m', type=int, default=49152)
    ap.add_argument('--reps', type=int, default=50)
    ap.add_argument('--flush', action='store_true', help='evict L2 between calls (the maps arrive from the LOS kernel)')
    args = ap.parse_args()
    import torch
    from imagine_b200 import _lib, engine
    torch.cuda.set_device(0)
    _lib.init(0)
    g = torch.Generator(device='cuda').manual_seed(1)
    sims = [torch.randn(args.N, args.m, dtype=torch.float64, device='cuda', generator=g) + 3.0 for _ in range(args.K)]
    meas = [s[0] * 1.01 for s in sims]
    var = [torch.full((args.m,), 0.5, dtype=torch.float64, device='cuda') for _ in range(args.K)]
    out = torch.zeros(args.K, dtype=torch.float64, device='cuda')
    flags = torch.zeros(args.K, dtype=torch.int32, device='cuda')
    junk = torch.empty(64 << 20, dtype=torch.float64, device='cuda') if args.flush else None
    engine.loglike_ensemble_batch_async(sims, meas, var, out, flags)
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(args.reps):
        if junk is not None:
            junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.loglike_ensemble_batch_async(sims, meas, var, out, flags)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    print('like_ms %.4f  K %d N %d m %d flush %s  sum %.12g' % (total / args.reps, args.K, args.N, args.m,
                                                               args.flush, float(out.sum())))


if __name__ == '__main__':
    main()
