#!/usr/bin/env python
"""Dense-route probe at the C2 shape (Nside 32: m = 12 288 pixels, 16 realisations): the
EnsembleLikelihood with a DENSE m x m covariance (S = U^T U / N by FP64 DMMA + OAS + measurement
covariance, blocked Cholesky with the right-hand side riding along, log-determinant) through the
C ABI, next to two library yardsticks measured in the same process:

  * cuBLAS DGEMM (torch.matmul, float64)                  -> the FP64 rate this GPU gives a library GEMM
  * cuSOLVER potrf (torch.linalg.cholesky) on the same A  -> a library factorisation of the same matrix

  python tools/dense_probe.py [--m 12288] [--N 16] [--reps 5] [--json out.json] [--once]

--once runs one warm-up and one timed call (ncu captures).  The Gram route on the same input is timed
beside it (that is the route IMB_ROUTE_AUTO takes for a diagonal measurement covariance).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--m', type=int, default=12288)
    ap.add_argument('--N', type=int, default=16)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--json', default=None)
    ap.add_argument('--once', action='store_true')
    args = ap.parse_args()
    import torch
    from imagine_b200 import _lib, engine
    torch.cuda.set_device(0)
    _lib.init(0)
    m, N = args.m, args.N
    g = torch.Generator(device='cuda').manual_seed(7)
    sims = torch.randn(N, m, dtype=torch.float64, device='cuda', generator=g) + 2.0
    meas = sims[0] * 1.02
    var = torch.full((m,), 0.4, dtype=torch.float64, device='cuda')
    dense_cov = torch.diag(var)
    # a dense covariance with off-diagonal structure: nearest-neighbour correlation 0.2
    idx = torch.arange(m - 1, device='cuda')
    dense_cov[idx, idx + 1] = 0.08
    dense_cov[idx + 1, idx] = 0.08
    reps = 1 if args.once else args.reps

    run_dense = lambda: engine.loglike_ensemble(sims, meas, dense_cov, route='dense', return_info=True)
    run_dense()
    dense_ms, (val_dense, info) = timed(run_dense, reps)
    if args.once:
        print('dense_ms %.3f value %.12g info %s' % (dense_ms, val_dense, info))
        return
    # same matrix, diagonal part only, both routes: the two routes must agree
    run_d = lambda: engine.loglike_ensemble(sims, meas, var, route='dense')
    run_g = lambda: engine.loglike_ensemble(sims, meas, var, route='gram')
    run_d(); run_g()
    dvar_ms, val_d = timed(run_d, reps)
    gram_ms, val_g = timed(run_g, 20)

    # yardsticks
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda', generator=g)
    b = torch.randn(n, n, dtype=torch.float64, device='cuda', generator=g)
    torch.matmul(a, b)
    mm_ms, _ = timed(lambda: torch.matmul(a, b), 5)
    dgemm_tflops = 2.0 * n ** 3 / (mm_ms * 1e-3) / 1e12
    del a, b
    u = sims - sims.mean(0, keepdim=True)
    A = dense_cov + u.T @ u / N * 0.7 + 0.1 * torch.eye(m, dtype=torch.float64, device='cuda')
    torch.linalg.cholesky(A)
    potrf_ms, Lc = timed(lambda: torch.linalg.cholesky(A), 3)
    syrk_ms, _ = timed(lambda: u.T @ u, 5)

    flops_chol = m ** 3 / 3.0
    line = {
        'probe': 'dense route, EnsembleLikelihood with a dense covariance', 'm': m, 'N': N, 'reps': reps,
        'dense_route_ms': dense_ms, 'dense_route_value': val_dense, 'dense_route_info': list(info),
        'dense_route_cholesky_tflops_if_all_time_were_factorisation': flops_chol / (dense_ms * 1e-3) / 1e12,
        'dense_route_diagonal_cov_ms': dvar_ms, 'gram_route_same_input_ms': gram_ms,
        'routes_agree_abs': abs(val_d - val_g), 'value_diagonal_cov': val_g,
        'yardstick': {
            'cublas_dgemm_8192_tflops': dgemm_tflops, 'cublas_dgemm_8192_ms': mm_ms,
            'cusolver_potrf_same_size_ms': potrf_ms,
            'cusolver_potrf_tflops': flops_chol / (potrf_ms * 1e-3) / 1e12,
            'cublas_syrk_like_utu_ms': syrk_ms,
        },
        'note': ('the dense route is the correctness path for dense measurement covariances (and the LU fallback '
                 'for non-SPD input); IMB_ROUTE_AUTO takes the Gram route whenever the measurement covariance '
                 'is diagonal, which is every configuration BASELINE.json names'),
    }
    print(json.dumps(line))
    if args.json:
        with open(args.json, 'w') as f:
            json.dump(line, f, indent=1)


if __name__ == '__main__':
    main()
