// Stage 3a: ensemble statistics and the log-likelihood routes that never form an m x m matrix:
//   * EnsembleLikelihood with a diagonal measurement covariance -> N x N Gram route
//     (Woodbury + matrix-determinant lemma; identical number to the reference's dense LU,
//     SURVEY.md A.5) -- O(m N^2) instead of O(m^3)
//   * EnsembleLikelihoodDiagonal, SimpleLikelihood (variance / no covariance)
// Reference arithmetic: oracle/likelihood.py; file:line citations in include/imagine_b200.h.
// Every reduction has a fixed order (no atomics) so results are run-to-run reproducible.
#include "common.cuh"
#include "linalg.cuh"

namespace imb {

// ---- column statistics -------------------------------------------------------------------
// mean[i] = (sum_r X[r,i]) / N, rows ascending (numpy add.reduce over axis 0, then divide)
__global__ void colmean_kernel(const double* __restrict__ X, int N, int m,
                               const double* __restrict__ meas, double* __restrict__ mean,
                               double* __restrict__ diff) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    double s = X[i];
    for (int r = 1; r < N; ++r) s = __dadd_rn(s, X[(size_t)r * m + i]);
    const double mu = __ddiv_rn(s, (double)N);
    if (mean) mean[i] = mu;
    if (diff) diff[i] = __dsub_rn(meas[i], mu);
  }
}

// numpy var(axis=0): mean, then sum of squared deviations / N, rows ascending
__global__ void colvar_kernel(const double* __restrict__ X, int N, int m, double* __restrict__ var) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    double s = X[i];
    for (int r = 1; r < N; ++r) s = __dadd_rn(s, X[(size_t)r * m + i]);
    const double mu = __ddiv_rn(s, (double)N);
    double d0 = __dsub_rn(X[i], mu);
    double q = __dmul_rn(d0, d0);
    for (int r = 1; r < N; ++r) {
      const double d = __dsub_rn(X[(size_t)r * m + i], mu);
      q = __dadd_rn(q, __dmul_rn(d, d));
    }
    var[i] = __ddiv_rn(q, (double)N);
  }
}

// ---- Gram matrix of the centred (optionally column-weighted) ensemble ----------------------
// partial[s][a][b] = sum_{i in slice s} (X[a,i]-mu_i) w_i (X[b,i]-mu_i)
// grid = (slices, nb*nb) with 64x64 output blocks; 256 threads, 4x4 outputs per thread.
constexpr int GT = 64;   // Gram block edge
constexpr int GK = 32;   // columns staged per iteration

__global__ void __launch_bounds__(256) gram_partial_kernel(const double* __restrict__ X, int N,
                                                           int m, const double* __restrict__ mean,
                                                           const double* __restrict__ w,
                                                           int cols_per_slice,
                                                           double* __restrict__ partial) {
  __shared__ double sA[GK][GT + 1];
  __shared__ double sB[GK][GT + 1];
  const int nb = (N + GT - 1) / GT;
  const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c_begin = blockIdx.x * cols_per_slice;
  const int c_end = min(m, c_begin + cols_per_slice);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int c0 = c_begin; c0 < c_end; c0 += GK) {
    // stage: 64 rows x 32 columns for each operand; lanes run along columns (coalesced)
    for (int e = threadIdx.x; e < GT * GK; e += 256) {
      const int k = e & (GK - 1), row = e >> 5;
      const int col = c0 + k;
      double a = 0.0, b = 0.0;
      if (col < c_end) {
        const double mu = mean[col];
        const int ra = bi * GT + row, rb = bj * GT + row;
        if (ra < N) {
          a = X[(size_t)ra * m + col] - mu;
          if (w) a *= w[col];
        }
        if (rb < N) b = X[(size_t)rb * m + col] - mu;
      }
      sA[k][row] = a;
      sB[k][row] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < GK; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partial + (size_t)blockIdx.x * N * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ra = bi * GT + ty * 4 + i;
    if (ra >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rb = bj * GT + tx * 4 + j;
      if (rb < N) out[(size_t)ra * N + rb] = acc[i][j];
    }
  }
}

// G[a][b] = sum_s partial[s][a][b] in slice order
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int slices, int nn,
                                   double* __restrict__ G) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += partial[(size_t)k * nn + e];
    G[e] = s;
  }
}

// OAS shrinkage from the Gram matrix: tr S = tr G / N, tr S^2 = ||G||_F^2 / N^2
// (covariance_estimator.py:222-234).  Single block.
__global__ void oas_scalars_kernel(const double* __restrict__ G, int N, int m, int estimator,
                                   LikeScalars* __restrict__ sc) {
  __shared__ double red[32];
  Comp t, v;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const double g = G[e];
    v.add(g * g);
    if (e / N == e % N) t.add(g);
  }
  const double trg = block_sum(t.value(), red);
  const double frob = block_sum(v.value(), red);
  if (threadIdx.x == 0) {
    const double n = (double)N, dm = (double)m;
    const double trs = trg / n, trs2 = frob / (n * n);
    double rho = 0.0;
    if (estimator == IMB_EST_OAS) {
      const double num = (1.0 - 2.0 / dm) * trs2 + trs * trs;
      const double den = (n + 1.0 - 2.0 / dm) * (trs2 - (trs * trs) / dm);
      rho = (den == 0.0) ? 1.0 : fmin(1.0, num / den);
    }
    sc->trs = trs;
    sc->trs2 = trs2;
    sc->rho = rho;
    sc->delta = (estimator == IMB_EST_OAS) ? rho * trs / dm : 0.0;
    sc->c = (1.0 - rho) / n;
  }
}

// EnsembleLikelihoodDiagonal / SimpleLikelihood('variance') / SimpleLikelihood(no cov):
//   mode 0: v = var + X.var(0); mode 1: v = var; mode 2: no covariance (-1/2 d.d)
__global__ void __launch_bounds__(256) diag_terms_kernel(const double* __restrict__ var,
                                                         const double* __restrict__ xvar,
                                                         const double* __restrict__ diff, int m,
                                                         int mode, double* __restrict__ part) {
  __shared__ double red[32];
  Comp q, l;
  double neg = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const double d = diff[i];
    if (mode == 2) {
      q.add(__dmul_rn(d, d));
      continue;
    }
    const double v = (mode == 0) ? __dadd_rn(var[i], xvar[i]) : var[i];
    q.add(__dmul_rn(d, __dmul_rn(1.0 / v, d)));
    l.add(log(__dmul_rn(__dmul_rn(v, 2.0), 3.141592653589793)));  // NaN for v < 0, like numpy
    if (v < 0.0) neg += 1.0;
  }
  const double a = block_sum(q.value(), red), b = block_sum(l.value(), red);
  const double c = block_sum(neg, red);
  if (threadIdx.x == 0) {
    part[4 * blockIdx.x] = a; part[4 * blockIdx.x + 1] = b; part[4 * blockIdx.x + 2] = c;
    part[4 * blockIdx.x + 3] = 0.0;
  }
}

__global__ void diag_finish_kernel(const double* __restrict__ part, int nparts, int mode,
                                   const double* __restrict__ var, int m,
                                   LikeScalars* __restrict__ sc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0, b = 0, c = 0;
  for (int k = 0; k < nparts; ++k) { a += part[4 * k]; b += part[4 * k + 1]; c += part[4 * k + 2]; }
  double sign = (((long long)c) & 1) ? -1.0 : 1.0;
  sc->quad = a; sc->logdet = b; sc->sign = sign; sc->info = 0; sc->nonpos = (int)c;
  sc->loglike = (mode == 2) ? -0.5 * a : (-0.5 * a - 0.5 * sign * b);
}

__global__ void gather_columns_kernel(const double* __restrict__ in, int rows, int m,
                                      const int* __restrict__ idx, int m_out,
                                      double* __restrict__ out) {
  const long long total = (long long)rows * m_out;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / m_out), k = (int)(e % m_out);
    out[e] = in[(size_t)r * m + idx[k]];
  }
}

__global__ void gather_cov_kernel(const double* __restrict__ cov, int m, const int* __restrict__ idx,
                                  int m_out, double* __restrict__ out) {
  const long long total = (long long)m_out * m_out;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(e / m_out), b = (int)(e % m_out);
    out[e] = cov[(size_t)idx[a] * m + idx[b]];
  }
}

int grid_blocks(long long work, int threads, int per_sm);

// ---- host orchestration ------------------------------------------------------------------
// Gram matrix G (N x N) of the centred ensemble into `G`; scratch `partial` sized slices*N*N.
int gram_slices(int m) {
  const int by_cols = (m + 4 * GK - 1) / (4 * GK);      // >= 128 columns per slice
  const int cap = state().sm_count;
  return by_cols < 1 ? 1 : (by_cols < cap ? by_cols : cap);
}

int launch_gram(const double* X, int N, int m, const double* mean, const double* w,
                double* partial, double* G, cudaStream_t s) {
  const int slices = gram_slices(m);
  int cps = (m + slices - 1) / slices;
  cps = (cps + GK - 1) / GK * GK;
  const int nb = (N + GT - 1) / GT;
  gram_partial_kernel<<<dim3(slices, nb * nb), 256, 0, s>>>(X, N, m, mean, w, cps, partial);
  IMB_LAUNCHED();
  gram_reduce_kernel<<<grid_blocks((long long)N * N, 256, 4), 256, 0, s>>>(partial, slices, N * N, G);
  IMB_LAUNCHED();
  return IMB_OK;
}

int read_scalars(const LikeScalars* d_sc, LikeScalars* h_out, cudaStream_t s) {
  LikeScalars* h = static_cast<LikeScalars*>(state().pinned);
  IMB_CUDA(cudaMemcpyAsync(h, d_sc, sizeof(LikeScalars), cudaMemcpyDeviceToHost, s));
  IMB_CUDA(cudaStreamSynchronize(s));
  *h_out = *h;
  return IMB_OK;
}

int launch_colmean(const double* X, int N, int m, const double* meas, double* mean, double* diff,
                   cudaStream_t s) {
  colmean_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(X, N, m, meas, mean, diff);
  IMB_LAUNCHED();
  return IMB_OK;
}

int launch_colvar(const double* X, int N, int m, double* var, cudaStream_t s) {
  colvar_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(X, N, m, var);
  IMB_LAUNCHED();
  return IMB_OK;
}

int launch_oas_scalars(const double* G, int N, int m, int estimator, LikeScalars* sc, cudaStream_t s) {
  oas_scalars_kernel<<<1, 256, 0, s>>>(G, N, m, estimator, sc);
  IMB_LAUNCHED();
  return IMB_OK;
}

int loglike_diag(const double* sims, int N, int m, const double* meas, const double* var, int mode,
                 double* out, cudaStream_t s) {
  const int nparts = grid_blocks(m, 256, 2);
  IMB_TRY(ensure_workspace(256 * 8 + sizeof(double) * ((size_t)3 * m + 4 * (size_t)nparts) +
                           sizeof(LikeScalars)));
  ++state().ws_epoch;
  Carver c(state().ws);
  LikeScalars* sc = c.take<LikeScalars>(1);
  double* mean = c.take<double>(m);
  double* diff = c.take<double>(m);
  double* xvar = c.take<double>(m);
  double* part = c.take<double>(4 * (size_t)nparts);
  const int cb = grid_blocks(m, 256, 4);
  colmean_kernel<<<cb, 256, 0, s>>>(sims, N, m, meas, mean, diff);
  IMB_LAUNCHED();
  if (mode == 0) {
    colvar_kernel<<<cb, 256, 0, s>>>(sims, N, m, xvar);
    IMB_LAUNCHED();
  }
  diag_terms_kernel<<<nparts, 256, 0, s>>>(var, xvar, diff, m, mode, part);
  IMB_LAUNCHED();
  diag_finish_kernel<<<1, 32, 0, s>>>(part, nparts, mode, var, m, sc);
  IMB_LAUNCHED();
  LikeScalars h;
  IMB_TRY(read_scalars(sc, &h, s));
  *out = h.loglike;
  return IMB_OK;
}

}  // namespace imb

using namespace imb;

extern "C" {

int imb_ensemble_mean(const double* sims, int N, int m, double* mean, void* stream) {
  IMB_TRY(check_init());
  if (!sims || !mean || N < 1 || m < 0) return IMB_EINVAL;
  if (m == 0) return IMB_OK;
  colmean_kernel<<<grid_blocks(m, 256, 4), 256, 0, (cudaStream_t)stream>>>(sims, N, m, nullptr, mean,
                                                                        nullptr);
  IMB_LAUNCHED();
  return IMB_OK;
}

int imb_ensemble_var(const double* sims, int N, int m, double* var, void* stream) {
  IMB_TRY(check_init());
  if (!sims || !var || N < 1 || m < 0) return IMB_EINVAL;
  if (m == 0) return IMB_OK;
  colvar_kernel<<<grid_blocks(m, 256, 4), 256, 0, (cudaStream_t)stream>>>(sims, N, m, var);
  IMB_LAUNCHED();
  return IMB_OK;
}

int imb_loglike_ensemble_diagonal(const double* sims, int N, int m, const double* meas,
                                  const double* var, double* out_loglike, void* stream) {
  IMB_TRY(check_init());
  if (!sims || !meas || !var || !out_loglike || N < 1 || m < 1) return IMB_EINVAL;
  return loglike_diag(sims, N, m, meas, var, 0, out_loglike, (cudaStream_t)stream);
}

int imb_loglike_ensemble(const double* sims, int N, int m, const double* meas, const double* cov,
                         int cov_kind, int estimator, int route, int use_trace_approximation,
                         double* out_loglike, int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (!sims || !meas || !cov || !out_loglike || N < 1 || m < 1) return IMB_EINVAL;
  if (cov_kind != IMB_COV_VARIANCE && cov_kind != IMB_COV_DENSE) return IMB_EINVAL;
  if (estimator < IMB_EST_EMPIRICAL || estimator > IMB_EST_DIAGONAL) return IMB_EINVAL;
  const bool gram_legal = cov_kind == IMB_COV_VARIANCE && estimator != IMB_EST_DIAGONAL &&
                          !use_trace_approximation;
  if (route == IMB_ROUTE_GRAM && !gram_legal) return IMB_EINVAL;
  if (out_info) out_info[0] = out_info[1] = 0;
  if ((route == IMB_ROUTE_AUTO || route == IMB_ROUTE_GRAM) && gram_legal) {
    int fallback = 0;
    IMB_TRY(loglike_gram_batch(1, &sims, N, m, &meas, &cov, estimator, out_loglike, out_info, &fallback, s));
    if (!fallback) return IMB_OK;
    if (route == IMB_ROUTE_GRAM) return IMB_OK;   // caller insisted; info[1] != 0 flags it
  }
  return loglike_dense(sims, N, m, meas, cov, cov_kind, estimator, use_trace_approximation,
                       /*ensemble=*/1, out_loglike, out_info, s);
}

int imb_loglike_simple(const double* sims, int N, int m, const double* meas, const double* cov,
                       int cov_kind, double* out_loglike, int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (!sims || !meas || !out_loglike || N < 1 || m < 1) return IMB_EINVAL;
  if (out_info) out_info[0] = out_info[1] = 0;
  if (cov_kind == IMB_COV_NONE) return loglike_diag(sims, N, m, meas, nullptr, 2, out_loglike, s);
  if (!cov) return IMB_EINVAL;
  if (cov_kind == IMB_COV_VARIANCE) return loglike_diag(sims, N, m, meas, cov, 1, out_loglike, s);
  if (cov_kind != IMB_COV_DENSE) return IMB_EINVAL;
  return loglike_dense(sims, N, m, meas, cov, IMB_COV_DENSE, IMB_EST_EMPIRICAL, 0, /*ensemble=*/0,
                       out_loglike, out_info, s);
}

int imb_gather_columns(const double* in, int rows, int m, const int32_t* idx, int m_out, double* out,
                       void* stream) {
  IMB_TRY(check_init());
  if (rows < 0 || m_out < 0) return IMB_EINVAL;
  if (rows == 0 || m_out == 0) return IMB_OK;
  if (!in || !idx || !out) return IMB_EINVAL;
  gather_columns_kernel<<<grid_blocks((long long)rows * m_out, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      in, rows, m, idx, m_out, out);
  IMB_LAUNCHED();
  return IMB_OK;
}

int imb_gather_cov(const double* cov, int m, const int32_t* idx, int m_out, double* out, void* stream) {
  IMB_TRY(check_init());
  if (m_out < 0) return IMB_EINVAL;
  if (m_out == 0) return IMB_OK;
  if (!cov || !idx || !out) return IMB_EINVAL;
  gather_cov_kernel<<<grid_blocks((long long)m_out * m_out, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      cov, m, idx, m_out, out);
  IMB_LAUNCHED();
  return IMB_OK;
}

}  // extern "C"
