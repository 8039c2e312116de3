// Stage 3a, batched: the Gram route of EnsembleLikelihood.call for ALL observables of one
// evaluation in a fixed sequence of 9 launches and ONE host read-back
// (likelihoods/ensemble_likelihood.py:86-105 loops over keys; every key has the same [N, m]).
//
//   mean, d = meas - mean                                   colmean   grid (m/256, K)
//   G = U U^T  (FP64 tensor cores, DMMA m8n8k4)             gram      grid (slices, nb^2, K)
//   tr S, tr S^2 -> rho, delta = rho tr S / m, c = (1-rho)/N            oas
//   D = var + delta, w = 1/D, sum d^2 w, sum log(2 pi D)    dvec
//   Gw = U diag(w) U^T                                      gram (weighted)
//   y = U (w o d)                                           wdot
//   K = I + c Gw; chol(K) with y riding along as an extra row; log|K|; quad; loglike   finish
//
// Identical number to the reference's dense LU route (Woodbury + determinant lemma).
// All reductions have a fixed order: results are reproducible run to run and independent of K.
#include "common.cuh"
#include "linalg.cuh"

namespace imb {

constexpr int KMAX = 16;           // observables per batch launch
constexpr int GCOLS = 256;         // columns per Gram slice (fixed => summation order depends on m only)

struct BatchPtrs {
  const double* sims[KMAX];
  const double* meas[KMAX];
  const double* var[KMAX];
};

// workspace of one key; all pointers into the library scratch
struct KeyWs {
  double *mean, *diff, *w, *partial, *G, *ypart, *dpart;
  LikeScalars* sc;
};
struct BatchWs { KeyWs k[KMAX]; };

__global__ void __launch_bounds__(256) bg_colmean_kernel(const __grid_constant__ BatchPtrs p,
                                                         const __grid_constant__ BatchWs ws, int N, int m) {
  const int key = blockIdx.y;
  const double* __restrict__ X = p.sims[key];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = X[i];
  for (int r = 1; r < N; ++r) s = __dadd_rn(s, X[(size_t)r * m + i]);
  const double mu = __ddiv_rn(s, (double)N);
  ws.k[key].mean[i] = mu;
  ws.k[key].diff[i] = __dsub_rn(p.meas[key][i], mu);
}

// acc = A * B^T over the slice columns; when the row blocks coincide (always for N <= 64) the
// tile is staged once and written to both operand buffers (weighted copy for A).
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) bg_gram_kernel(const __grid_constant__ BatchPtrs p,
                                                      const __grid_constant__ BatchWs ws, int N, int m,
                                                      int cols_per_slice) {
  __shared__ double sA[KC][LDS_];
  __shared__ double sB[KC][LDS_];
  const int key = blockIdx.z;
  const KeyWs& kw = ws.k[key];
  const int nb = (N + TB - 1) / TB;
  const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
  const int c0 = blockIdx.x * cols_per_slice;
  const int c_end = min(m, c0 + cols_per_slice);
  const double* __restrict__ X = p.sims[key];
  const double* __restrict__ mean = kw.mean;
  const double* __restrict__ w = kw.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
  for (int k0 = c0; k0 < c_end; k0 += KC) {
    for (int e = threadIdx.x; e < TB * KC; e += 256) {
      const int k = e & (KC - 1), i = e >> 4;
      const int col = k0 + k;
      double a = 0.0, b = 0.0;
      if (col < c_end) {
        const double mu = mean[col];
        const int ra = bi * TB + i, rb = bj * TB + i;
        if (rb < N) b = X[(size_t)rb * m + col] - mu;
        if (bi == bj) a = b;
        else if (ra < N) a = X[(size_t)ra * m + col] - mu;
        if (WEIGHTED) a *= w[col];
      }
      sA[k][i] = a;
      sB[k][i] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      const double a = sA[kk + fk][warp * 8 + fr];
#pragma unroll
      for (int q = 0; q < 8; ++q) dmma884(acc[q][0], acc[q][1], a, sB[kk + fk][q * 8 + fr]);
    }
    __syncthreads();
  }
  const int i = bi * TB + warp * 8 + (lane >> 2);
  if (i >= N) return;
  double* out = kw.partial + (size_t)blockIdx.x * N * N;
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = bj * TB + q * 8 + 2 * (lane & 3) + h;
      if (j < N) out[(size_t)i * N + j] = acc[q][h];
    }
}

__global__ void __launch_bounds__(256) bg_reduce_kernel(const __grid_constant__ BatchWs ws, int nn,
                                                        int slices) {
  const KeyWs& kw = ws.k[blockIdx.y];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nn) return;
  double s = 0.0;
  for (int k = 0; k < slices; ++k) s += kw.partial[(size_t)k * nn + e];
  kw.G[e] = s;
}

// OAS shrinkage (covariance_estimator.py:222-234) from tr G and ||G||_F
__global__ void __launch_bounds__(256) bg_oas_kernel(const __grid_constant__ BatchWs ws, int N, int m,
                                                     int estimator) {
  __shared__ double red[32];
  const KeyWs& kw = ws.k[blockIdx.x];
  Comp t, v;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const double g = kw.G[e];
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
    LikeScalars* sc = kw.sc;
    sc->trs = trs; sc->trs2 = trs2; sc->rho = rho;
    sc->delta = (estimator == IMB_EST_OAS) ? rho * trs / dm : 0.0;
    sc->c = (1.0 - rho) / n;
  }
}

__global__ void __launch_bounds__(256) bg_dvec_kernel(const __grid_constant__ BatchPtrs p,
                                                      const __grid_constant__ BatchWs ws, int m) {
  __shared__ double red[32];
  const KeyWs& kw = ws.k[blockIdx.y];
  const double* __restrict__ var = p.var[blockIdx.y];
  const double delta = kw.sc->delta;
  Comp q, l;
  double nonpos = 0.0, neg = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const double D = __dadd_rn(var[i], delta);
    const double wi = 1.0 / D;
    kw.w[i] = wi;
    const double d = kw.diff[i];
    q.add(__dmul_rn(d, __dmul_rn(wi, d)));
    l.add(log(__dmul_rn(__dmul_rn(fabs(D), 2.0), 3.141592653589793)));
    if (!(D > 0.0)) nonpos += 1.0;
    if (D < 0.0) neg += 1.0;
  }
  const double a = block_sum(q.value(), red), b = block_sum(l.value(), red);
  const double c = block_sum(nonpos, red), d = block_sum(neg, red);
  if (threadIdx.x == 0) {
    double* o = kw.dpart + 4 * blockIdx.x;
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
  }
}

// ypart[s][r] = sum_{cols of slice s} (X[r,col]-mean[col]) w[col] d[col]; warp per row, 8 rows in flight
__global__ void __launch_bounds__(256) bg_wdot_kernel(const __grid_constant__ BatchPtrs p,
                                                      const __grid_constant__ BatchWs ws, int N, int m,
                                                      int cols_per_slice) {
  const KeyWs& kw = ws.k[blockIdx.y];
  const double* __restrict__ X = p.sims[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * cols_per_slice, c_end = min(m, c0 + cols_per_slice);
  for (int r = warp; r < N; r += 8) {
    Comp s;
    const double* row = X + (size_t)r * m;
    for (int c = c0 + lane; c < c_end; c += 32) s.add((row[c] - kw.mean[c]) * kw.w[c] * kw.diff[c]);
    const double v = warp_sum(s.value());
    if (lane == 0) kw.ypart[(size_t)blockIdx.x * N + r] = v;
  }
}

// One block per key.  T = [K; y^T] is (N+1) x N (leading dimension ldt) in shared or global memory.
// Right-looking Cholesky with ONE barrier per column: the update of column block k reads the
// unscaled column k and divides by the pivot, while the scaled column is written to a separate
// array L (so no thread overwrites what another still reads).
__global__ void __launch_bounds__(256) bg_finish_kernel(const __grid_constant__ BatchWs ws, int N,
                                                        int slices, int dparts, int use_smem, int ldt) {
  extern __shared__ double smT[];
  __shared__ double red[32];
  __shared__ double sh[4];
  __shared__ int bad;
  const KeyWs& kw = ws.k[blockIdx.x];
  LikeScalars* sc = kw.sc;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* T = use_smem ? smT : kw.partial;         // partial is free again: reuse as global scratch
  {
    double a = 0, b = 0, c = 0, d = 0;              // fixed order: thread t takes parts t, t+nt, ...
    for (int k = tid; k < dparts; k += nt) {
      a += kw.dpart[4 * k]; b += kw.dpart[4 * k + 1]; c += kw.dpart[4 * k + 2]; d += kw.dpart[4 * k + 3];
    }
    a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red); d = block_sum(d, red);
    if (tid == 0) { sh[0] = a; sh[1] = b; sh[2] = c; sh[3] = d; bad = 0; }
  }
  const double c = sc->c;
  for (int e = tid; e < N * N; e += nt) {
    const int i = e / N, j = e % N;
    T[(size_t)i * ldt + j] = (i == j ? 1.0 : 0.0) + c * kw.G[e];
  }
  for (int r = tid; r < N; r += nt) {
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += kw.ypart[(size_t)k * N + r];
    T[(size_t)N * ldt + r] = s;
  }
  __syncthreads();
  // after step k: column k of T holds UNSCALED values a_ik (i >= k); pivot = a_kk
  double ld = 0.0, zz = 0.0;
  for (int k = 0; k < N; ++k) {
    const double pv = T[(size_t)k * ldt + k];
    if (tid == 0 && !(pv > 0.0) && bad == 0) bad = k + 1;
    const double inv = 1.0 / pv;
    // rows k+1 .. N (row N = rhs), columns k+1 .. min(i, N-1): warp per row, lanes along columns
    for (int i = k + 1 + (tid >> 5); i <= N; i += (nt >> 5)) {
      const double lik = T[(size_t)i * ldt + k] * inv;
      const int jmax = i < N ? i : N - 1;
      for (int j = k + 1 + (tid & 31); j <= jmax; j += 32)
        T[(size_t)i * ldt + j] -= lik * T[(size_t)j * ldt + k];
    }
    if (tid == 0) {
      ld += log(pv);                             // log|K| = sum log(pivot)
      const double z = T[(size_t)N * ldt + k];   // z_k = a_Nk / sqrt(pivot)
      zz += z * z * inv;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double quad = sh[0] - c * zz;
    const double logdet = sh[1] + ld;
    const double sign = (((long long)sh[3]) & 1) ? -1.0 : 1.0;
    sc->s_quad = sh[0]; sc->s_logdet = sh[1];
    sc->quad = quad; sc->logdet = logdet; sc->sign = sign;
    sc->nonpos = (int)sh[2];
    sc->info = bad;
    sc->loglike = -0.5 * (quad + sign * logdet);
  }
}

// ---- host ---------------------------------------------------------------------------------
static int run_batch(int K, const double* const* sims, int N, int m, const double* const* meas,
                     const double* const* var, int estimator, LikeScalars* h_out, cudaStream_t s) {
  const int nn = N * N;
  const int dparts = grid_blocks(m, 256, 1);
  const int cps = GCOLS;
  const int slices = (m + cps - 1) / cps;
  const size_t tn = (size_t)(N + 1) * (N + 1);
  const size_t partial_elems = (size_t)slices * nn > tn ? (size_t)slices * nn : tn;
  const size_t per_key = 256 * 10 + sizeof(LikeScalars) +
                         sizeof(double) * ((size_t)3 * m + partial_elems + nn + (size_t)slices * N + 4 * (size_t)dparts);
  IMB_TRY(ensure_workspace(per_key * K + 256));
  Carver c(state().ws);
  BatchPtrs p;
  BatchWs ws;
  LikeScalars* sc0 = c.take<LikeScalars>(K);
  for (int k = 0; k < K; ++k) {
    p.sims[k] = sims[k]; p.meas[k] = meas[k]; p.var[k] = var[k];
    KeyWs& w = ws.k[k];
    w.sc = sc0 + k;
    w.mean = c.take<double>(m);
    w.diff = c.take<double>(m);
    w.w = c.take<double>(m);
    w.partial = c.take<double>(partial_elems);
    w.G = c.take<double>(nn);
    w.ypart = c.take<double>((size_t)slices * N);
    w.dpart = c.take<double>(4 * (size_t)dparts);
  }
  const int nb = (N + TB - 1) / TB;
  bg_colmean_kernel<<<dim3((m + 255) / 256, K), 256, 0, s>>>(p, ws, N, m);
  IMB_LAUNCHED();
  bg_gram_kernel<false><<<dim3(slices, nb * nb, K), 256, 0, s>>>(p, ws, N, m, cps);
  IMB_LAUNCHED();
  bg_reduce_kernel<<<dim3((nn + 255) / 256, K), 256, 0, s>>>(ws, nn, slices);
  IMB_LAUNCHED();
  bg_oas_kernel<<<K, 256, 0, s>>>(ws, N, m, estimator);
  IMB_LAUNCHED();
  bg_dvec_kernel<<<dim3(dparts, K), 256, 0, s>>>(p, ws, m);
  IMB_LAUNCHED();
  bg_gram_kernel<true><<<dim3(slices, nb * nb, K), 256, 0, s>>>(p, ws, N, m, cps);
  IMB_LAUNCHED();
  bg_reduce_kernel<<<dim3((nn + 255) / 256, K), 256, 0, s>>>(ws, nn, slices);
  IMB_LAUNCHED();
  bg_wdot_kernel<<<dim3(slices, K), 256, 0, s>>>(p, ws, N, m, cps);
  IMB_LAUNCHED();
  const int ldt = N + 1;
  const size_t smem = sizeof(double) * (size_t)(N + 1) * ldt;
  const int use_smem = smem <= 200 * 1024;
  static size_t attr_bytes = 0;
  if (use_smem && smem > 48 * 1024 && smem > attr_bytes) {
    IMB_CUDA(cudaFuncSetAttribute(bg_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  bg_finish_kernel<<<K, 256, use_smem ? smem : 0, s>>>(ws, N, slices, dparts, use_smem, ldt);
  IMB_LAUNCHED();
  LikeScalars* h = static_cast<LikeScalars*>(state().pinned);
  IMB_CUDA(cudaMemcpyAsync(h, sc0, sizeof(LikeScalars) * K, cudaMemcpyDeviceToHost, s));
  IMB_CUDA(cudaStreamSynchronize(s));
  for (int k = 0; k < K; ++k) h_out[k] = h[k];
  return IMB_OK;
}

// Gram route for K keys; fallback[k] = 1 where D had non-positive entries / K was not SPD
int loglike_gram_batch(int K, const double* const* sims, int N, int m, const double* const* meas,
                       const double* const* var, int estimator, double* out, int32_t* info,
                       int* fallback, cudaStream_t s) {
  for (int k0 = 0; k0 < K; k0 += KMAX) {
    const int kb = (K - k0) < KMAX ? (K - k0) : KMAX;
    LikeScalars h[KMAX];
    IMB_TRY(run_batch(kb, sims + k0, N, m, meas + k0, var + k0, estimator, h, s));
    for (int k = 0; k < kb; ++k) {
      out[k0 + k] = h[k].loglike;
      fallback[k0 + k] = (h[k].nonpos > 0 || h[k].info != 0) ? 1 : 0;
      if (info) {
        info[2 * (k0 + k)] = IMB_INFO_ROUTE_GRAM | IMB_INFO_CHOLESKY;
        info[2 * (k0 + k) + 1] = h[k].info;
      }
    }
  }
  return IMB_OK;
}

}  // namespace imb

using namespace imb;

extern "C" int imb_loglike_ensemble_batch(int K, const double* const* sims, int N, int m,
                                          const double* const* meas, const double* const* var,
                                          int estimator, double* out_loglike, int32_t* out_info,
                                          void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (K < 1 || !sims || !meas || !var || !out_loglike || N < 1 || m < 1) return IMB_EINVAL;
  if (estimator != IMB_EST_EMPIRICAL && estimator != IMB_EST_OAS) return IMB_EINVAL;
  for (int k = 0; k < K; ++k)
    if (!sims[k] || !meas[k] || !var[k]) return IMB_EINVAL;
  if (state().pinned_bytes < sizeof(LikeScalars) * KMAX) return IMB_EUNSUPPORTED;
  int* fallback = new int[K];
  int rc = loglike_gram_batch(K, sims, N, m, meas, var, estimator, out_loglike, out_info, fallback, s);
  for (int k = 0; rc == IMB_OK && k < K; ++k) {
    if (!fallback[k]) continue;      // non-positive D (or K not SPD): dense LU route for that key
    rc = loglike_dense(sims[k], N, m, meas[k], var[k], IMB_COV_VARIANCE, estimator, 0, 1,
                       out_loglike + k, out_info ? out_info + 2 * k : nullptr, s);
  }
  delete[] fallback;
  return rc;
}
