// Stage 3a, batched: the Gram route of EnsembleLikelihood.call for ALL observables of one
// evaluation (likelihoods/ensemble_likelihood.py:86-105 loops over keys; every key has the same
// [N, m]) in three phases, so that an ensemble whose PIXELS are sharded over ranks can insert one
// small all-reduce after each of the first two:
//
//   phase 1  mean, d = meas - mean; partial G = U U^T over the local columns (FP64 tensor cores)
//            ---- all-reduce G [K, N, N] ----
//   phase 2  tr S, tr S^2 -> rho, delta = rho tr S / m, c = (1-rho)/N          (m = TOTAL pixels)
//            D = var + delta, w = 1/D; partial sums of d^2 w, log(2 pi D); partial Gw = U diag(w) U^T,
//            partial y = U (w o d)
//            ---- all-reduce P2 [K, N*N + N + 4] ----
//   phase 3  K = I + c Gw; Cholesky with y riding along as an extra row; log|K|; quad; loglike
//
// On one GPU the three phases run back to back (11 launches, ONE host read-back for all keys).
// Identical number to the reference's dense LU route (Woodbury + determinant lemma).
// Every reduction has a fixed order: results are reproducible run to run and independent of K.
// Rows may live in a blocked layout (what an all-to-all of per-rank map blocks produces):
//   row r  ->  base + (r / rows_per_block) * block_stride + (r % rows_per_block) * row_stride.
#include "common.cuh"
#include "linalg.cuh"

namespace imb {

constexpr int KMAX = 16;           // observables per batch launch
// columns per Gram slice: fixed, so the summation order depends on m only (batching never changes a
// key's value); 192 fills 148 SMs x 3 CTAs in 2.9 waves at C3 (5 keys x 256 slices), 256 left a third wave
// that was 16 % full
constexpr int GCOLS = 192;

struct RowLayout {
  int rpb;                 // rows per block
  long long bstride;       // doubles between blocks
  long long rstride;       // doubles between rows of a block
  __host__ __device__ __forceinline__ size_t off(int r) const {
    return (size_t)(r / rpb) * (size_t)bstride + (size_t)(r % rpb) * (size_t)rstride;
  }
};

struct BatchPtrs {
  const double* sims[KMAX];
  const double* meas[KMAX];
  const double* var[KMAX];
};

// workspace of one key; all pointers into the library scratch (G / P2 may be caller buffers)
struct KeyWs {
  double *mean, *diff, *w, *partial, *ypart, *dpart;
  double *G;               // [N*N]  Gram matrix (partial after phase 1, full after the all-reduce)
  double *P2;              // [N*N + N + 4]  Gw, y, (sum d^2 w, sum log 2 pi D, #(D<=0), #(D<0))
  LikeScalars* sc;
};
struct BatchWs { KeyWs k[KMAX]; };

__global__ void __launch_bounds__(256) bg_colmean_kernel(const __grid_constant__ BatchPtrs p,
                                                         const __grid_constant__ BatchWs ws, int N, int m,
                                                         RowLayout L) {
  const int key = blockIdx.y;
  const double* __restrict__ X = p.sims[key];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = X[L.off(0) + i];                       // rows ascending: numpy add.reduce(axis=0)
  for (int r = 1; r < N; ++r) s = __dadd_rn(s, X[L.off(r) + i]);
  const double mu = __ddiv_rn(s, (double)N);
  ws.k[key].mean[i] = mu;
  ws.k[key].diff[i] = __dsub_rn(p.meas[key][i], mu);
}

// acc = A * B^T over the slice columns; when the row blocks coincide (always for N <= 64) the
// tile is staged once and written to both operand buffers (weighted copy for A).
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) bg_gram_kernel(const __grid_constant__ BatchPtrs p,
                                                      const __grid_constant__ BatchWs ws, int N, int m,
                                                      int cols_per_slice, RowLayout L) {
  __shared__ double sA[KC][LDS_];
  __shared__ double sB[KC][LDS_];
  const int key = blockIdx.z;
  const KeyWs& kw = ws.k[key];
  const int nb = (N + TB - 1) / TB;
  const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
  if (bj > bi) return;                 // G and Gw are symmetric: upper tiles are mirrored by the reduce
  const int c0 = blockIdx.x * cols_per_slice;
  const int c_end = min(m, c0 + cols_per_slice);
  const double* __restrict__ X = p.sims[key];
  const double* __restrict__ mean = kw.mean;
  const double* __restrict__ w = kw.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  // each thread stages the same 4 (row, k) slots every iteration: hoist the row offsets
  const int sk = threadIdx.x & (KC - 1);
  size_t offA[4], offB[4];
  bool okA[4], okB[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i = (threadIdx.x >> 4) + 16 * t;
    const int ra = bi * TB + i, rb = bj * TB + i;
    okA[t] = ra < N; okB[t] = rb < N;
    offA[t] = okA[t] ? L.off(ra) : 0;
    offB[t] = okB[t] ? L.off(rb) : 0;
  }
  double acc[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
  for (int k0 = c0; k0 < c_end; k0 += KC) {
    const int col = k0 + sk;
    const bool live = col < c_end;
    const double mu = live ? mean[col] : 0.0;
    const double wc = (WEIGHTED && live) ? w[col] : 1.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = (threadIdx.x >> 4) + 16 * t;
      double a = 0.0, b = 0.0;
      if (live) {
        if (okB[t]) b = X[offB[t] + col] - mu;
        if (bi == bj) a = b;
        else if (okA[t]) a = X[offA[t] + col] - mu;
        if (WEIGHTED) a *= wc;
      }
      sA[sk][i] = a;
      sB[sk][i] = b;
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

// dst[e] = sum_s partial[s][e] in slice order; which = 0 -> G, 1 -> P2 (Gw section)
__global__ void __launch_bounds__(256) bg_reduce_kernel(const __grid_constant__ BatchWs ws, int nn,
                                                        int slices, int which, int N) {
  const KeyWs& kw = ws.k[blockIdx.y];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nn) return;
  const int i = e / N, j = e % N;
  const int src = (j / TB > i / TB) ? j * N + i : e;      // upper 64x64 tiles were not computed
  double s = 0.0;
  for (int k = 0; k < slices; ++k) s += kw.partial[(size_t)k * nn + src];
  (which == 0 ? kw.G : kw.P2)[e] = s;
}

// OAS shrinkage (covariance_estimator.py:222-234) from tr G and ||G||_F; m = TOTAL pixel count
__global__ void __launch_bounds__(256) bg_oas_kernel(const __grid_constant__ BatchWs ws, int N,
                                                     double m_total, int estimator) {
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
    const double n = (double)N, dm = m_total;
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

// ypart[s][r] = sum_{cols of slice s} (X[r,col]-mean[col]) w[col] d[col]; warp per row
__global__ void __launch_bounds__(256) bg_wdot_kernel(const __grid_constant__ BatchPtrs p,
                                                      const __grid_constant__ BatchWs ws, int N, int m,
                                                      int cols_per_slice, RowLayout L) {
  const KeyWs& kw = ws.k[blockIdx.y];
  const double* __restrict__ X = p.sims[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * cols_per_slice, c_end = min(m, c0 + cols_per_slice);
  for (int r = warp; r < N; r += 8) {
    Comp s;
    const double* row = X + L.off(r);
    for (int c = c0 + lane; c < c_end; c += 32) s.add((row[c] - kw.mean[c]) * kw.w[c] * kw.diff[c]);
    const double v = warp_sum(s.value());
    if (lane == 0) kw.ypart[(size_t)blockIdx.x * N + r] = v;
  }
}

// P2 tail: y[r] = sum_s ypart[s][r]; (sum d^2 w, sum log 2 pi D, #(D<=0), #(D<0)) = sum of dpart
__global__ void __launch_bounds__(256) bg_pack2_kernel(const __grid_constant__ BatchWs ws, int N,
                                                       int slices, int dparts) {
  __shared__ double red[32];
  const KeyWs& kw = ws.k[blockIdx.x];
  const int tid = threadIdx.x, nt = blockDim.x;
  double* tail = kw.P2 + (size_t)N * N;
  // y[r]: warp per row, lanes stride over the slices, fixed-order shuffle tree
  for (int r = tid >> 5; r < N; r += nt >> 5) {
    double s = 0.0;
    for (int k = tid & 31; k < slices; k += 32) s += kw.ypart[(size_t)k * N + r];
    s = warp_sum(s);
    if ((tid & 31) == 0) tail[r] = s;
  }
  double a = 0, b = 0, c = 0, d = 0;                 // fixed order: thread t takes parts t, t+nt, ...
  for (int k = tid; k < dparts; k += nt) {
    a += kw.dpart[4 * k]; b += kw.dpart[4 * k + 1]; c += kw.dpart[4 * k + 2]; d += kw.dpart[4 * k + 3];
  }
  a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red); d = block_sum(d, red);
  if (tid == 0) { tail[N] = a; tail[N + 1] = b; tail[N + 2] = c; tail[N + 3] = d; }
}

// One block per key.  T = [K; y^T] is (N+1) x N (leading dimension ldt) in shared or global memory.
// Right-looking Cholesky with ONE barrier per column: the update reads the UNSCALED column k and
// divides by the pivot on the fly (column k itself is never rewritten).
__global__ void __launch_bounds__(256) bg_finish_kernel(const __grid_constant__ BatchWs ws, int N,
                                                        int use_smem, int ldt) {
  extern __shared__ double smT[];
  __shared__ int bad;
  const KeyWs& kw = ws.k[blockIdx.x];
  LikeScalars* sc = kw.sc;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* T = use_smem ? smT : kw.partial;           // partial is free again: global scratch
  const double* tail = kw.P2 + (size_t)N * N;
  if (tid == 0) bad = 0;
  const double c = sc->c;
  for (int e = tid; e < N * N; e += nt) {
    const int i = e / N, j = e % N;
    T[(size_t)i * ldt + j] = (i == j ? 1.0 : 0.0) + c * kw.P2[e];
  }
  for (int r = tid; r < N; r += nt) T[(size_t)N * ldt + r] = tail[r];
  __syncthreads();
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
      ld += log(pv);                                 // log|K| = sum log(pivot)
      const double z = T[(size_t)N * ldt + k];       // z_k = a_Nk / sqrt(pivot)
      zz += z * z * inv;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double s_quad = tail[N], s_logdet = tail[N + 1];
    const double quad = s_quad - c * zz;
    const double logdet = s_logdet + ld;
    const double sign = (((long long)tail[N + 3]) & 1) ? -1.0 : 1.0;
    sc->s_quad = s_quad; sc->s_logdet = s_logdet;
    sc->quad = quad; sc->logdet = logdet; sc->sign = sign;
    sc->nonpos = (int)tail[N + 2];
    sc->info = bad;
    sc->loglike = -0.5 * (quad + sign * logdet);
  }
}

// ---- host ---------------------------------------------------------------------------------
struct BatchCtx {
  int K, N, m, slices, dparts, nn;
  RowLayout L;
  BatchPtrs p;
  BatchWs ws;
  LikeScalars* sc0;
};

// Deterministic carving: the three phases of one evaluation see the same pointers as long as no
// other library call grows the workspace in between.  G / P2 may be caller-owned (contiguous
// [K, nn] and [K, nn + N + 4]) so that torch.distributed can all-reduce them.
static int make_ctx(BatchCtx* c, int K, const double* const* sims, int N, int m, RowLayout L,
                    const double* const* meas, const double* const* var, double* G_ext, double* P2_ext) {
  if (K < 1 || K > KMAX) return IMB_EINVAL;
  c->K = K; c->N = N; c->m = m; c->L = L;
  c->nn = N * N;
  c->dparts = grid_blocks(m, 256, 1);
  c->slices = (m + GCOLS - 1) / GCOLS;
  const size_t nn = c->nn, p2 = nn + N + 4;
  const size_t tn = (size_t)(N + 1) * (N + 1);
  const size_t partial_elems = (size_t)c->slices * nn > tn ? (size_t)c->slices * nn : tn;
  const size_t per_key = 256 * 12 + sizeof(LikeScalars) +
                         sizeof(double) * ((size_t)3 * m + partial_elems + nn + p2 + (size_t)c->slices * N +
                                           4 * (size_t)c->dparts);
  IMB_TRY(ensure_workspace(per_key * K + 512));
  Carver cv(state().ws);
  c->sc0 = cv.take<LikeScalars>(K);
  double* G_int = cv.take<double>((size_t)K * nn);
  double* P2_int = cv.take<double>((size_t)K * p2);
  for (int k = 0; k < K; ++k) {
    c->p.sims[k] = sims ? sims[k] : nullptr;
    c->p.meas[k] = meas ? meas[k] : nullptr;
    c->p.var[k] = var ? var[k] : nullptr;
    KeyWs& w = c->ws.k[k];
    w.sc = c->sc0 + k;
    w.G = (G_ext ? G_ext : G_int) + (size_t)k * nn;
    w.P2 = (P2_ext ? P2_ext : P2_int) + (size_t)k * p2;
    w.mean = cv.take<double>(m);
    w.diff = cv.take<double>(m);
    w.w = cv.take<double>(m);
    w.partial = cv.take<double>(partial_elems);
    w.ypart = cv.take<double>((size_t)c->slices * N);
    w.dpart = cv.take<double>(4 * (size_t)c->dparts);
  }
  return IMB_OK;
}

static int phase1(const BatchCtx& c, cudaStream_t s) {
  const int nb = (c.N + TB - 1) / TB;
  bg_colmean_kernel<<<dim3((c.m + 255) / 256, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, c.L);
  IMB_LAUNCHED();
  bg_gram_kernel<false><<<dim3(c.slices, nb * nb, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, GCOLS, c.L);
  IMB_LAUNCHED();
  bg_reduce_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 0, c.N);
  IMB_LAUNCHED();
  return IMB_OK;
}

static int phase2(const BatchCtx& c, double m_total, int estimator, cudaStream_t s) {
  const int nb = (c.N + TB - 1) / TB;
  bg_oas_kernel<<<c.K, 256, 0, s>>>(c.ws, c.N, m_total, estimator);
  IMB_LAUNCHED();
  bg_dvec_kernel<<<dim3(c.dparts, c.K), 256, 0, s>>>(c.p, c.ws, c.m);
  IMB_LAUNCHED();
  bg_gram_kernel<true><<<dim3(c.slices, nb * nb, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, GCOLS, c.L);
  IMB_LAUNCHED();
  bg_reduce_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 1, c.N);
  IMB_LAUNCHED();
  bg_wdot_kernel<<<dim3(c.slices, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, GCOLS, c.L);
  IMB_LAUNCHED();
  bg_pack2_kernel<<<c.K, 256, 0, s>>>(c.ws, c.N, c.slices, c.dparts);
  IMB_LAUNCHED();
  return IMB_OK;
}

static int phase3(const BatchCtx& c, LikeScalars* h_out, cudaStream_t s) {
  const int ldt = c.N + 1;
  const size_t smem = sizeof(double) * (size_t)(c.N + 1) * ldt;
  const int use_smem = smem <= 200 * 1024;
  static size_t attr_bytes = 0;
  if (use_smem && smem > 48 * 1024 && smem > attr_bytes) {
    IMB_CUDA(cudaFuncSetAttribute(bg_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  bg_finish_kernel<<<c.K, 256, use_smem ? smem : 0, s>>>(c.ws, c.N, use_smem, ldt);
  IMB_LAUNCHED();
  LikeScalars* h = static_cast<LikeScalars*>(state().pinned);
  IMB_CUDA(cudaMemcpyAsync(h, c.sc0, sizeof(LikeScalars) * c.K, cudaMemcpyDeviceToHost, s));
  IMB_CUDA(cudaStreamSynchronize(s));
  for (int k = 0; k < c.K; ++k) h_out[k] = h[k];
  return IMB_OK;
}

// device-side export for the asynchronous entry points: value + a flag (0 ok, else Gram route invalid)
__global__ void bg_export_kernel(const __grid_constant__ BatchWs ws, int K, double* __restrict__ out,
                                 int* __restrict__ flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const LikeScalars* sc = ws.k[k].sc;
  out[k] = sc->loglike;
  if (flag) flag[k] = (sc->nonpos > 0 || sc->info != 0) ? (sc->info != 0 ? sc->info : -1) : 0;
}

static int finish_async(const BatchCtx& c, double* out_device, int32_t* flag_device, cudaStream_t s) {
  const int ldt = c.N + 1;
  const size_t smem = sizeof(double) * (size_t)(c.N + 1) * ldt;
  const int use_smem = smem <= 200 * 1024;
  static size_t attr_bytes = 0;
  if (use_smem && smem > 48 * 1024 && smem > attr_bytes) {
    IMB_CUDA(cudaFuncSetAttribute(bg_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  bg_finish_kernel<<<c.K, 256, use_smem ? smem : 0, s>>>(c.ws, c.N, use_smem, ldt);
  IMB_LAUNCHED();
  bg_export_kernel<<<1, 32, 0, s>>>(c.ws, c.K, out_device, flag_device);
  IMB_LAUNCHED();
  return IMB_OK;
}

static void unpack(const LikeScalars* h, int K, double* out, int32_t* info, int* fallback) {
  for (int k = 0; k < K; ++k) {
    out[k] = h[k].loglike;
    if (fallback) fallback[k] = (h[k].nonpos > 0 || h[k].info != 0) ? 1 : 0;
    if (info) {
      info[2 * k] = IMB_INFO_ROUTE_GRAM | IMB_INFO_CHOLESKY;
      info[2 * k + 1] = h[k].info ? h[k].info : (h[k].nonpos > 0 ? -1 : 0);
    }
  }
}

// Gram route for K keys on one GPU; fallback[k] = 1 where D had non-positive entries / K was not SPD
int loglike_gram_batch(int K, const double* const* sims, int N, int m, const double* const* meas,
                       const double* const* var, int estimator, double* out, int32_t* info,
                       int* fallback, cudaStream_t s) {
  const RowLayout L{N, 0, (long long)m};
  for (int k0 = 0; k0 < K; k0 += KMAX) {
    const int kb = (K - k0) < KMAX ? (K - k0) : KMAX;
    BatchCtx c;
    IMB_TRY(make_ctx(&c, kb, sims + k0, N, m, L, meas + k0, var + k0, nullptr, nullptr));
    IMB_TRY(phase1(c, s));
    IMB_TRY(phase2(c, (double)m, estimator, s));
    LikeScalars h[KMAX];
    IMB_TRY(phase3(c, h, s));
    unpack(h, kb, out + k0, info ? info + 2 * k0 : nullptr, fallback + k0);
  }
  return IMB_OK;
}

}  // namespace imb

using namespace imb;

extern "C" {

int imb_loglike_ensemble_batch(int K, const double* const* sims, int N, int m,
                               const double* const* meas, const double* const* var, int estimator,
                               double* out_loglike, int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (K < 1 || !sims || !meas || !var || !out_loglike || N < 1 || m < 1) return IMB_EINVAL;
  if (estimator != IMB_EST_EMPIRICAL && estimator != IMB_EST_OAS) return IMB_EINVAL;
  for (int k = 0; k < K; ++k)
    if (!sims[k] || !meas[k] || !var[k]) return IMB_EINVAL;
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

int imb_loglike_ensemble_batch_async(int K, const double* const* sims, int N, int m,
                                     const double* const* meas, const double* const* var,
                                     int estimator, double* out_device, int32_t* flag_device,
                                     void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (K < 1 || K > KMAX || !sims || !meas || !var || !out_device || N < 1 || m < 1) return IMB_EINVAL;
  if (estimator != IMB_EST_EMPIRICAL && estimator != IMB_EST_OAS) return IMB_EINVAL;
  for (int k = 0; k < K; ++k)
    if (!sims[k] || !meas[k] || !var[k]) return IMB_EINVAL;
  const RowLayout L{N, 0, (long long)m};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, sims, N, m, L, meas, var, nullptr, nullptr));
  IMB_TRY(phase1(c, s));
  IMB_TRY(phase2(c, (double)m, estimator, s));
  return finish_async(c, out_device, flag_device, s);
}

int imb_gram_phase3_async(int K, int N, int m_local, double* G, double* P2, double* out_device,
                          int32_t* flag_device, void* stream) {
  IMB_TRY(check_init());
  if (K < 1 || K > KMAX || N < 1 || m_local < 1 || !G || !P2 || !out_device) return IMB_EINVAL;
  const RowLayout L{N, 0, (long long)m_local};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, nullptr, N, m_local, L, nullptr, nullptr, G, P2));
  return finish_async(c, out_device, flag_device, (cudaStream_t)stream);
}

static int sharded_args_ok(int K, const double* const* sims, int N, int m_local, int rpb) {
  if (K < 1 || K > KMAX || !sims || N < 1 || m_local < 1 || rpb < 1) return 0;
  for (int k = 0; k < K; ++k)
    if (!sims[k]) return 0;
  return 1;
}

int imb_gram_phase1(int K, const double* const* sims, int N, int m_local, int rows_per_block,
                    int64_t block_stride, int64_t row_stride, const double* const* meas, double* G,
                    void* stream) {
  IMB_TRY(check_init());
  if (!sharded_args_ok(K, sims, N, m_local, rows_per_block) || !meas || !G) return IMB_EINVAL;
  const RowLayout L{rows_per_block, (long long)block_stride, (long long)row_stride};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, sims, N, m_local, L, meas, nullptr, G, nullptr));
  return phase1(c, (cudaStream_t)stream);
}

int imb_gram_phase2(int K, const double* const* sims, int N, int m_local, int rows_per_block,
                    int64_t block_stride, int64_t row_stride, const double* const* meas,
                    const double* const* var, int64_t m_total, int estimator, double* G, double* P2,
                    void* stream) {
  IMB_TRY(check_init());
  if (!sharded_args_ok(K, sims, N, m_local, rows_per_block) || !meas || !var || !G || !P2 || m_total < 1)
    return IMB_EINVAL;
  if (estimator != IMB_EST_EMPIRICAL && estimator != IMB_EST_OAS) return IMB_EINVAL;
  const RowLayout L{rows_per_block, (long long)block_stride, (long long)row_stride};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, sims, N, m_local, L, meas, var, G, P2));
  return phase2(c, (double)m_total, estimator, (cudaStream_t)stream);
}

int imb_gram_phase3(int K, int N, int m_local, double* G, double* P2, double* out_loglike,
                    int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  if (K < 1 || K > KMAX || N < 1 || m_local < 1 || !G || !P2 || !out_loglike) return IMB_EINVAL;
  const RowLayout L{N, 0, (long long)m_local};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, nullptr, N, m_local, L, nullptr, nullptr, G, P2));
  LikeScalars h[KMAX];
  IMB_TRY(phase3(c, h, (cudaStream_t)stream));
  unpack(h, K, out_loglike, out_info, nullptr);
  return IMB_OK;
}

}  // extern "C"
