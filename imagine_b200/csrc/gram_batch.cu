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

constexpr int KMAX = 32;           // keys per batch launch (observables x parameter points)
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

// ---- N <= 64: one 64 x 64 output tile per CTA, operands streamed by the TMA engine ------------------
// The first Gram kernel (bg_gram_kernel, kept for N > 64) runs LDG -> st.shared -> barrier -> DMMA ->
// barrier per 16 columns and reaches 0.84 TB/s (ncu r01: DMMA pipe 37 %, 17 warps per issue waiting on
// the loads).  Here raw rows go global -> shared asynchronously: one `cp.async.bulk` (SASS UBLKCP) per
// row and 32-column tile, completing on an mbarrier, in a ring of GSTAGES tiles; centring (x - mean),
// the weights and y = U (w o d) are applied to the fragments in registers.  Only the lower-triangular
// 8 x 8 blocks are computed (36 of 64): warp w of each half takes row blocks {w, 7 - w} = 9 DMMAs per
// 4 columns for every warp; the two warp halves take alternate k-steps and are summed in fixed order.
#ifndef IMB_GRAM_STAGES
#define IMB_GRAM_STAGES 3
#endif
#ifndef IMB_GRAM_TILE
#define IMB_GRAM_TILE 64
#endif
constexpr int GSTAGES = IMB_GRAM_STAGES;
constexpr int GTILE = IMB_GRAM_TILE;       // columns per stage (512-byte row segments: DRAM page locality)
constexpr int GLD = GTILE + 4;             // shared leading dimension: conflict-free 64-bit fragment loads
constexpr int GVEC = 3 * GTILE;            // mean | w | d of the tile's columns
constexpr size_t GRAM64_SMEM = sizeof(double) * ((size_t)GSTAGES * (64 * GLD + GVEC)) + 8 * GSTAGES + 128;

__device__ __forceinline__ void mbar_init(unsigned int bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one 32-column tile: this warp's k-steps (alternate ones per warp half), row blocks WQ and 7 - WQ
template <bool WEIGHTED, int WQ>
__device__ __forceinline__ void gram64_tile(const double* __restrict__ xs, const double* __restrict__ vs,
                                            int nc, int N, int half, int fr, int fk, double (&acc0)[4][2],
                                            double (&acc1)[8][2], double& y0, double& y1) {
  constexpr int R0 = WQ, R1 = 7 - WQ;
#pragma unroll
  for (int ks = 0; ks < GTILE / 8; ++ks) {
    const int k = 4 * (2 * ks + half) + fk;            // this lane's column of the k-step
    const bool valid = k < nc;
    const double mu = vs[k];
    double f[R1 + 1];
#pragma unroll
    for (int q = 0; q <= R1; ++q) f[q] = (valid && 8 * q + fr < N) ? xs[(8 * q + fr) * GLD + k] - mu : 0.0;
    double a0 = f[R0], a1 = f[R1];
    if (WEIGHTED) {
      const double wc = valid ? vs[GTILE + k] : 0.0, dd = valid ? vs[2 * GTILE + k] : 0.0;
      a0 *= wc; a1 *= wc;
      y0 = fma(a0, dd, y0);
      y1 = fma(a1, dd, y1);
    }
#pragma unroll
    for (int q = 0; q <= R0; ++q) dmma884(acc0[q][0], acc0[q][1], a0, f[q]);
#pragma unroll
    for (int q = 0; q <= R1; ++q) dmma884(acc1[q][0], acc1[q][1], a1, f[q]);
  }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(256, 2) bg_gram64_kernel(const __grid_constant__ BatchPtrs p,
                                                           const __grid_constant__ BatchWs ws, int N, int m,
                                                           int cols_per_slice, RowLayout L) {
  extern __shared__ __align__(128) unsigned char g_smem[];
  double* sX = reinterpret_cast<double*>(g_smem);                       // [GSTAGES][64][GLD]
  double* sV = sX + (size_t)GSTAGES * 64 * GLD;                         // [GSTAGES][3][GTILE]
  const unsigned int bar0 = (unsigned int)__cvta_generic_to_shared(sV + (size_t)GSTAGES * GVEC);
  const int key = blockIdx.y;
  const KeyWs& kw = ws.k[key];
  const double* __restrict__ X = p.sims[key];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * cols_per_slice;
  const int c_end = min(m, c0 + cols_per_slice);
  const int ntiles = (c_end - c0 + GTILE - 1) / GTILE;

  // rows >= N stay zero for the whole kernel; stale columns of a short last tile must be finite
  for (int e = tid; e < GSTAGES * (64 * GLD + GVEC); e += 256) sX[e] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GSTAGES; ++s) mbar_init(bar0 + 8 * s, 1);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic writes before async writes
  __syncthreads();

  // warp 0 issues a tile: 2 rows per lane + the column vectors
  auto issue = [&](int t) {
    const int s = t % GSTAGES;
    const int col = c0 + t * GTILE;
    const int nc = min(GTILE, c_end - col);
    const unsigned int bytes = (unsigned int)nc * 8u;
    const unsigned int bar = bar0 + 8 * s;
    if (lane == 0) mbar_expect_tx(bar, bytes * (unsigned int)(N + (WEIGHTED ? 3 : 1)));
    __syncwarp();
    const unsigned int xs = (unsigned int)__cvta_generic_to_shared(sX + (size_t)s * 64 * GLD);
    for (int r = lane; r < N; r += 32) bulk_g2s(xs + (unsigned int)r * GLD * 8u, X + L.off(r) + col, bytes, bar);
    const unsigned int vs = (unsigned int)__cvta_generic_to_shared(sV + (size_t)s * GVEC);
    if (lane == 0) bulk_g2s(vs, kw.mean + col, bytes, bar);
    if (WEIGHTED && lane == 1) bulk_g2s(vs + GTILE * 8u, kw.w + col, bytes, bar);
    if (WEIGHTED && lane == 2) bulk_g2s(vs + 2 * GTILE * 8u, kw.diff + col, bytes, bar);
  };
  if (warp == 0)
    for (int t = 0; t < GSTAGES && t < ntiles; ++t) issue(t);

  const int half = warp >> 2, wq = warp & 3;
  const int r0 = wq, r1 = 7 - wq;                     // the two row blocks of this warp
  const int fr = lane >> 2, fk = lane & 3;
  double acc0[4][2], acc1[8][2];                       // blocks (r0, q <= r0) and (r1, q <= r1)
#pragma unroll
  for (int q = 0; q < 4; ++q) acc0[q][0] = acc0[q][1] = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc1[q][0] = acc1[q][1] = 0.0;
  double y0 = 0.0, y1 = 0.0;

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % GSTAGES;
    mbar_wait(bar0 + 8 * s, (unsigned int)((t / GSTAGES) & 1));
    const double* xs = sX + (size_t)s * 64 * GLD;
    const double* vs = sV + (size_t)s * GVEC;
    const int nc = min(GTILE, c_end - (c0 + t * GTILE));
    switch (wq) {                                      // row blocks as compile-time constants
      case 0: gram64_tile<WEIGHTED, 0>(xs, vs, nc, N, half, fr, fk, acc0, acc1, y0, y1); break;
      case 1: gram64_tile<WEIGHTED, 1>(xs, vs, nc, N, half, fr, fk, acc0, acc1, y0, y1); break;
      case 2: gram64_tile<WEIGHTED, 2>(xs, vs, nc, N, half, fr, fk, acc0, acc1, y0, y1); break;
      default: gram64_tile<WEIGHTED, 3>(xs, vs, nc, N, half, fr, fk, acc0, acc1, y0, y1); break;
    }
    __syncthreads();                                   // every warp is done with stage s
    if (warp == 0 && t + GSTAGES < ntiles) issue(t + GSTAGES);
  }

  // ---- the second warp half hands its sums over; fixed order: half 0 + half 1 ----------------------
  double* red = sX;                                    // [4 warps][12 blocks][2][32 lanes] + y
  __syncthreads();
  if (half == 1) {
    double* o = red + (size_t)wq * (12 * 64 + 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) { o[(q * 2) * 32 + lane] = acc0[q][0]; o[(q * 2 + 1) * 32 + lane] = acc0[q][1]; }
#pragma unroll
    for (int q = 0; q < 8; ++q) { o[((4 + q) * 2) * 32 + lane] = acc1[q][0]; o[((4 + q) * 2 + 1) * 32 + lane] = acc1[q][1]; }
    o[12 * 64 + lane] = y0; o[12 * 64 + 32 + lane] = y1;
  }
  __syncthreads();
  if (half == 0) {
    const double* o = red + (size_t)wq * (12 * 64 + 64);
    double* out = kw.partial + (size_t)blockIdx.x * N * N;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q > r0) continue;
      const int i = 8 * r0 + fr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 8 * q + 2 * fk + h;
        if (i < N && j < N) out[(size_t)i * N + j] = acc0[q][h] + o[(q * 2 + h) * 32 + lane];
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q > r1) continue;
      const int i = 8 * r1 + fr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 8 * q + 2 * fk + h;
        if (i < N && j < N) out[(size_t)i * N + j] = acc1[q][h] + o[((4 + q) * 2 + h) * 32 + lane];
      }
    }
    if (WEIGHTED) {
      y0 += o[12 * 64 + lane]; y1 += o[12 * 64 + 32 + lane];
      // the 4 lanes of a row hold its 4 k-phases
      y0 += __shfl_xor_sync(0xffffffffu, y0, 1); y0 += __shfl_xor_sync(0xffffffffu, y0, 2);
      y1 += __shfl_xor_sync(0xffffffffu, y1, 1); y1 += __shfl_xor_sync(0xffffffffu, y1, 2);
      if (fk == 0) {
        if (8 * r0 + fr < N) kw.ypart[(size_t)blockIdx.x * N + 8 * r0 + fr] = y0;
        if (8 * r1 + fr < N) kw.ypart[(size_t)blockIdx.x * N + 8 * r1 + fr] = y1;
      }
    }
  }
}

// dst[e] = sum_s partial[s][e] in slice order, upper 8 x 8 blocks mirrored; which = 0 -> G, 1 -> P2
__global__ void __launch_bounds__(256) bg_reduce64_kernel(const __grid_constant__ BatchWs ws, int nn,
                                                          int slices, int which, int N) {
  const KeyWs& kw = ws.k[blockIdx.y];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nn) return;
  const int i = e / N, j = e % N;
  const int src = (j / 8 > i / 8) ? j * N + i : e;
  double s = 0.0;
  for (int k = 0; k < slices; ++k) s += kw.partial[(size_t)k * nn + src];
  (which == 0 ? kw.G : kw.P2)[e] = s;
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
__global__ void __launch_bounds__(1024) bg_oas_kernel(const __grid_constant__ BatchWs ws, int N,
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
  if (tid == 0) bad = 0x7fffffff;
  const double c = sc->c;
  for (int e = tid; e < N * N; e += nt) {
    const int i = e / N, j = e % N;
    T[(size_t)i * ldt + j] = (i == j ? 1.0 : 0.0) + c * kw.P2[e];
  }
  for (int r = tid; r < N; r += nt) T[(size_t)N * ldt + r] = tail[r];
  __syncthreads();
  for (int k = 0; k < N; ++k) {
    // nothing but the update sits between two barriers: the pivots' logarithms and the z^2 terms are
    // summed after the loop (column k and z_k = T[N][k] are final once column k - 1 is eliminated)
    const double inv = 1.0 / T[(size_t)k * ldt + k];
    // rows k+1 .. N (row N = rhs), columns k+1 .. min(i, N-1): warp per row, lanes along columns
    for (int i = k + 1 + (tid >> 5); i <= N; i += (nt >> 5)) {
      const double lik = T[(size_t)i * ldt + k] * inv;
      const int jmax = i < N ? i : N - 1;
      for (int j = k + 1 + (tid & 31); j <= jmax; j += 32)
        T[(size_t)i * ldt + j] -= lik * T[(size_t)j * ldt + k];
    }
    __syncthreads();
  }
  // log|K| = sum log(pivot), z^T z = sum z_k^2 / pivot_k: thread t takes k = t, t + nt, ... (fixed order)
  __shared__ double red[32];
  double ld = 0.0, zz = 0.0;
  int first_bad = 0x7fffffff;
  for (int k = tid; k < N; k += nt) {
    const double pv = T[(size_t)k * ldt + k], z = T[(size_t)N * ldt + k];
    if (!(pv > 0.0)) first_bad = min(first_bad, k + 1);
    ld += log(pv);
    zz += z * z / pv;
  }
  ld = block_sum(ld, red);
  zz = block_sum(zz, red);
  atomicMin(&bad, first_bad);
  __syncthreads();
  if (tid == 0) {
    const double s_quad = tail[N], s_logdet = tail[N + 1];
    const double quad = s_quad - c * zz;
    const double logdet = s_logdet + ld;
    const double sign = (((long long)tail[N + 3]) & 1) ? -1.0 : 1.0;
    sc->s_quad = s_quad; sc->s_logdet = s_logdet;
    sc->quad = quad; sc->logdet = logdet; sc->sign = sign;
    sc->nonpos = (int)tail[N + 2];
    sc->info = bad == 0x7fffffff ? 0 : bad;
    sc->loglike = -0.5 * (quad + sign * logdet);
  }
}

// The same for N <= 64 on 1024 threads = 16 row groups x 64 columns: a thread touches at most 5 rows per
// column step, the row loop has exact bounds and constant pointer strides (the step is bound by the
// instructions the block issues: index arithmetic and clamps cost more than the update itself), the
// strictly upper triangle is updated along and never read, and the pivot enters as u_j = t_jk / p_k.
constexpr int FRG = 16;
__global__ void __launch_bounds__(64 * FRG) bg_finish64_kernel(const __grid_constant__ BatchWs ws, int N) {
  __shared__ double T[65][65];
  __shared__ int bad;
  __shared__ double red[32];
  const KeyWs& kw = ws.k[blockIdx.x];
  LikeScalars* sc = kw.sc;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const double* tail = kw.P2 + (size_t)N * N;
  if (tid == 0) bad = 0x7fffffff;
  const double c = sc->c;
  for (int i = ty; i <= 64; i += FRG) {
    double v = 0.0;
    if (tx < N) {
      if (i < N) v = (i == tx ? 1.0 : 0.0) + c * kw.P2[(size_t)i * N + tx];
      else if (i == N) v = tail[tx];
    }
    T[i][tx] = v;
  }
  __syncthreads();
  for (int k = 0; k + 1 < N; ++k) {
    if (tx > k && tx < N && k + 1 + ty <= N) {
      const double u = T[tx][k] * (1.0 / T[k][k]);
      // rows k + 1 + ty, + 16, ... up to N (row N = rhs): exact bounds, pointers stepped by a constant
      const double* pk = &T[k + 1 + ty][k];
      double* pt = &T[k + 1 + ty][tx];
      for (int i = k + 1 + ty; i <= N; i += FRG, pk += FRG * 65, pt += FRG * 65) *pt = fma(-*pk, u, *pt);
    }
    __syncthreads();
  }
  double ld = 0.0, zz = 0.0;
  int first_bad = 0x7fffffff;
  if (tid < N) {
    const double pv = T[tid][tid], z = T[N][tid];
    if (!(pv > 0.0)) first_bad = tid + 1;
    ld = log(pv);
    zz = z * z / pv;
  }
  ld = block_sum(ld, red);
  zz = block_sum(zz, red);
  atomicMin(&bad, first_bad);
  __syncthreads();
  if (tid == 0) {
    const double s_quad = tail[N], s_logdet = tail[N + 1];
    const double quad = s_quad - c * zz;
    const double logdet = s_logdet + ld;
    const double sign = (((long long)tail[N + 3]) & 1) ? -1.0 : 1.0;
    sc->s_quad = s_quad; sc->s_logdet = s_logdet;
    sc->quad = quad; sc->logdet = logdet; sc->sign = sign;
    sc->nonpos = (int)tail[N + 2];
    sc->info = bad == 0x7fffffff ? 0 : bad;
    sc->loglike = -0.5 * (quad + sign * logdet);
  }
}

// 64 < N <= ~370 (C4: N = 256): T = (N + 1) x N does not fit in shared memory, a 64-column PANEL of it
// does.  One block of 1024 threads per key walks the panels left to right: the panel (all rows from its
// diagonal block down to the right-hand-side row) is factored in shared memory with the same branch-free
// column step as above, then the columns to its right are updated in place in global memory (L2-resident,
// 0.5 MB per key) with the panel and its pivot-scaled transpose as shared-memory operands.  The kernel
// it replaces for these sizes kept T in global memory for every one of the N column steps: 2.4 ms at
// N = 256, 56 % of the C4 likelihood stage.
constexpr int FPW = 64;
__host__ __device__ inline size_t finish_panel_smem(int N) {
  return sizeof(double) * ((size_t)(N + 2) * 65 + (size_t)FPW * 65 + 2 * (size_t)N + FPW);
}
__global__ void __launch_bounds__(1024) bg_finish_panel_kernel(const __grid_constant__ BatchWs ws, int N, int ldt) {
  extern __shared__ double fsm[];
  double (*P)[65] = reinterpret_cast<double (*)[65]>(fsm);
  double (*Us)[65] = reinterpret_cast<double (*)[65]>(fsm + (size_t)(N + 2) * 65);
  double* piv = fsm + (size_t)(N + 2) * 65 + (size_t)FPW * 65;
  double* zr = piv + N;
  double* pinv = zr + N;
  __shared__ int bad;
  __shared__ double red[32];
  const KeyWs& kw = ws.k[blockIdx.x];
  LikeScalars* sc = kw.sc;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* T = kw.partial;                      // free again: global scratch, (N + 1) x ldt
  const double* tail = kw.P2 + (size_t)N * N;
  if (tid == 0) bad = 0x7fffffff;
  const double c = sc->c;
  for (int i = ty; i < N; i += FRG)
    for (int j = tx; j < N; j += 64) T[(size_t)i * ldt + j] = (i == j ? 1.0 : 0.0) + c * kw.P2[(size_t)i * N + j];
  for (int r = tid; r < N; r += 1024) T[(size_t)N * ldt + r] = tail[r];
  __syncthreads();
  for (int p0 = 0; p0 < N; p0 += FPW) {
    const int pw = min(FPW, N - p0);
    const int R = N + 1 - p0;                  // panel rows: p0 .. N (the rhs)
    for (int r = ty; r < R; r += FRG) P[r][tx] = tx < pw ? T[(size_t)(p0 + r) * ldt + p0 + tx] : 0.0;
    __syncthreads();
    for (int k = 0; k < pw; ++k) {
      if (tx > k && tx < pw && k + 1 + ty < R) {
        const double u = P[tx][k] * (1.0 / P[k][k]);
        const double* pk = &P[k + 1 + ty][k];
        double* pt = &P[k + 1 + ty][tx];
        for (int i = k + 1 + ty; i < R; i += FRG, pk += FRG * 65, pt += FRG * 65) *pt = fma(-*pk, u, *pt);
      }
      __syncthreads();
    }
    if (tid < pw) {
      const double pv = P[tid][tid];
      piv[p0 + tid] = pv;
      zr[p0 + tid] = P[R - 1][tid];
      pinv[tid] = 1.0 / pv;
    }
    __syncthreads();
    for (int jb = p0 + FPW; jb < N; jb += FPW) {
      const int j = jb + tx;
#pragma unroll
      for (int kk = 0; kk < FPW / FRG; ++kk) {
        const int k = ty * (FPW / FRG) + kk;
        Us[tx][k] = (j < N && k < pw) ? P[j - p0][k] * pinv[k] : 0.0;
      }
      __syncthreads();
      for (int i = jb + ty; i <= N; i += FRG) {
        const double* pr = P[i - p0];
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 16
        for (int k = 0; k < FPW; k += 2) {
          a0 = fma(pr[k], Us[tx][k], a0);
          a1 = fma(pr[k + 1], Us[tx][k + 1], a1);
        }
        if (j < N && j <= i) T[(size_t)i * ldt + j] -= a0 + a1;
      }
      __syncthreads();
    }
  }
  double ld = 0.0, zz = 0.0;
  int first_bad = 0x7fffffff;
  if (tid < N) {
    const double pv = piv[tid], z = zr[tid];
    if (!(pv > 0.0)) first_bad = tid + 1;
    ld = log(pv);
    zz = z * z / pv;
  }
  ld = block_sum(ld, red);
  zz = block_sum(zz, red);
  atomicMin(&bad, first_bad);
  __syncthreads();
  if (tid == 0) {
    const double s_quad = tail[N], s_logdet = tail[N + 1];
    const double quad = s_quad - c * zz;
    const double logdet = s_logdet + ld;
    const double sign = (((long long)tail[N + 3]) & 1) ? -1.0 : 1.0;
    sc->s_quad = s_quad; sc->s_logdet = s_logdet;
    sc->quad = quad; sc->logdet = logdet; sc->sign = sign;
    sc->nonpos = (int)tail[N + 2];
    sc->info = bad == 0x7fffffff ? 0 : bad;
    sc->loglike = -0.5 * (quad + sign * logdet);
  }
}

// the Cholesky finish for c.N: N <= 64, a panel kernel while a panel fits in shared memory, else global
static int launch_finish(const BatchWs& ws, int K, int N, cudaStream_t s) {
  if (N <= 64) {
    bg_finish64_kernel<<<K, 64 * FRG, 0, s>>>(ws, N);
    IMB_LAUNCHED();
    return IMB_OK;
  }
  const int ldt = N + 1;
  const size_t psm = finish_panel_smem(N);
  if (psm <= 220 * 1024 && N <= 1024) {
    static size_t attr_panel = 0;
    if (psm > attr_panel) {
      IMB_CUDA(cudaFuncSetAttribute(bg_finish_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
      attr_panel = psm;
    }
    bg_finish_panel_kernel<<<K, 1024, psm, s>>>(ws, N, ldt);
    IMB_LAUNCHED();
    return IMB_OK;
  }
  const size_t smem = sizeof(double) * (size_t)(N + 1) * ldt;
  const int use_smem = smem <= 200 * 1024;
  static size_t attr_bytes = 0;
  if (use_smem && smem > 48 * 1024 && smem > attr_bytes) {
    IMB_CUDA(cudaFuncSetAttribute(bg_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  bg_finish_kernel<<<K, 256, use_smem ? smem : 0, s>>>(ws, N, use_smem, ldt);
  IMB_LAUNCHED();
  return IMB_OK;
}

// ---- host ---------------------------------------------------------------------------------
struct BatchCtx {
  int K, N, m, slices, dparts, nn;
  int cols;        // columns per Gram slice
  bool use64;      // N <= 64 and 16-byte aligned rows: the TMA-ring Gram kernel
  RowLayout L;
  BatchPtrs p;
  BatchWs ws;
  LikeScalars* sc0;
};

// Deterministic carving: the three phases of one evaluation see the same pointers as long as no
// other library call grows the workspace in between.  G / P2 may be caller-owned (contiguous
// [K, nn] and [K, nn + N + 4]) so that torch.distributed can all-reduce them.
// phase: 0 = a self-contained evaluation, 1 = first phase of a sharded one (both claim the workspace),
// 2 / 3 = later phases (the workspace must still be the one phase 1 claimed: IMB_ESTATE otherwise)
static int make_ctx(BatchCtx* c, int K, const double* const* sims, int N, int m, RowLayout L,
                    const double* const* meas, const double* const* var, double* G_ext, double* P2_ext,
                    int phase = 0) {
  if (K < 1 || K > KMAX) return IMB_EINVAL;
  c->K = K; c->N = N; c->m = m; c->L = L;
  c->nn = N * N;
  c->dparts = grid_blocks(m, 256, 1);
  // slice width: a function of m only, so a key's summation order never depends on the batch
  c->use64 = N <= 64 && m % 2 == 0 && L.bstride % 2 == 0 && L.rstride % 2 == 0;
  for (int k = 0; c->use64 && sims && k < K; ++k)
    if (reinterpret_cast<uintptr_t>(sims[k]) % 16 != 0) c->use64 = false;
  c->cols = !c->use64 ? (m >= 8192 ? 512 : GCOLS) : (m >= 32768 ? 1024 : (m >= 8192 ? 512 : GCOLS));
  c->slices = (m + c->cols - 1) / c->cols;
  const size_t max_slices = (m + GCOLS - 1) / GCOLS;     // sizing is the same for both kernels
  const size_t nn = c->nn, p2 = nn + N + 4;
  const size_t tn = (size_t)(N + 1) * (N + 1);
  const size_t partial_elems = max_slices * nn > tn ? max_slices * nn : tn;
  const size_t per_key = 256 * 12 + sizeof(LikeScalars) +
                         sizeof(double) * ((size_t)3 * m + partial_elems + nn + p2 + max_slices * N +
                                           4 * (size_t)c->dparts);
  if (phase >= 2 && (state().phase_epoch != state().ws_epoch || per_key * K + 512 > state().ws_bytes))
    return IMB_ESTATE;
  IMB_TRY(ensure_workspace(per_key * K + 512));
  if (phase < 2) state().phase_epoch = ++state().ws_epoch;
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
    w.ypart = cv.take<double>(max_slices * N);
    w.dpart = cv.take<double>(4 * (size_t)c->dparts);
  }
  return IMB_OK;
}

static int gram64_attr() {
  static bool done = false;
  if (done) return IMB_OK;
  IMB_CUDA(cudaFuncSetAttribute(bg_gram64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)GRAM64_SMEM));
  IMB_CUDA(cudaFuncSetAttribute(bg_gram64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)GRAM64_SMEM));
  done = true;
  return IMB_OK;
}

static int phase1(const BatchCtx& c, cudaStream_t s) {
  const int nb = (c.N + TB - 1) / TB;
  bg_colmean_kernel<<<dim3((c.m + 255) / 256, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, c.L);
  IMB_LAUNCHED();
  if (c.use64) {
    IMB_TRY(gram64_attr());
    bg_gram64_kernel<false><<<dim3(c.slices, c.K), 256, GRAM64_SMEM, s>>>(c.p, c.ws, c.N, c.m, c.cols, c.L);
    IMB_LAUNCHED();
    bg_reduce64_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 0, c.N);
    IMB_LAUNCHED();
    return IMB_OK;
  }
  bg_gram_kernel<false><<<dim3(c.slices, nb * nb, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, c.cols, c.L);
  IMB_LAUNCHED();
  bg_reduce_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 0, c.N);
  IMB_LAUNCHED();
  return IMB_OK;
}

static int phase2(const BatchCtx& c, double m_total, int estimator, cudaStream_t s) {
  const int nb = (c.N + TB - 1) / TB;
  bg_oas_kernel<<<c.K, c.N > 64 ? 1024 : 256, 0, s>>>(c.ws, c.N, m_total, estimator);
  IMB_LAUNCHED();
  bg_dvec_kernel<<<dim3(c.dparts, c.K), 256, 0, s>>>(c.p, c.ws, c.m);
  IMB_LAUNCHED();
  if (c.use64) {                       // Gw and y = U (w o d) in one pass over the maps
    IMB_TRY(gram64_attr());
    bg_gram64_kernel<true><<<dim3(c.slices, c.K), 256, GRAM64_SMEM, s>>>(c.p, c.ws, c.N, c.m, c.cols, c.L);
    IMB_LAUNCHED();
    bg_reduce64_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 1, c.N);
    IMB_LAUNCHED();
  } else {
    bg_gram_kernel<true><<<dim3(c.slices, nb * nb, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, c.cols, c.L);
    IMB_LAUNCHED();
    bg_reduce_kernel<<<dim3((c.nn + 255) / 256, c.K), 256, 0, s>>>(c.ws, c.nn, c.slices, 1, c.N);
    IMB_LAUNCHED();
    bg_wdot_kernel<<<dim3(c.slices, c.K), 256, 0, s>>>(c.p, c.ws, c.N, c.m, c.cols, c.L);
    IMB_LAUNCHED();
  }
  bg_pack2_kernel<<<c.K, 256, 0, s>>>(c.ws, c.N, c.slices, c.dparts);
  IMB_LAUNCHED();
  return IMB_OK;
}

static int phase3(const BatchCtx& c, LikeScalars* h_out, cudaStream_t s) {
  IMB_TRY(launch_finish(c.ws, c.K, c.N, s));
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
  IMB_TRY(launch_finish(c.ws, c.K, c.N, s));
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
  IMB_TRY(make_ctx(&c, K, nullptr, N, m_local, L, nullptr, nullptr, G, P2, 3));
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
  IMB_TRY(make_ctx(&c, K, sims, N, m_local, L, meas, nullptr, G, nullptr, 1));
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
  IMB_TRY(make_ctx(&c, K, sims, N, m_local, L, meas, var, G, P2, 2));
  return phase2(c, (double)m_total, estimator, (cudaStream_t)stream);
}

int imb_gram_phase3(int K, int N, int m_local, double* G, double* P2, double* out_loglike,
                    int32_t* out_info, void* stream) {
  IMB_TRY(check_init());
  if (K < 1 || K > KMAX || N < 1 || m_local < 1 || !G || !P2 || !out_loglike) return IMB_EINVAL;
  const RowLayout L{N, 0, (long long)m_local};
  BatchCtx c;
  IMB_TRY(make_ctx(&c, K, nullptr, N, m_local, L, nullptr, nullptr, G, P2, 3));
  LikeScalars h[KMAX];
  IMB_TRY(phase3(c, h, (cudaStream_t)stream));
  unpack(h, K, out_loglike, out_info, nullptr);
  return IMB_OK;
}

}  // extern "C"
