// Internal interfaces shared by likelihood.cu and dense.cu (host launchers only: kernels stay
// private to their translation unit, so no relocatable device code is needed).
#pragma once
#include "common.cuh"

namespace imb {

// scalars shared between kernels of one likelihood evaluation
struct LikeScalars {
  double trs, trs2, rho, delta, c;   // tr S, tr S^2, shrinkage, rho*trs/m, (1-rho)/N
  double s_quad, s_logdet;           // sum d^2/D, sum log(2 pi |D|)
  double sign;                       // sign of the determinant
  double loglike, quad, logdet;
  int info, nonpos;
};

int grid_blocks(long long work, int threads, int per_sm);
int gram_slices(int m);
int launch_colmean(const double* X, int N, int m, const double* meas, double* mean, double* diff,
                   cudaStream_t s);
int launch_colvar(const double* X, int N, int m, double* var, cudaStream_t s);
// G = U U^T (optionally column-weighted) of the centred ensemble; partial sized gram_slices(m)*N*N
int launch_gram(const double* X, int N, int m, const double* mean, const double* w,
                double* partial, double* G, cudaStream_t s);
int launch_oas_scalars(const double* G, int N, int m, int estimator, LikeScalars* sc, cudaStream_t s);
int read_scalars(const LikeScalars* d_sc, LikeScalars* h_out, cudaStream_t s);

// Dense route of EnsembleLikelihood / SimpleLikelihood (dense.cu).
//   ensemble = 1: A = meas_cov + cov_est(sims);  ensemble = 0: A = meas_cov (dense only).
int loglike_dense(const double* sims, int N, int m, const double* meas, const double* cov,
                  int cov_kind, int estimator, int use_trace_approximation, int ensemble,
                  double* out_loglike, int32_t* out_info, cudaStream_t s);

// Batched Gram route (gram_batch.cu): K keys of identical shape, one read-back.
int loglike_gram_batch(int K, const double* const* sims, int N, int m, const double* const* meas,
                       const double* const* var, int estimator, double* out, int32_t* info,
                       int* fallback, cudaStream_t s);

#ifdef __CUDACC__
// ---- FP64 tensor-core tile core shared by the dense route (dense.cu) and the batched Gram
// route (gram_batch.cu) ------------------------------------------------------------------
constexpr int TB = 64;       // output tile edge / Cholesky panel width
constexpr int KC = 16;       // k-chunk staged in shared memory
constexpr int LDS_ = TB + 4; // smem leading dimension: conflict-free DMMA fragment loads

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One 64x64 tile of acc = A * B^T over k in [0, K): warp w owns rows 8w..8w+7, all 8 column
// blocks.  Operand element getters are functors so the same core serves SYRK and Cholesky.
template <typename LoadA, typename LoadB>
__device__ __forceinline__ void tile_mma(int K, LoadA loadA, LoadB loadB, double (&acc)[8][2],
                                         double (*sA)[LDS_], double (*sB)[LDS_]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  for (int k0 = 0; k0 < K; k0 += KC) {
    for (int e = threadIdx.x; e < TB * KC; e += 256) {
      int i, k;
      loadA.index(e, i, k);
      sA[k][i] = (k0 + k < K) ? loadA(i, k0 + k) : 0.0;
      loadB.index(e, i, k);
      sB[k][i] = (k0 + k < K) ? loadB(i, k0 + k) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
      const double a = sA[kk + fk][warp * 8 + fr];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const double b = sB[kk + fk][nb * 8 + fr];
        dmma884(acc[nb][0], acc[nb][1], a, b);
      }
    }
    __syncthreads();
  }
}

#endif  // __CUDACC__
}  // namespace imb
