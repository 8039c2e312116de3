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
}  // namespace imb
