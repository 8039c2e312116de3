// Stage 3b: the dense covariance route (dense m x m measurement covariances; IMB_ROUTE_AUTO takes the Gram
// route of gram_batch.cu whenever the measurement covariance is diagonal).
//   S = U^T U / N as a tensor-core GEMM (FP64 DMMA, mma.sync m8n8k4 -- tcgen05 has no f64 kind), fused
//   with the OAS shrinkage and the "+ measurement covariance" epilogue;
//   two-level blocked right-looking Cholesky with lookahead (64-wide panels inside 256-wide block
//   columns: tile factor, register-resident triangular solve, 128 x 128 DMMA trailing updates fed by a
//   cp.async ring); the right-hand side d rides along as an extra row, so z = L^-1 d falls out of the
//   factorisation.  m = 12 288: 33.6 ms for the whole route (cuSOLVER potrf alone on the same matrix:
//   22 ms; cuBLAS DGEMM on this GPU: 35 TFLOP/s) -- profiles/r02_bench_c2_dense.json.
//   Non-symmetric or non-positive-definite input (the reference is LU throughout and its tests feed
//   np.random.rand matrices, tests/test_likelihood.py:63-113) takes a partially pivoted LU instead.
//   THE LU IS A CORRECTNESS PATH ONLY: unblocked, two launches per column, O(m) launches -- it exists so
//   that every input the reference accepts gives the reference's answer, not to be fast.
// Reference arithmetic: oracle/likelihood.py (empirical_mcov / oas_mcov / ensemble_loglike).
#include <vector>
#include "common.cuh"
#include "linalg.cuh"

namespace imb {

// ---- build A = M + cov_est(sims) --------------------------------------------------------
// operand(i,k) = X[k*m + i0+i] - mean[i0+i]  (i contiguous)
struct CenteredOperand {
  const double* X; const double* mean; int m, N, i0;
  __device__ __forceinline__ void index(int e, int& i, int& k) const { i = e & (TB - 1); k = e >> 6; }
  __device__ __forceinline__ double operator()(int i, int k) const {
    const int col = i0 + i;
    return col < m ? X[(size_t)k * m + col] - mean[col] : 0.0;
  }
};

// out[i*ld + j] = M_ij + ((1-rho) * S_ij + (i==j) * delta),  S = U^T U / N
//   cov_kind: 0 none (M = 0), 1 variance (M = diag), 2 dense
__global__ void __launch_bounds__(256) build_cov_kernel(const double* __restrict__ X, int N, int m,
                                                        const double* __restrict__ mean,
                                                        const LikeScalars* __restrict__ sc,
                                                        const double* __restrict__ M, int cov_kind,
                                                        double* __restrict__ out, int ld) {
  __shared__ double sA[KC][LDS_];
  __shared__ double sB[KC][LDS_];
  double acc[8][2];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
  const int i0 = blockIdx.y * TB, j0 = blockIdx.x * TB;
  CenteredOperand A{X, mean, m, N, i0}, B{X, mean, m, N, j0};
  tile_mma(N, A, B, acc, sA, sB);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double one_minus_rho = 1.0 - sc->rho, delta = sc->delta, n = (double)N;
  const int i = i0 + warp * 8 + (lane >> 2);
  if (i >= m) return;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = j0 + nb * 8 + 2 * (lane & 3) + h;
      if (j >= m) continue;
      const double s = __ddiv_rn(acc[nb][h], n);
      const double cov = __dadd_rn(__dmul_rn(one_minus_rho, s), (i == j) ? delta : 0.0);
      double mij = 0.0;
      if (cov_kind == IMB_COV_DENSE) mij = M[(size_t)i * m + j];
      else if (cov_kind == IMB_COV_VARIANCE) mij = (i == j) ? M[i] : 0.0;
      out[(size_t)i * ld + j] = (cov_kind == IMB_COV_NONE) ? cov : __dadd_rn(mij, cov);
    }
  }
}

// out = M (+ diag(extra)) copied into a matrix with leading dimension ld
__global__ void copy_cov_kernel(const double* __restrict__ M, int cov_kind, int m,
                                const double* __restrict__ extra_diag, double* __restrict__ out, int ld) {
  const long long total = (long long)m * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e % m);
    double v = 0.0;
    if (cov_kind == IMB_COV_DENSE) v = M[e];
    else if (cov_kind == IMB_COV_VARIANCE) v = (i == j) ? M[i] : 0.0;
    if (extra_diag && i == j) v = (cov_kind == IMB_COV_NONE) ? extra_diag[i] : __dadd_rn(v, extra_diag[i]);
    out[(size_t)i * ld + j] = v;
  }
}

__global__ void place_rhs_kernel(const double* __restrict__ d, int m, double* __restrict__ dst, int stride) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    dst[(size_t)i * stride] = d[i];
}

// flag[0] = 1 if A (ld) is not exactly symmetric
__global__ void symmetry_kernel(const double* __restrict__ A, int m, int ld, int* flag) {
  const long long total = (long long)m * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e % m);
    if (j < i && A[(size_t)i * ld + j] != A[(size_t)j * ld + i]) *flag = 1;
  }
}

// ---- blocked Cholesky (lower, in place, row-major, ld = m; rows 0..m, row m = rhs) ----------
// factor the kb x kb diagonal tile at k0; pivots[k0+k] = value under the square root.
// The tile is reduced in its L D L^T form -- column k keeps t_ik = l_ik l_kk and the update is
// T[i][j] -= t_ik (t_jk / p_k) -- so that the serial chain of a step is ONE division, no square root and
// one barrier (sqrt + divide + two barriers per step made this kernel 17 % of the factorisation);
// L = T / sqrt(p) column by column falls out at the end, in parallel.
constexpr int PRG = 16;            // row groups of the tile factor: 1024 threads = 16 row groups x 64 columns
__global__ void __launch_bounds__(64 * PRG) potf2_kernel(double* __restrict__ A, int ld, int k0, int kb,
                                                         double* __restrict__ pivots, int* __restrict__ info) {
  // the strictly upper triangle is updated along and never read; 16 row groups: a thread touches at most
  // 4 rows per step, with exact loop bounds and constant pointer strides (a step is bound by the
  // instructions the block issues)
  __shared__ double T[TB][TB + 1];
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  for (int i = ty; i < TB; i += PRG)
    T[i][tx] = (i < kb && tx < kb) ? A[(size_t)(k0 + i) * ld + k0 + tx] : 0.0;
  __syncthreads();
  for (int k = 0; k + 1 < kb; ++k) {
    if (tx > k && tx < kb && k + 1 + ty < kb) {
      const double u = T[tx][k] / T[k][k];
      // rows k + 1 + ty, + 16, ...: exact bounds, pointers stepped by a constant
      const double* pk = &T[k + 1 + ty][k];
      double* pt = &T[k + 1 + ty][tx];
      for (int i = k + 1 + ty; i < kb; i += PRG, pk += PRG * (TB + 1), pt += PRG * (TB + 1)) *pt = fma(-*pk, u, *pt);
    }
    __syncthreads();          // column k + 1 (and its pivot) is final
  }
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0x7fffffff;
  __syncthreads();
  if (tx < kb) {
    const double p = T[tx][tx];
    const double lkk = sqrt(p);
    if (ty == 0) {
      pivots[k0 + tx] = p;
      if (!(p > 0.0)) atomicMin(&s_bad, k0 + tx + 1);
    }
    for (int i = tx + ty; i < kb; i += PRG)
      A[(size_t)(k0 + i) * ld + k0 + tx] = (i == tx) ? lkk : T[i][tx] / lkk;
  }
  __syncthreads();
  // info[0] keeps the FIRST non-positive pivot of the whole factorisation (LAPACK's info)
  if (tid == 0 && s_bad != 0x7fffffff && *info == 0) *info = s_bad;
}

// rows i in [k0+kb, nrows): A[i, k0:k0+kb] <- A[i, k0:k0+kb] * L11^-T ; 64 rows per block, a row per
// thread.  The row lives in registers (the substitution is fully unrolled), L11 is read as shared-memory
// broadcasts, and every dot product runs on two independent FMA chains (DFMA latency is 8 cycles and a
// block is two warps: the chain length is the run time; the diagonal is applied as a reciprocal for the
// same reason).  kb < 64: L11 is padded with the identity.
__global__ void __launch_bounds__(64) trsm_kernel(double* __restrict__ A, int ld, int k0, int kb,
                                                  int nrows) {
  extern __shared__ __align__(16) double sm[];
  double (*L)[TB] = reinterpret_cast<double (*)[TB]>(sm);
  double (*R)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(sm + TB * TB);
  const int tid = threadIdx.x;
  const int r0 = k0 + kb + blockIdx.x * TB;
  __shared__ double rdiag[TB];
  // row e / 64 per iteration, 64 contiguous doubles across the block; 16 independent loads in flight
#pragma unroll 16
  for (int e = tid; e < TB * TB; e += 64) {
    const int i = e >> 6, j = e & 63;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < kb && j <= i) v = A[(size_t)(k0 + i) * ld + k0 + j];
    L[i][j] = v;
  }
#pragma unroll 16
  for (int e = tid; e < TB * TB; e += 64) {
    const int i = e >> 6, j = e & 63;
    R[i][j] = (r0 + i < nrows && j < kb) ? A[(size_t)(r0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  rdiag[tid] = 1.0 / L[tid][tid];          // 64 divisions off the substitution's serial chain
  __syncthreads();
  double r[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) r[c] = R[tid][c];
#pragma unroll
  for (int c = 0; c < TB; ++c) {
    double s0 = r[c], s1 = 0.0;
#pragma unroll
    for (int j = 0; j + 1 < c; j += 2) {
      s0 = fma(-r[j], L[c][j], s0);
      s1 = fma(-r[j + 1], L[c][j + 1], s1);
    }
    if (c & 1) s0 = fma(-r[c - 1], L[c][c - 1], s0);
    r[c] = (s0 + s1) * rdiag[c];
  }
#pragma unroll
  for (int c = 0; c < TB; ++c) R[tid][c] = r[c];
  __syncthreads();
  for (int e = tid; e < TB * TB; e += 64) {
    const int i = e >> 6, j = e & 63;
    if (r0 + i < nrows && j < kb) A[(size_t)(r0 + i) * ld + k0 + j] = R[i][j];
  }
}

// operand(i,k) = A[(r0+i)*ld + k0+k]  (k contiguous)
struct PanelOperand {
  const double* A; int ld, r0, k0, nrows;
  __device__ __forceinline__ void index(int e, int& i, int& k) const { k = e & (KC - 1); i = e >> 4; }
  __device__ __forceinline__ double operator()(int i, int k) const {
    return (r0 + i < nrows) ? A[(size_t)(r0 + i) * ld + k0 + k] : 0.0;
  }
};

// trailing update: A[i][j] -= sum_k A[i][k0+k] A[j][k0+k], i in [t0, nrows), j in [t0, m), tiles j<=i
__global__ void __launch_bounds__(256) chol_update_kernel(double* __restrict__ A, int ld, int k0, int kb,
                                                          int t0, int nrows, int m) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ double sA[KC][LDS_];
  __shared__ double sB[KC][LDS_];
  double acc[8][2];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
  const int i0 = t0 + blockIdx.y * TB, j0 = t0 + blockIdx.x * TB;
  PanelOperand Pa{A, ld, i0, k0, nrows}, Pb{A, ld, j0, k0, m};
  tile_mma(kb, Pa, Pb, acc, sA, sB);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = i0 + warp * 8 + (lane >> 2);
  if (i >= nrows) return;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = j0 + nb * 8 + 2 * (lane & 3) + h;
      if (j < m && j <= i) A[(size_t)i * ld + j] -= acc[nb][h];
    }
}

// The same update on 128 x 128 tiles for even ld and K a multiple of 16 (every panel but a ragged last
// one): the two 128 x 16 operand chunks of a k-step arrive by 16-byte cp.async (LDGSTS) into a double
// ring (3 stages) whose rows keep the k index contiguous, exactly as the panel lies in memory -- no transposing
// 8-byte stores (the 64-tile kernel above spends its L1 data pipe on them: ncu r02, 77 % of the
// wavefront peak at 43 % DMMA) -- and a row stride of 20 doubles makes the DMMA fragment reads
// conflict-free.  Warp (wi, wj) of the 4 x 4 warp grid owns 32 rows x 32 columns: 4 A fragments and
// 4 B fragments feed 16 DMMAs per 4-deep k-step (8 shared loads per 16 tensor instructions instead
// of 9 per 8); 16 warps per SM keep the tensor pipe fed across the fixed issue gaps of a DMMA chain
// (8 warps of 32 x 64: 60 % DMMA-active).
constexpr int UT = 128;            // tile edge
constexpr int UK = 16;             // k-chunk
constexpr int UKS = UK + 4;        // shared row stride in doubles
constexpr int UST = 3;             // cp.async stages
constexpr size_t UPD_SMEM = sizeof(double) * 2 * UST * UT * UKS;   // operands x stages

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

constexpr int UTHREADS = 512;      // 16 warps: 4 x 4 warp grid, 32 x 32 outputs per warp
__global__ void __launch_bounds__(UTHREADS, 1) chol_update128_kernel(double* __restrict__ A, int ld, int k0, int K,
                                                                     int t0, int nrows, int jmax) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  extern __shared__ __align__(16) double usm[];
  double (*sA)[UT][UKS] = reinterpret_cast<double (*)[UT][UKS]>(usm);
  double (*sB)[UT][UKS] = reinterpret_cast<double (*)[UT][UKS]>(usm + UST * UT * UKS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wi = warp >> 2, wj = warp & 3, fr = lane >> 2, fk = lane & 3;
  const int i0 = t0 + bi * UT, j0 = t0 + bj * UT;
  double acc[4][4][2];
#pragma unroll
  for (int ra = 0; ra < 4; ++ra)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc[ra][nb][0] = acc[ra][nb][1] = 0.0;

  auto stage = [&](int c, int buf) {
#pragma unroll
    for (int q = 0; q < UT * 8 / UTHREADS; ++q) {
      const int p = tid + UTHREADS * q, row = p >> 3, seg = (p & 7) * 2;
      const int ri = i0 + row, rj = j0 + row;
      const bool okA = ri < nrows, okB = rj < jmax;
      // out-of-range rows are zero-filled (src-size 0); the address stays a valid one
      cp_async16(&sA[buf][row][seg], A + (size_t)(okA ? ri : t0) * ld + k0 + c * UK + seg, okA ? 16 : 0);
      cp_async16(&sB[buf][row][seg], A + (size_t)(okB ? rj : t0) * ld + k0 + c * UK + seg, okB ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // chunk c + 2 is requested right after the barrier that ends the compute of chunk c - 1 (whose buffer
  // it takes): one barrier per chunk, two chunks in flight during the DMMAs
  const int nch = K / UK;
  stage(0, 0);
  if (nch > 1) stage(1, 1);
  // the output tile is read-modify-written by the epilogue with nothing to overlap (one block per SM):
  // pull it into L2 now.  A warp's 32 x 32 outputs are 32 rows of 256 contiguous bytes.
  {
    const int i = i0 + 32 * wi + lane;
    if (i < nrows) {
      const double* row = A + (size_t)i * ld + j0 + 32 * wj;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 31));
    }
  }
  for (int c = 0; c < nch; ++c) {
    const int buf = c % UST;
    if (c + 1 < nch) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (c + 2 < nch) stage(c + 2, (c + 2) % UST);
#pragma unroll
    for (int kk = 0; kk < UK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int ra = 0; ra < 4; ++ra) a[ra] = sA[buf][32 * wi + 8 * ra + fr][kk + fk];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) b[nb] = sB[buf][32 * wj + 8 * nb + fr][kk + fk];
#pragma unroll
      for (int ra = 0; ra < 4; ++ra)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) dmma884(acc[ra][nb][0], acc[ra][nb][1], a[ra], b[nb]);
    }
  }

#pragma unroll
  for (int ra = 0; ra < 4; ++ra) {
    const int i = i0 + 32 * wi + 8 * ra + fr;
    if (i >= nrows) continue;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int j = j0 + 32 * wj + 8 * nb + 2 * fk;
      double* p = A + (size_t)i * ld + j;
      if (j + 1 < jmax && j + 1 <= i) {            // both elements in the lower triangle: one 16-byte RMW
        double2 v = *reinterpret_cast<double2*>(p);
        v.x -= acc[ra][nb][0];
        v.y -= acc[ra][nb][1];
        *reinterpret_cast<double2*>(p) = v;
      } else {
        if (j < jmax && j <= i) p[0] -= acc[ra][nb][0];
        if (j + 1 < jmax && j + 1 <= i) p[1] -= acc[ra][nb][1];
      }
    }
  }
}

// ---- partially pivoted LU on the augmented matrix [A | d] (ld = m + 1) ------------------------
__global__ void __launch_bounds__(1024) lu_pivot_kernel(double* __restrict__ A, int m, int ld, int k,
                                                        double* __restrict__ pivots, int* __restrict__ info) {
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ int prow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double best = -1.0;
  int bi = k;
  for (int i = k + tid; i < m; i += blockDim.x) {
    const double v = fabs(A[(size_t)i * ld + k]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { sval[warp] = best; sidx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (blockDim.x >> 5) ? sval[lane] : -1.0;
    bi = lane < (blockDim.x >> 5) ? sidx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) prow = bi;
  }
  __syncthreads();
  const int p = prow;
  if (p != k) {
    for (int j = tid; j <= m; j += blockDim.x) {
      const double a = A[(size_t)k * ld + j], b = A[(size_t)p * ld + j];
      A[(size_t)k * ld + j] = b;
      A[(size_t)p * ld + j] = a;
    }
  }
  __syncthreads();
  const double ukk = A[(size_t)k * ld + k];
  if (tid == 0) {
    pivots[k] = ukk;
    if (p != k) info[1] ^= 1;                       // swap parity
    if (ukk == 0.0 && info[0] == 0) info[0] = k + 1;  // exactly singular
  }
  for (int i = k + 1 + tid; i < m; i += blockDim.x) A[(size_t)i * ld + k] /= ukk;
}

__global__ void __launch_bounds__(256) lu_update_kernel(double* __restrict__ A, int m, int ld, int k) {
  // A[i][j] -= l_i * u_j for i in (k, m), j in (k, m]  (column m = rhs)
  const int j = k + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j > m) return;
  const double ukj = A[(size_t)k * ld + j];
  for (int i = k + 1 + blockIdx.y; i < m; i += gridDim.y)
    A[(size_t)i * ld + j] -= A[(size_t)i * ld + k] * ukj;
}

// back substitution U x = y (y = column m), then quad = d . x ; single block
__global__ void __launch_bounds__(1024) lu_backsolve_kernel(double* __restrict__ A, int m, int ld,
                                                            const double* __restrict__ d,
                                                            LikeScalars* __restrict__ sc) {
  __shared__ double red[32];
  __shared__ double xk;
  const int tid = threadIdx.x;
  for (int k = m - 1; k >= 0; --k) {
    if (tid == 0) {
      xk = A[(size_t)k * ld + m] / A[(size_t)k * ld + k];
      A[(size_t)k * ld + m] = xk;
    }
    __syncthreads();
    const double x = xk;
    for (int i = tid; i < k; i += blockDim.x) A[(size_t)i * ld + m] -= A[(size_t)i * ld + k] * x;
    __syncthreads();
  }
  Comp q;
  for (int i = tid; i < m; i += blockDim.x) q.add(d[i] * A[(size_t)i * ld + m]);
  const double v = block_sum(q.value(), red);
  if (tid == 0) sc->quad = v;
}

// quad = || row m of the Cholesky workspace ||^2 ; single block
__global__ void __launch_bounds__(1024) chol_quad_kernel(const double* __restrict__ z, int m,
                                                         LikeScalars* __restrict__ sc) {
  __shared__ double red[32];
  Comp q;
  for (int i = threadIdx.x; i < m; i += blockDim.x) q.add(z[i] * z[i]);
  const double v = block_sum(q.value(), red);
  if (threadIdx.x == 0) sc->quad = v;
}

// logdet = sum_k log((|p_k| * 2) * pi) in a fixed order, sign = parity * prod sign(p_k);
// then loglike = -1/2 (quad + sign * logdet).  use_abs = 0 mirrors numpy's log of the raw
// value (trace approximation, sign forced to +1).  Single block.
__global__ void __launch_bounds__(1024) logdet_finish_kernel(const double* __restrict__ piv, int m,
                                                             int use_abs, const int* __restrict__ info,
                                                             LikeScalars* __restrict__ sc, int have_quad) {
  __shared__ double red[32];
  Comp l;
  double neg = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double p = piv[i];
    l.add(log(__dmul_rn(__dmul_rn(use_abs ? fabs(p) : p, 2.0), 3.141592653589793)));
    if (p < 0.0) neg += 1.0;
  }
  const double ld = block_sum(l.value(), red);
  const double ng = block_sum(neg, red);
  if (threadIdx.x == 0) {
    double sign = 1.0;
    if (use_abs) {
      long long flips = (long long)ng + (info ? info[1] : 0);
      sign = (flips & 1) ? -1.0 : 1.0;
    }
    sc->logdet = ld;
    sc->sign = sign;
    if (have_quad) sc->loglike = -0.5 * (sc->quad + sign * ld);
  }
}

__global__ void add_vec_kernel(const double* __restrict__ a, const double* __restrict__ b, int m,
                               double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    out[i] = __dadd_rn(a[i], b[i]);
}

__global__ void diag_of_kernel(const double* __restrict__ M, int m, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    out[i] = M[(size_t)i * m + i];
}

// ---- host orchestration ------------------------------------------------------------------
struct DenseWs {
  LikeScalars* sc;
  int* info;          // [0] first bad pivot (1-based), [1] swap parity, [2] asymmetry flag
  double *mean, *diff, *xvar, *tvec, *pivots, *G, *partial, *A;
};

static int carve_dense(int N, int m, DenseWs* w) {
  const int slices = gram_slices(m);
  const size_t a_elems = (size_t)(m + 1) * (m + 1);
  const size_t need = 256 * 16 + sizeof(LikeScalars) + 64 +
                      sizeof(double) * ((size_t)5 * m + (size_t)(slices + 1) * N * N + a_elems);
  IMB_TRY(ensure_workspace(need));
  ++state().ws_epoch;
  Carver c(state().ws);
  w->sc = c.take<LikeScalars>(1);
  w->info = c.take<int>(4);
  w->mean = c.take<double>(m);
  w->diff = c.take<double>(m);
  w->xvar = c.take<double>(m);
  w->tvec = c.take<double>(m);
  w->pivots = c.take<double>(m);
  w->G = c.take<double>((size_t)N * N);
  w->partial = c.take<double>((size_t)slices * N * N);
  w->A = c.take<double>(a_elems);
  return IMB_OK;
}

// writes A (leading dimension ld) = meas_cov + estimator covariance of sims
static int build_matrix(const double* sims, int N, int m, const double* cov, int cov_kind,
                        int estimator, int ensemble, DenseWs& w, int ld, cudaStream_t s) {
  if (!ensemble) {
    copy_cov_kernel<<<grid_blocks((long long)m * m, 256, 8), 256, 0, s>>>(cov, cov_kind, m, nullptr, w.A, ld);
    IMB_LAUNCHED();
  } else if (estimator == IMB_EST_DIAGONAL) {
    copy_cov_kernel<<<grid_blocks((long long)m * m, 256, 8), 256, 0, s>>>(cov, cov_kind, m, w.xvar, w.A, ld);
    IMB_LAUNCHED();
  } else {
    const int nt = (m + TB - 1) / TB;
    build_cov_kernel<<<dim3(nt, nt), 256, 0, s>>>(sims, N, m, w.mean, w.sc, cov, cov_kind, w.A, ld);
    IMB_LAUNCHED();
  }
  return IMB_OK;
}

// trailing update A[i][j] -= sum_{k < K} A[i][k0+k] A[j][k0+k] for i in [t0, nrows), j in [t0, jmax), j <= i
static int launch_update(double* A, int ld, int k0, int K, int t0, int nrows, int jmax, cudaStream_t s) {
  if (t0 >= jmax || t0 >= nrows || K <= 0) return IMB_OK;
  if (ld % 2 == 0 && K % UK == 0) {
    const int nty = (nrows - t0 + UT - 1) / UT, ntx = (jmax - t0 + UT - 1) / UT;
    chol_update128_kernel<<<dim3(ntx, nty), UTHREADS, UPD_SMEM, s>>>(A, ld, k0, K, t0, nrows, jmax);
  } else {
    const int nty = (nrows - t0 + TB - 1) / TB, ntx = (jmax - t0 + TB - 1) / TB;
    chol_update_kernel<<<dim3(ntx, nty), 256, 0, s>>>(A, ld, k0, K, t0, nrows, jmax);
  }
  IMB_LAUNCHED();
  return IMB_OK;
}

// Two-level right-looking factorisation with one block column of lookahead.
//   * 64-wide panels (tile factor + triangular solve) update only the rest of their 256-wide block
//     column; the trailing matrix is read and written once per BLOCK COLUMN, by a rank-256 update (a
//     quarter of the passes over the matrix, and 256 k-steps of DMMA per tile visit).
//   * The rank-256 update of block column b is split: the next block column first ("head"), then
//     everything to its right ("rest").  The panel chain of block column b + 1 (latency-bound: one small
//     block, then a few dozen) runs on a second, high-priority stream as soon as the head is done and
//     overlaps the rest update, which only touches columns further right.
constexpr int OB = 256;
struct Lookahead {
  cudaStream_t panel = nullptr;
  std::vector<cudaEvent_t> ev;
  int get(size_t n) {
    if (!panel) {
      int lo = 0, hi = 0;
      IMB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      IMB_CUDA(cudaStreamCreateWithPriority(&panel, cudaStreamNonBlocking, hi));
    }
    while (ev.size() < n) {
      cudaEvent_t e;
      IMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    return IMB_OK;
  }
};

static int cholesky(DenseWs& w, int m, cudaStream_t s) {
  const int ld = m, nrows = m + 1;
  static bool attr_set = false;
  static Lookahead la;
  const size_t trsm_smem = sizeof(double) * (TB * TB + TB * (TB + 1));
  if (!attr_set) {
    IMB_CUDA(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem));
    IMB_CUDA(cudaFuncSetAttribute(chol_update128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)UPD_SMEM));
    attr_set = true;
  }
  const int nblk = (m + OB - 1) / OB;
  IMB_TRY(la.get(2 * (size_t)nblk + 1));
  cudaStream_t sp = la.panel;
  // events: [2 b] panel chain of block column b done (on sp), [2 b + 1] head update of b done (on s)
  IMB_CUDA(cudaEventRecord(la.ev[2 * nblk], s));             // everything queued on s before this call
  IMB_CUDA(cudaStreamWaitEvent(sp, la.ev[2 * nblk], 0));
  for (int b = 0; b < nblk; ++b) {
    const int K0 = b * OB;
    const int kend = (m - K0) < OB ? m : K0 + OB;
    if (b > 0) IMB_CUDA(cudaStreamWaitEvent(sp, la.ev[2 * (b - 1) + 1], 0));
    for (int k0 = K0; k0 < kend; k0 += TB) {
      const int kb = (m - k0) < TB ? (m - k0) : TB;
      potf2_kernel<<<1, 64 * PRG, 0, sp>>>(w.A, ld, k0, kb, w.pivots, w.info);
      IMB_LAUNCHED();
      const int t0 = k0 + kb;
      if (nrows - t0 <= 0) break;
      trsm_kernel<<<(nrows - t0 + TB - 1) / TB, 64, trsm_smem, sp>>>(w.A, ld, k0, kb, nrows);
      IMB_LAUNCHED();
      IMB_TRY(launch_update(w.A, ld, k0, kb, t0, nrows, kend, sp));     // inside the block column
    }
    IMB_CUDA(cudaEventRecord(la.ev[2 * b], sp));
    IMB_CUDA(cudaStreamWaitEvent(s, la.ev[2 * b], 0));
    const int head_end = (m - kend) < OB ? m : kend + OB;
    IMB_TRY(launch_update(w.A, ld, K0, kend - K0, kend, nrows, head_end, s));      // the next block column
    IMB_CUDA(cudaEventRecord(la.ev[2 * b + 1], s));
    IMB_TRY(launch_update(w.A, ld, K0, kend - K0, head_end, nrows, m, s));         // everything to its right
  }
  return IMB_OK;
}

static int lu(DenseWs& w, int m, cudaStream_t s) {
  const int ld = m + 1;
  for (int k = 0; k < m; ++k) {
    lu_pivot_kernel<<<1, m - k >= 4096 ? 1024 : 256, 0, s>>>(w.A, m, ld, k, w.pivots, w.info);
    IMB_LAUNCHED();
    const int cols = m - k;  // columns k+1 .. m inclusive
    if (k + 1 < m) {
      const int rows = m - k - 1;
      const int gy = rows < 64 ? rows : (rows < 2048 ? 64 : 256);
      lu_update_kernel<<<dim3((cols + 255) / 256, gy), 256, 0, s>>>(w.A, m, ld, k);
      IMB_LAUNCHED();
    }
  }
  return IMB_OK;
}

int loglike_dense(const double* sims, int N, int m, const double* meas, const double* cov,
                  int cov_kind, int estimator, int use_trace_approximation, int ensemble,
                  double* out_loglike, int32_t* out_info, cudaStream_t s) {
  DenseWs w;
  IMB_TRY(carve_dense(N, m, &w));
  IMB_CUDA(cudaMemsetAsync(w.info, 0, 4 * sizeof(int), s));
  IMB_TRY(launch_colmean(sims, N, m, meas, w.mean, w.diff, s));
  if (ensemble) {
    if (estimator == IMB_EST_DIAGONAL || use_trace_approximation) IMB_TRY(launch_colvar(sims, N, m, w.xvar, s));
    if (estimator != IMB_EST_DIAGONAL) {
      IMB_TRY(launch_gram(sims, N, m, w.mean, nullptr, w.partial, w.G, s));
      IMB_TRY(launch_oas_scalars(w.G, N, m, estimator, w.sc, s));
    }
  }
  // symmetric input -> try Cholesky first (A = M + symmetric PSD estimate)
  bool try_chol = true;
  if (cov_kind == IMB_COV_DENSE) {
    symmetry_kernel<<<grid_blocks((long long)m * m, 256, 8), 256, 0, s>>>(cov, m, m, w.info + 2);
    IMB_LAUNCHED();
    int* h = static_cast<int*>(state().pinned);
    IMB_CUDA(cudaMemcpyAsync(h, w.info + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
    IMB_CUDA(cudaStreamSynchronize(s));
    try_chol = (*h == 0);
  }
  int bits = IMB_INFO_ROUTE_DENSE;
  int lapack_info = 0;
  bool done = false;
  if (try_chol) {
    IMB_TRY(build_matrix(sims, N, m, cov, cov_kind, estimator, ensemble, w, m, s));
    place_rhs_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(w.diff, m, w.A + (size_t)m * m, 1);
    IMB_LAUNCHED();
    IMB_TRY(cholesky(w, m, s));
    int* h = static_cast<int*>(state().pinned);
    IMB_CUDA(cudaMemcpyAsync(h, w.info, sizeof(int), cudaMemcpyDeviceToHost, s));
    IMB_CUDA(cudaStreamSynchronize(s));
    if (*h == 0) {
      chol_quad_kernel<<<1, 1024, 0, s>>>(w.A + (size_t)m * m, m, w.sc);
      IMB_LAUNCHED();
      bits |= IMB_INFO_CHOLESKY;
      done = true;
    } else {
      lapack_info = *h;   // not positive definite: fall through to LU
      IMB_CUDA(cudaMemsetAsync(w.info, 0, 4 * sizeof(int), s));
    }
  }
  if (!done) {
    IMB_TRY(build_matrix(sims, N, m, cov, cov_kind, estimator, ensemble, w, m + 1, s));
    place_rhs_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(w.diff, m, w.A + m, m + 1);
    IMB_LAUNCHED();
    IMB_TRY(lu(w, m, s));
    lu_backsolve_kernel<<<1, 1024, 0, s>>>(w.A, m, m + 1, w.diff, w.sc);
    IMB_LAUNCHED();
    bits |= IMB_INFO_LU;
  }
  if (ensemble && use_trace_approximation) {
    // logdet ~ sum log(2 pi (meas_var + X.var(0))), sign = +1 (ensemble_likelihood.py:100-103)
    if (cov_kind == IMB_COV_DENSE) {
      diag_of_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(cov, m, w.tvec);
      IMB_LAUNCHED();
      add_vec_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(w.tvec, w.xvar, m, w.tvec);
    } else {
      add_vec_kernel<<<grid_blocks(m, 256, 4), 256, 0, s>>>(cov, w.xvar, m, w.tvec);
    }
    IMB_LAUNCHED();
    logdet_finish_kernel<<<1, 1024, 0, s>>>(w.tvec, m, 0, nullptr, w.sc, 1);
  } else {
    logdet_finish_kernel<<<1, 1024, 0, s>>>(w.pivots, m, 1, w.info, w.sc, 1);
  }
  IMB_LAUNCHED();
  LikeScalars h;
  IMB_TRY(read_scalars(w.sc, &h, s));
  int* hi = static_cast<int*>(state().pinned);
  IMB_CUDA(cudaMemcpyAsync(hi, w.info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  IMB_CUDA(cudaStreamSynchronize(s));
  if (!done && hi[0] != 0) {
    if (out_info) { out_info[0] = bits; out_info[1] = hi[0]; }
    return IMB_ESINGULAR;
  }
  if (h.sign < 0) bits |= IMB_INFO_NEGATIVE_DET;
  *out_loglike = h.loglike;
  if (out_info) { out_info[0] = bits; out_info[1] = lapack_info; }
  return IMB_OK;
}

}  // namespace imb

using namespace imb;

extern "C" int imb_mcov(const double* sims, int N, int m, int estimator, double* out_mean,
                        double* out_cov, double* out_rho, void* stream) {
  IMB_TRY(check_init());
  cudaStream_t s = (cudaStream_t)stream;
  if (!sims || !out_mean || !out_cov || N < 1 || m < 1) return IMB_EINVAL;
  if (estimator < IMB_EST_EMPIRICAL || estimator > IMB_EST_DIAGONAL) return IMB_EINVAL;
  const int slices = gram_slices(m);
  IMB_TRY(ensure_workspace(256 * 8 + sizeof(LikeScalars) +
                           sizeof(double) * ((size_t)m + (size_t)(slices + 1) * N * N)));
  ++state().ws_epoch;
  Carver c(state().ws);
  LikeScalars* sc = c.take<LikeScalars>(1);
  double* xvar = c.take<double>(m);
  double* G = c.take<double>((size_t)N * N);
  double* partial = c.take<double>((size_t)slices * N * N);
  IMB_TRY(launch_colmean(sims, N, m, nullptr, out_mean, nullptr, s));
  LikeScalars h;
  h.rho = 0.0;
  if (estimator == IMB_EST_DIAGONAL) {
    IMB_TRY(launch_colvar(sims, N, m, xvar, s));
    copy_cov_kernel<<<grid_blocks((long long)m * m, 256, 8), 256, 0, s>>>(nullptr, IMB_COV_NONE, m, xvar,
                                                                       out_cov, m);
    IMB_LAUNCHED();
    IMB_CUDA(cudaStreamSynchronize(s));
  } else {
    IMB_TRY(launch_gram(sims, N, m, out_mean, nullptr, partial, G, s));
    IMB_TRY(launch_oas_scalars(G, N, m, estimator, sc, s));
    const int nt = (m + TB - 1) / TB;
    build_cov_kernel<<<dim3(nt, nt), 256, 0, s>>>(sims, N, m, out_mean, sc, nullptr, IMB_COV_NONE, out_cov, m);
    IMB_LAUNCHED();
    IMB_TRY(read_scalars(sc, &h, s));
  }
  if (out_rho) *out_rho = h.rho;
  return IMB_OK;
}
