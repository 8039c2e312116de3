// Stage 2: line-of-sight integration (Faraday depth, dispersion measure, synchrotron I/Q/U).
//
// One warp owns one sky direction (HEALPix pixel) and a tile of RT realisations; the 32
// lanes take 32 consecutive radial samples, so the eight trilinear corner gathers of a
// load instruction fall on a short run of cache lines along the ray.  Geometry (sample
// position, cell, weights) and the deterministic n_e / CRE interpolants are computed once
// per sample and reused for every realisation of the tile.  The running Faraday depth is a
// warp inclusive scan per 32-sample chunk plus a carried prefix; the ray sums finish with a
// warp-shuffle tree.  Numeric definition: oracle/los.py (the reference delegates this
// stage to hammurabiX, simulators/hammurabi.py:193-227).
#include <cstdlib>
#include "common.cuh"
#include "los.cuh"

namespace imb {

int make_dev_grid(const imb_grid* grid, DevGrid* g);
int launch_los_rl(const LosArgs& a, int rl, bool has_ne, bool has_cre, bool var_p, cudaStream_t s);

template <bool HAS_NE, bool HAS_CRE, bool VAR_P, int RT, int BS>
__global__ void __launch_bounds__(256) los_kernel(const __grid_constant__ LosArgs a) {
  const int lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= a.m) return;
  const int r0 = blockIdx.y * RT;

  const double nx = __ldg(a.dirs + 3 * pix), ny = __ldg(a.dirs + 3 * pix + 1),
               nz = __ldg(a.dirs + 3 * pix + 2);
  // sky basis (IAU): East = (-sin l, cos l, 0), North = (-sin b cos l, -sin b sin l, cos b)
  const double sth = sqrt(nx * nx + ny * ny);
  const double cph = sth > 0.0 ? nx / sth : 1.0, sph = sth > 0.0 ? ny / sth : 0.0;
  const double ex_ = -sph, ey_ = cph;
  const double nox = -nz * cph, noy = -nz * sph, noz = sth;

  // conservative in-box step range from a slab test (exact test still done per sample)
  double tmin = 0.0, tmax = a.dr * (double)a.n_steps;
  {
    const double nn[3] = {nx, ny, nz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double lo = a.g.lo[k], hi = lo + (double)(a.g.n[k] - 1) / a.g.inv_step[k];
      if (nn[k] != 0.0) {
        const double t1 = (lo - a.obs[k]) / nn[k], t2 = (hi - a.obs[k]) / nn[k];
        tmin = fmax(tmin, fmin(t1, t2));
        tmax = fmin(tmax, fmax(t1, t2));
      } else if (a.obs[k] < lo || a.obs[k] > hi) {
        tmax = -1.0;
      }
    }
  }
  int j_lo = 0, j_hi = -1;
  if (tmax >= tmin) {
    j_lo = max(0, (int)floor(tmin / a.dr - 0.5) - 1);
    j_hi = min(a.n_steps - 1, (int)ceil(tmax / a.dr - 0.5) + 1);
  }

  double accI[RT], accQ[RT], accU[RT], carry[RT], accDM[RT];
#pragma unroll
  for (int t = 0; t < RT; ++t) accI[t] = accQ[t] = accU[t] = carry[t] = accDM[t] = 0.0;

  const long long sy = a.g.n[2], sx = (long long)a.g.n[1] * a.g.n[2];
  const double umx = (double)(a.g.n[0] - 1), umy = (double)(a.g.n[1] - 1),
               umz = (double)(a.g.n[2] - 1);

  for (int jb = j_lo; jb <= j_hi; jb += 32) {
    const int j = jb + lane;
    // -- geometry: bit-identical to oracle/los.py:sample_geometry (no fma on this path)
    const double r = __dmul_rn((double)j + 0.5, a.dr);
    const double px = __dadd_rn(a.obs[0], __dmul_rn(r, nx));
    const double py = __dadd_rn(a.obs[1], __dmul_rn(r, ny));
    const double pz = __dadd_rn(a.obs[2], __dmul_rn(r, nz));
    double ux = __dmul_rn(__dsub_rn(px, a.g.lo[0]), a.g.inv_step[0]);
    double uy = __dmul_rn(__dsub_rn(py, a.g.lo[1]), a.g.inv_step[1]);
    double uz = __dmul_rn(__dsub_rn(pz, a.g.lo[2]), a.g.inv_step[2]);
    const bool inside = (j <= j_hi) && ux >= 0.0 && ux <= umx && uy >= 0.0 && uy <= umy &&
                        uz >= 0.0 && uz <= umz;
    if (!inside) ux = uy = uz = 0.0;
    const int ix = min((int)ux, a.g.n[0] - 2), iy = min((int)uy, a.g.n[1] - 2),
              iz = min((int)uz, a.g.n[2] - 2);
    Cell c;
    c.base = ((long long)ix * a.g.n[1] + iy) * a.g.n[2] + iz;
    c.sy = sy; c.sx = sx;
    c.tx = ux - (double)ix; c.ty = uy - (double)iy; c.tz = uz - (double)iz;

    double kp = a.sync_k, pif = a.pol_frac, expo = a.expo;
    if (HAS_CRE && VAR_P) {
      const double R = sqrt(px * px + py * py);
      const double p = a.alpha - a.beta * R - a.theta * pz;
      kp = sync_coeff_var(a, p);
      pif = (p + 1.0) / (p + 7.0 / 3.0);
      expo = (p + 1.0) * 0.25;
    }

    double ne_s = 0.0, cre_s = 0.0;
    const bool ne_shared = HAS_NE && a.ne_row == 0, cre_shared = HAS_CRE && a.cre_row == 0;
    if (ne_shared) ne_s = inside ? interp_scalar(a.ne, c) : 0.0;
    if (cre_shared) cre_s = inside ? interp_scalar(a.cre, c) : 0.0;

#pragma unroll
    for (int t = 0; t < RT; ++t) {
      const int rr = r0 + t;
      if (rr >= a.n_real) break;
      double bx = 0.0, by = 0.0, bz = 0.0;
      if (inside) interp_vector<BS>(a.B + a.B_row * rr, c, bx, by, bz);
      double phi = 0.0;
      if (HAS_NE) {
        if (!ne_shared) ne_s = inside ? interp_scalar(a.ne + a.ne_row * rr, c) : 0.0;
        const double bpar = -(bx * nx + by * ny + bz * nz);
        const double f = __dmul_rn(__dmul_rn(__dmul_rn(0.812, ne_s), bpar), a.dl_pc);
        double x = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += up;
        }
        phi = (carry[t] + x) - 0.5 * f;
        carry[t] += __shfl_sync(0xffffffffu, x, 31);
        accDM[t] += ne_s * a.dl_pc;
      }
      if (HAS_CRE) {
        if (!cre_shared) cre_s = inside ? interp_scalar(a.cre + a.cre_row * rr, c) : 0.0;
        const double bn = bx * nox + by * noy + bz * noz;
        const double be = bx * ex_ + by * ey_;
        const double bp2 = bn * bn + be * be;
        if (bp2 > 0.0) {
          const double pw = (expo == 1.0) ? bp2 : pow(bp2, expo);
          const double em = cre_s * kp * pw * a.dl_cm;
          const double g = -em * pif / bp2;
          const double qc = g * (bn * bn - be * be), uc = g * (2.0 * bn * be);
          accI[t] += em;
          if (HAS_NE) {
            double sa, ca;
            sincos(2.0 * phi * a.lam2, &sa, &ca);
            accQ[t] += qc * ca - uc * sa;
            accU[t] += qc * sa + uc * ca;
          } else {
            accQ[t] += qc;
            accU[t] += uc;
          }
        }
      }
    }
  }

#pragma unroll
  for (int t = 0; t < RT; ++t) {
    const int rr = r0 + t;
    if (rr >= a.n_real) break;
    const long long o = out_index(a, rr, pix);
    if (HAS_NE) {
      const double dm = warp_sum(accDM[t]);
      if (lane == 0) {
        if (a.out_dm) a.out_dm[o] = dm;
        if (a.out_fd) a.out_fd[o] = carry[t];
      }
    }
    if (HAS_CRE) {
      const double I = warp_sum(accI[t]), Q = warp_sum(accQ[t]), U = warp_sum(accU[t]);
      if (lane == 0) {
        if (a.out_I) a.out_I[o] = I;
        if (a.out_Q) a.out_Q[o] = Q;
        if (a.out_U) a.out_U[o] = U;
      }
    }
  }
}

template <bool HAS_NE, bool HAS_CRE, bool VAR_P, int RT>
static void launch_bs(const LosArgs& a, int bs, dim3 grid, cudaStream_t s) {
  if (bs == 4) los_kernel<HAS_NE, HAS_CRE, VAR_P, RT, 4><<<grid, 256, 0, s>>>(a);
  else los_kernel<HAS_NE, HAS_CRE, VAR_P, RT, 3><<<grid, 256, 0, s>>>(a);
}

template <bool HAS_NE, bool HAS_CRE, bool VAR_P>
static void launch_rt(const LosArgs& a, int bs, int rt, cudaStream_t s) {
  const int warps = 8;
  dim3 grid((a.m + warps - 1) / warps, (a.n_real + rt - 1) / rt);
  if (rt == 4) launch_bs<HAS_NE, HAS_CRE, VAR_P, 4>(a, bs, grid, s);
  else if (rt == 2) launch_bs<HAS_NE, HAS_CRE, VAR_P, 2>(a, bs, grid, s);
  else launch_bs<HAS_NE, HAS_CRE, VAR_P, 1>(a, bs, grid, s);
}

__global__ void pack_pairs_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                  long long n, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<double2*>(out)[i] = make_double2(x[i], y[i]);
}

int grid_blocks(long long work, int threads, int per_sm);

}  // namespace imb

using namespace imb;

// CGS constants, identical literals to oracle/los.py
static const double kCLight = 2.99792458e10, kQe = 4.80320471e-10, kMe = 9.10938356e-28,
                    kKb = 1.38064852e-16, kGevErg = 1.6021766208e-3,
                    kKpcCm = 3.0856775814913673e21;

extern "C" int imb_los_integrate(const imb_grid* grid, const imb_los_cfg* cfg, const double* dirs,
                                 int m, const double* B, int rows_B, int b_stride,
                                 const double* ne, int rows_ne, const double* cre, int rows_cre,
                                 int n_real, double* out_dm, double* out_fd, double* out_I,
                                 double* out_Q, double* out_U, void* stream) {
  IMB_TRY(check_init());
  LosArgs a;
  IMB_TRY(make_dev_grid(grid, &a.g));
  if (m == 0 && cfg && n_real >= 1) return IMB_OK;            // nothing to integrate
  if (!cfg || !dirs || !B || m < 0 || n_real < 1) return IMB_EINVAL;
  if (b_stride != 3 && b_stride != 4 && b_stride != 24 && b_stride != 48 && b_stride != 96)
    return IMB_EINVAL;
  if (b_stride >= 4 && (reinterpret_cast<uintptr_t>(B) % 32 != 0)) return IMB_EINVAL;
  if (b_stride >= 24 && rows_B != n_real) return IMB_EINVAL;   // interleaved layouts are per realisation
  for (int k = 0; k < 3; ++k)
    if (grid->n[k] < 2 || !(grid->hi[k] > grid->lo[k])) return IMB_EINVAL;  // trilinear needs a cell
  if (cfg->n_steps < 0 || !(cfg->dr > 0.0)) return IMB_EINVAL;
  if ((rows_B != 1 && rows_B != n_real) || (ne && rows_ne != 1 && rows_ne != n_real) ||
      (cre && rows_cre != 1 && rows_cre != n_real))
    return IMB_EINVAL;
  const bool want_ne = out_dm || out_fd, want_cre = out_I || out_Q || out_U;
  if ((want_ne && !ne) || (want_cre && !cre)) return IMB_EINVAL;
  if (!want_ne && !want_cre) return IMB_EINVAL;
  if (m == 0) return IMB_OK;
  const bool has_ne = ne != nullptr, has_cre = want_cre;

  for (int k = 0; k < 3; ++k) a.obs[k] = cfg->observer[k];
  a.dr = cfg->dr;
  a.dl_pc = cfg->dr * 1.0e3;
  a.dl_cm = cfg->dr * kKpcCm;
  a.n_steps = cfg->n_steps;
  a.m = m;
  a.n_real = n_real;
  const double nu = cfg->nu_ghz * 1.0e9;
  const double lam = (kCLight * 1.0e-2) / nu;
  a.lam2 = lam * lam;
  a.alpha = cfg->alpha; a.beta = cfg->beta; a.theta = cfg->theta;
  const bool var_p = has_cre && (cfg->beta != 0.0 || cfg->theta != 0.0);
  a.sync_k = cfg->sync_k;
  a.pol_frac = (cfg->alpha + 1.0) / (cfg->alpha + 7.0 / 3.0);
  a.expo = (cfg->alpha + 1.0) / 4.0;
  {
    // K(p) = kA/(p+1) * Gamma * Gamma * exp(kE0 + p*kE1); see oracle/los.py:sync_coefficient
    const double mec2 = kMe * kCLight * kCLight;
    const double t_conv = kCLight * kCLight / (2.0 * kKb * nu * nu);
    const double W = 2.0 * M_PI * kMe * kCLight * nu / (3.0 * kQe);
    a.kA = t_conv * 0.5 * sqrt(3.0) * kQe * kQe * kQe / (2.0 * M_PI * mec2) *
           (4.0 * M_PI / kCLight) / (kGevErg * 1.0e4);
    const double lw = log(W), le0 = log(cfg->e0_gev * kGevErg), lm = log(mec2), l6 = log(1.0e-6);
    a.kE0 = 0.5 * lw + lm + 0.5 * l6;
    a.kE1 = -0.5 * lw + le0 - lm + 0.5 * l6;
  }
  a.dirs = dirs;
  a.B = B;     a.B_row = rows_B > 1 ? a.g.ncell * b_stride : 0;
  a.ne = ne;   a.ne_row = (ne && rows_ne > 1) ? a.g.ncell : 0;
  a.cre = cre; a.cre_row = (cre && rows_cre > 1) ? a.g.ncell : 0;
  a.out_dm = out_dm; a.out_fd = out_fd; a.out_I = out_I; a.out_Q = out_Q; a.out_U = out_U;
  a.necre = nullptr;
  a.out_mp = m; a.out_dest = 0;
  if (cfg->out_cols_per_dest > 0) {
    if (m % cfg->out_cols_per_dest != 0 || cfg->out_dest_stride < 0) return IMB_EINVAL;
    a.out_mp = cfg->out_cols_per_dest;
    a.out_dest = cfg->out_dest_stride;
  }
  a.n_peers = 0;
  if (cfg->n_peers > 0) {
    if (cfg->n_peers > 16 || cfg->out_cols_per_dest <= 0 || m / cfg->out_cols_per_dest != cfg->n_peers)
      return IMB_EINVAL;
    a.n_peers = cfg->n_peers;
    for (int d = 0; d < cfg->n_peers; ++d) a.peer_delta[d] = cfg->peer_delta[d];
  }
  if (b_stride >= 24 && has_ne && has_cre && rows_ne == 1 && rows_cre == 1) {
    // pack the two shared scalar grids into (n_e, CRE) pairs: one 128-bit gather per vertex
    IMB_TRY(ensure_buffer(&state().aux, &state().aux_bytes, sizeof(double) * 2 * (size_t)a.g.ncell));
    double* packed = static_cast<double*>(state().aux);
    pack_pairs_kernel<<<grid_blocks(a.g.ncell, 256, 8), 256, 0, (cudaStream_t)stream>>>(ne, cre, a.g.ncell, packed);
    IMB_LAUNCHED();
    a.necre = packed;
  }

  cudaStream_t s = (cudaStream_t)stream;
  if (b_stride >= 24) {                       // lane = realisation (los_rl.cu); always one row per realisation
    IMB_TRY(launch_los_rl(a, b_stride / 3, has_ne, has_cre, var_p, s));
    IMB_LAUNCHED();
    return IMB_OK;
  }
  // row-major interchange layouts: realisation tile, reuse geometry + deterministic interpolants
  int rt = 1;
  const bool per_real = rows_B > 1 || rows_ne > 1 || rows_cre > 1;
  if (!per_real && cfg->out_cols_per_dest > 0) return IMB_EUNSUPPORTED;
  if (!per_real) {
    // every realisation is identical: compute one row, replicate below
    a.n_real = 1;
  } else {
    rt = n_real >= 4 ? 4 : (n_real >= 2 ? 2 : 1);
  }
  if (has_ne && has_cre) {
    if (var_p) launch_rt<true, true, true>(a, b_stride, rt, s);
    else launch_rt<true, true, false>(a, b_stride, rt, s);
  } else if (has_ne) {
    launch_rt<true, false, false>(a, b_stride, rt, s);
  } else {
    if (var_p) launch_rt<false, true, true>(a, b_stride, rt, s);
    else launch_rt<false, true, false>(a, b_stride, rt, s);
  }
  IMB_LAUNCHED();
  if (!per_real && n_real > 1) {
    double* outs[5] = {out_dm, out_fd, out_I, out_Q, out_U};
    for (double* o : outs) {
      if (!o) continue;
      for (int r = 1; r < n_real; ++r)
        IMB_CUDA(cudaMemcpyAsync(o + (size_t)r * m, o, sizeof(double) * (size_t)m,
                                 cudaMemcpyDeviceToDevice, s));
    }
  }
  return IMB_OK;
}
