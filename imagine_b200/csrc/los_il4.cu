// Stage 2, production kernel: line-of-sight integration over the realisation-interleaved field
// layout B[group][cell][4 realisations][4] (b_stride 16), built for OCCUPANCY.
//
// The first interleaved kernel (los.cu:los_il4_kernel) is latency bound: ncu shows ~45 % of the
// stall samples on the first use of the gathered corners and only 16 resident warps per SM,
// because 8 x 256-bit loads in flight + per-group ray accumulators need 128 registers.  Here
//   * the ray accumulators (DM, Faraday carry, I, Q, U per realisation group) live in shared
//     memory, [slot][thread] so every access is conflict free;
//   * per-pixel constants (direction, sky basis) are staged once per warp in shared memory and
//     re-read as broadcasts;
//   * the eight corners are consumed as four z-pairs (lerp in z as soon as a pair lands);
//   * p = 3 (the reference CRE default, fields/hamx/cre_analytic.py:38) gets a division-free,
//     pow-free emissivity.
// That brings the kernel to <= 80 registers (3 blocks = 24 warps per SM) or 64 (32 warps).
// Work decomposition is unchanged: warp = one pixel x 8 consecutive samples x 4 realisations
// (lane = 4*sample + realisation), looping over NG groups so geometry and the shared
// n_e / CRE interpolants are reused.  Numeric definition: oracle/los.py.
#include <cstdlib>
#include "los.cuh"

namespace imb {

struct PixConst {          // per warp, in shared memory
  double n[3], north[3], east[2];
  int j_lo, j_hi;
};

template <bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3, int NG, int MINB>
__global__ void __launch_bounds__(256, MINB) los_il4s_kernel(const __grid_constant__ LosArgs a) {
  extern __shared__ double s_acc[];                     // [NG*nslot][256]
  __shared__ PixConst s_pix[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = lane >> 2, r4 = lane & 3;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= a.m) return;                               // whole warp exits together
  const int g0 = blockIdx.y * NG;
  const int n_groups = (a.n_real + 3) >> 2;
  PixConst& pc = s_pix[warp];

  if (lane == 0) {
    const double nx = a.dirs[3 * pix], ny = a.dirs[3 * pix + 1], nz = a.dirs[3 * pix + 2];
    const double sth = sqrt(nx * nx + ny * ny);
    const double cph = sth > 0.0 ? nx / sth : 1.0, sph = sth > 0.0 ? ny / sth : 0.0;
    pc.n[0] = nx; pc.n[1] = ny; pc.n[2] = nz;
    pc.east[0] = -sph; pc.east[1] = cph;
    pc.north[0] = -nz * cph; pc.north[1] = -nz * sph; pc.north[2] = sth;
    // conservative in-box step range (exact in-box test is still done per sample)
    double tmin = 0.0, tmax = a.dr * (double)a.n_steps;
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
    int j_lo = 0, j_hi = -1;
    if (tmax >= tmin) {
      j_lo = max(0, (int)floor(tmin / a.dr - 0.5) - 1);
      j_hi = min(a.n_steps - 1, (int)ceil(tmax / a.dr - 0.5) + 1);
    }
    pc.j_lo = j_lo; pc.j_hi = j_hi;
  }
#pragma unroll
  // slots per group: carry, I, Q, U (+ DM only when n_e differs per realisation)
  const int nslot = (HAS_NE && a.ne_row != 0) ? 5 : 4;
  for (int q = 0; q < NG * nslot; ++q) s_acc[q * 256 + tid] = 0.0;
  __syncwarp();

  const int j_hi = pc.j_hi;
  const long long sy = a.g.n[2], sx = (long long)a.g.n[1] * a.g.n[2];
  const long long group_stride = a.g.ncell * 16;

  double dm_total = 0.0;                 // shared-n_e dispersion measure of this lane's samples
  for (int jb = pc.j_lo; jb <= j_hi; jb += 8) {
    const int j = jb + slot;
    // -- geometry: bit-identical to oracle/los.py:sample_geometry (no fma on this path)
    const double r = __dmul_rn((double)j + 0.5, a.dr);
    const double px = __dadd_rn(a.obs[0], __dmul_rn(r, pc.n[0]));
    const double py = __dadd_rn(a.obs[1], __dmul_rn(r, pc.n[1]));
    const double pz = __dadd_rn(a.obs[2], __dmul_rn(r, pc.n[2]));
    double ux = __dmul_rn(__dsub_rn(px, a.g.lo[0]), a.g.inv_step[0]);
    double uy = __dmul_rn(__dsub_rn(py, a.g.lo[1]), a.g.inv_step[1]);
    double uz = __dmul_rn(__dsub_rn(pz, a.g.lo[2]), a.g.inv_step[2]);
    const bool inside = (j <= j_hi) && ux >= 0.0 && ux <= (double)(a.g.n[0] - 1) && uy >= 0.0 &&
                        uy <= (double)(a.g.n[1] - 1) && uz >= 0.0 && uz <= (double)(a.g.n[2] - 1);
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
    if (HAS_NE && HAS_CRE && a.necre) {
      if (inside) interp_pair(a.necre, c, ne_s, cre_s);
    } else {
      if (ne_shared) ne_s = inside ? interp_scalar(a.ne, c) : 0.0;
      if (cre_shared) cre_s = inside ? interp_scalar(a.cre, c) : 0.0;
    }
    // the dispersion measure does not depend on the realisation when n_e is shared
    double dm_chunk = 0.0;
    if (ne_shared) dm_chunk = ne_s * a.dl_pc;

    dm_total += dm_chunk;
#pragma unroll 1
    for (int t = 0; t < NG; ++t) {
      const int g = g0 + t;
      if (g >= n_groups) break;
      const int rr = g * 4 + r4;
      const bool live = inside && rr < a.n_real;
      double bx = 0.0, by = 0.0, bz = 0.0;
      if (live) {
        const double* p = a.B + group_stride * g + c.base * 16 + r4 * 4;
        // four z-pairs, each reduced to three doubles as soon as it lands
        double4 lo4 = ldg256(p), hi4 = ldg256(p + 16);
        double x00 = lerp(lo4.x, hi4.x, c.tz), y00 = lerp(lo4.y, hi4.y, c.tz), z00 = lerp(lo4.z, hi4.z, c.tz);
        lo4 = ldg256(p + c.sy * 16); hi4 = ldg256(p + c.sy * 16 + 16);
        double x01 = lerp(lo4.x, hi4.x, c.tz), y01 = lerp(lo4.y, hi4.y, c.tz), z01 = lerp(lo4.z, hi4.z, c.tz);
        x00 = lerp(x00, x01, c.ty); y00 = lerp(y00, y01, c.ty); z00 = lerp(z00, z01, c.ty);
        lo4 = ldg256(p + c.sx * 16); hi4 = ldg256(p + c.sx * 16 + 16);
        double x10 = lerp(lo4.x, hi4.x, c.tz), y10 = lerp(lo4.y, hi4.y, c.tz), z10 = lerp(lo4.z, hi4.z, c.tz);
        lo4 = ldg256(p + (c.sx + c.sy) * 16); hi4 = ldg256(p + (c.sx + c.sy) * 16 + 16);
        x01 = lerp(lo4.x, hi4.x, c.tz); y01 = lerp(lo4.y, hi4.y, c.tz); z01 = lerp(lo4.z, hi4.z, c.tz);
        x10 = lerp(x10, x01, c.ty); y10 = lerp(y10, y01, c.ty); z10 = lerp(z10, z01, c.ty);
        bx = lerp(x00, x10, c.tx); by = lerp(y00, y10, c.tx); bz = lerp(z00, z10, c.tx);
      }
      double* acc = s_acc + (size_t)(t * nslot) * 256 + tid;  // [0] carry, [1] I, [2] Q, [3] U, [4] DM
      double phi = 0.0;
      if (HAS_NE) {
        if (!ne_shared) ne_s = live ? interp_scalar(a.ne + a.ne_row * rr, c) : 0.0;
        const double bpar = -(bx * pc.n[0] + by * pc.n[1] + bz * pc.n[2]);
        const double f = __dmul_rn(__dmul_rn(__dmul_rn(0.812, ne_s), bpar), a.dl_pc);
        double x = f;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {          // inclusive scan over the 8 samples of this realisation
          const double up = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += up;
        }
        const double carry = acc[0];
        phi = (carry + x) - 0.5 * f;
        acc[0] = carry + __shfl_sync(0xffffffffu, x, 28 + r4);
        if (!ne_shared) acc[1024] += ne_s * a.dl_pc;
      }
      if (HAS_CRE) {
        if (!cre_shared) cre_s = live ? interp_scalar(a.cre + a.cre_row * rr, c) : 0.0;
        const double bn = bx * pc.north[0] + by * pc.north[1] + bz * pc.north[2];
        const double be = bx * pc.east[0] + by * pc.east[1];
        const double bn2 = bn * bn, be2 = be * be, bp2 = bn2 + be2;
        if (bp2 > 0.0) {
          double em, qc, uc;
          if (P3) {                                  // (p+1)/4 == 1: no pow, no division
            const double ck = cre_s * kp * a.dl_cm;
            em = ck * bp2;
            const double gq = -ck * pif;
            qc = gq * (bn2 - be2);
            uc = gq * (2.0 * bn * be);
          } else {
            em = cre_s * kp * pow(bp2, expo) * a.dl_cm;
            const double gq = -em * pif / bp2;
            qc = gq * (bn2 - be2);
            uc = gq * (2.0 * bn * be);
          }
          acc[256] += em;
          if (HAS_NE) {
            double sa, ca;
            sincos(2.0 * phi * a.lam2, &sa, &ca);
            acc[512] += qc * ca - uc * sa;
            acc[768] += qc * sa + uc * ca;
          } else {
            acc[512] += qc;
            acc[768] += uc;
          }
        }
      }
    }
  }

  const bool dm_shared = HAS_NE && a.ne_row == 0;
  for (int t = 0; t < NG; ++t) {
    const int g = g0 + t;
    if (g >= n_groups) break;
    const int rr = g * 4 + r4;
    const double* acc = s_acc + (size_t)(t * nslot) * 256 + tid;
    double dm = dm_shared ? dm_total : acc[1024], I = acc[256], Q = acc[512], U = acc[768];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      if (HAS_NE) dm += __shfl_xor_sync(0xffffffffu, dm, o);
      if (HAS_CRE) {
        I += __shfl_xor_sync(0xffffffffu, I, o);
        Q += __shfl_xor_sync(0xffffffffu, Q, o);
        U += __shfl_xor_sync(0xffffffffu, U, o);
      }
    }
    if (slot == 0 && rr < a.n_real) {
      const long long o = out_index(a, rr, pix);
      if (HAS_NE) {
        if (a.out_dm) a.out_dm[o] = dm;
        if (a.out_fd) a.out_fd[o] = acc[0];
      }
      if (HAS_CRE) {
        if (a.out_I) a.out_I[o] = I;
        if (a.out_Q) a.out_Q[o] = Q;
        if (a.out_U) a.out_U[o] = U;
      }
    }
  }
}

template <bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3, int NG, int MINB>
static int launch_one(const LosArgs& a, dim3 grid, cudaStream_t s) {
  auto kern = los_il4s_kernel<HAS_NE, HAS_CRE, VAR_P, P3, NG, MINB>;
  const size_t smem = sizeof(double) * NG * ((HAS_NE && a.ne_row != 0) ? 5 : 4) * 256;
  const size_t smem_max = sizeof(double) * NG * 5 * 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(los_il4s_kernel)");
    configured = true;
  }
  kern<<<grid, 256, smem, s>>>(a);
  return IMB_OK;
}

template <bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3>
static int launch_variant(const LosArgs& a, cudaStream_t s) {
  const int n_groups = (a.n_real + 3) / 4;
  int ng = n_groups >= 4 ? 4 : (n_groups >= 2 ? 2 : 1);
  int minb = 4;
  if (const char* e = getenv("IMB_LOS_NG")) ng = atoi(e);
  if (const char* e = getenv("IMB_LOS_MINB")) minb = atoi(e);
  dim3 grid((a.m + 7) / 8, (n_groups + ng - 1) / ng);
  if (ng >= 4) return minb >= 4 ? launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 4, 4>(a, grid, s)
                                : launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 4, 3>(a, grid, s);
  if (ng >= 2) return minb >= 4 ? launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 2, 4>(a, grid, s)
                                : launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 2, 3>(a, grid, s);
  return minb >= 4 ? launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 1, 4>(a, grid, s)
                   : launch_one<HAS_NE, HAS_CRE, VAR_P, P3, 1, 3>(a, grid, s);
}

// entry used by imb_los_integrate for b_stride == 16
int launch_los_il4s(const LosArgs& a, bool has_ne, bool has_cre, bool var_p, cudaStream_t s) {
  const bool p3 = has_cre && !var_p && a.expo == 1.0;
  if (has_ne && has_cre) {
    if (var_p) return launch_variant<true, true, true, false>(a, s);
    return p3 ? launch_variant<true, true, false, true>(a, s) : launch_variant<true, true, false, false>(a, s);
  }
  if (has_ne) return launch_variant<true, false, false, false>(a, s);
  if (var_p) return launch_variant<false, true, true, false>(a, s);
  return p3 ? launch_variant<false, true, false, true>(a, s) : launch_variant<false, true, false, false>(a, s);
}

}  // namespace imb
