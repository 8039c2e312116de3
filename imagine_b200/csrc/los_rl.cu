// Stage 2, realisation-lane kernel: line-of-sight integration over the wide realisation-interleaved
// layout B[group][cell][component][RL] (b_stride 3*RL, RL = 32, 16 or 8 realisations per group).
//
// los_il4s_kernel (lane = 4*sample + realisation) is bound by the L1 data pipe (ncu: 84-86 % of
// the wavefront peak): every 256-bit corner gather costs the 8-wavefront minimum, a quarter of the
// bytes are padding, the (n_e, CRE) gathers and the shared-memory ray accumulators ride the same
// pipe.  Here a LANE IS A REALISATION and the warp marches ONE ray sample by sample:
//   * every corner component is one 64-bit load per lane, RL consecutive doubles per request
//     (256 B = 2 wavefronts at RL 32, no padding): 48 wavefronts per 32 sample-realisations
//     instead of ~140;
//   * geometry, the in-box test, the shared n_e / CRE interpolants and the spectral-index terms do
//     not depend on the realisation: the 32 lanes compute them for 32 consecutive steps at once
//     (phase A), park them in shared memory, and the march (phase B) re-reads them as broadcasts;
//     only in-box steps are marched (ballot mask);
//   * the ray accumulators (Faraday depth, I, Q, U) are plain registers: no scan, no shuffles, no
//     shared-memory read-modify-write in the loop;
//   * RL = 16 / 8: the 2 / 4 lane groups march consecutive segments of the ray; when the segments are
//     joined, each outer one is rotated by the total Faraday depth of those inside it (the rotation
//     is linear in Q+iU).
// Measured and dropped: an L1 prefetch (CCTL.PF1, one line per lane) of the next in-box step's corner
// lines -- 4.24 vs 3.47 ms; the prefetch requests cost more LSU time than the latency they hide.
// 128-thread blocks at 72 registers (28 instead of 24 warps / SM): 3.58 vs 3.47 ms.  Two realisation
// groups per warp (geometry and broadcasts shared, half the warps): 3.88-4.24 vs 3.48 ms.  Not built:
// reusing the 24 corner values when the next step stays in the same cell (32 % of the steps at C3)
// needs ~120 registers (280 bytes of spills at 80), i.e. 16 instead of 24 warps / SM.
// Numeric definition: oracle/los.py.
#include <cstdlib>
#include "los.cuh"

namespace imb {

struct __align__(16) RlPix {   // per warp, in shared memory
  // sky basis by its angles: n = (st cp, st sp, ct), north = (-ct cp, -ct sp, st), east = (-sp, cp, 0);
  // the march re-reads (st, ct) and (cp, sp) as two 128-bit words per step instead of n|north|east
  double st, ct, cp, sp;
  double n[3];
  int j_lo, j_hi;
};

struct __align__(16) RlGeo {   // per (warp, step slot): everything realisation-independent
  double tx, ty, tz, ne, cre;
  int base, pad;               // base < 0: step is outside the box
};

// SHARED: every scalar grid present (n_e, CRE) is one grid for all realisations -- the usual case;
// the per-realisation gathers of RandomThermalElectrons are then compiled out of the march
template <int RL, bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3, bool SHARED, int MINB>
__global__ void __launch_bounds__(256, MINB) los_rl_kernel(const __grid_constant__ LosArgs a) {
  constexpr int S = 32 / RL;                 // ray segments marched side by side by one warp
  __shared__ RlPix s_pix[8];
  __shared__ RlGeo s_geo[8][32];
  __shared__ double s_var[VAR_P ? 8 : 1][32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = lane / RL, rl = lane % RL;
  const int pix = blockIdx.x * 8 + warp;
  if (pix >= a.m) return;                               // whole warp exits together
  const int g = blockIdx.y;
  const int rr = g * RL + rl;
  const bool live = rr < a.n_real;
  RlPix& pc = s_pix[warp];

  if (lane == 0) {
    const double nx = a.dirs[3 * pix], ny = a.dirs[3 * pix + 1], nz = a.dirs[3 * pix + 2];
    const double sth = sqrt(nx * nx + ny * ny);
    const double cph = sth > 0.0 ? nx / sth : 1.0, sph = sth > 0.0 ? ny / sth : 0.0;
    pc.n[0] = nx; pc.n[1] = ny; pc.n[2] = nz;
    pc.st = sth; pc.ct = nz; pc.cp = cph; pc.sp = sph;
    // conservative in-box step range (the exact in-box test is still done per step)
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
  __syncwarp();

  const int j_lo = pc.j_lo, j_hi = pc.j_hi;
  const int seg_len = (j_hi - j_lo + 1 + S - 1) / S;     // <= 0 when the ray misses the box
  const long long sy = a.g.n[2], sx = (long long)a.g.n[1] * a.g.n[2];
  const bool ne_shared = HAS_NE && (SHARED || a.ne_row == 0), cre_shared = HAS_CRE && (SHARED || a.cre_row == 0);
  // corner strides of the wide layout, in doubles
  constexpr long long CS = 3 * RL;
  const long long oy = sy * CS, ox = sx * CS;
  const double* __restrict__ Bg = a.B + (long long)g * a.g.ncell * CS + rl;

  const double two_lam2 = 2.0 * a.lam2;
  double dm_geo = 0.0;                         // shared-n_e dispersion measure of this lane's steps
  double carry = 0.0, dm = 0.0, I = 0.0, Q = 0.0, U = 0.0;

  for (int cb = 0; cb < seg_len; cb += RL) {
    // ---- phase A: lane (seg, rl) prepares step j of its segment --------------------------------
    {
      const int off = cb + rl;
      const int j = j_lo + seg * seg_len + off;
      // geometry: bit-identical to oracle/los.py:sample_geometry (no fma on this path)
      const double r = __dmul_rn((double)j + 0.5, a.dr);
      const double px = __dadd_rn(a.obs[0], __dmul_rn(r, pc.n[0]));
      const double py = __dadd_rn(a.obs[1], __dmul_rn(r, pc.n[1]));
      const double pz = __dadd_rn(a.obs[2], __dmul_rn(r, pc.n[2]));
      double ux = __dmul_rn(__dsub_rn(px, a.g.lo[0]), a.g.inv_step[0]);
      double uy = __dmul_rn(__dsub_rn(py, a.g.lo[1]), a.g.inv_step[1]);
      double uz = __dmul_rn(__dsub_rn(pz, a.g.lo[2]), a.g.inv_step[2]);
      const bool inside = off < seg_len && j <= j_hi && ux >= 0.0 && ux <= (double)(a.g.n[0] - 1) &&
                          uy >= 0.0 && uy <= (double)(a.g.n[1] - 1) && uz >= 0.0 &&
                          uz <= (double)(a.g.n[2] - 1);
      if (!inside) ux = uy = uz = 0.0;
      const int ix = min((int)ux, a.g.n[0] - 2), iy = min((int)uy, a.g.n[1] - 2),
                iz = min((int)uz, a.g.n[2] - 2);
      Cell c;
      c.base = ((long long)ix * a.g.n[1] + iy) * a.g.n[2] + iz;
      c.sy = sy; c.sx = sx;
      c.tx = ux - (double)ix; c.ty = uy - (double)iy; c.tz = uz - (double)iz;
      double ne_s = 0.0, cre_s = 0.0;
      if (inside) {
        if (HAS_NE && HAS_CRE && a.necre) {
          interp_pair(a.necre, c, ne_s, cre_s);
        } else {
          if (ne_shared) ne_s = interp_scalar(a.ne, c);
          if (cre_shared) cre_s = interp_scalar(a.cre, c);
        }
      }
      if (ne_shared) dm_geo += ne_s * a.dl_pc;
      double kp = a.sync_k;
      if (HAS_CRE && VAR_P) {
        const double R = sqrt(px * px + py * py);
        const double p = a.alpha - a.beta * R - a.theta * pz;
        kp = inside ? sync_coeff_var(a, p) : 0.0;
        s_var[warp][lane][0] = kp;
        s_var[warp][lane][1] = (p + 1.0) / (p + 7.0 / 3.0);
        s_var[warp][lane][2] = (p + 1.0) * 0.25;
      }
      // the realisation-independent leading factors are folded in here, in the oracle's order:
      // 0.812 n_e (Faraday) and CRE K(p) [dl] (emissivity)
      RlGeo& ge = s_geo[warp][lane];
      ge.tx = c.tx; ge.ty = c.ty; ge.tz = c.tz;
      ge.ne = __dmul_rn(0.812, ne_s);
      ge.cre = P3 ? cre_s * kp * a.dl_cm : cre_s * kp;
      ge.base = inside ? (int)c.base : -1;
    }
    __syncwarp();

    // ---- phase B: march the in-box steps of this chunk, one realisation per lane ----------------
    unsigned int todo = 0;
    {
      const RlGeo& mine = s_geo[warp][lane];
      const unsigned int mask = __ballot_sync(0xffffffffu, mine.base >= 0);
      if (S == 1) {
        todo = mask;
      } else {
#pragma unroll
        for (int q = 0; q < S; ++q) todo |= (mask >> (q * RL)) & ((RL < 32 ? (1u << RL) : 0u) - 1u);
      }
    }
    for (; todo; todo &= todo - 1) {
      const int sl = seg * RL + (__ffs(todo) - 1);
      // three 128-bit broadcast reads: (tx, ty) (tz, n_e) (CRE, base)
      const double2* gp = reinterpret_cast<const double2*>(&s_geo[warp][sl]);
      const double2 g2 = gp[2];
      const int base = __double2loint(g2.y);
      if (S > 1 && base < 0) continue;                  // the other segment's step, not mine
      const double2 g0 = gp[0], g1 = gp[1];
      const double tx = g0.x, ty = g0.y, tz = g1.x;
      const double* p = Bg + (long long)base * CS;
      double bx, by, bz;
      {
        // all 24 corner loads are issued before the first lerp: one exposed memory round trip per
        // step instead of one per z-pair (ncu source page: the stall samples sat on the first use
        // after each group of six loads)
        const double* q1 = p + oy;
        const double* q2 = p + ox;
        const double* q3 = q2 + oy;
        double v0[6], v1[6], v2[6], v3[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) v0[k] = ldg64(p + (k / 3) * CS + (k % 3) * RL);
#pragma unroll
        for (int k = 0; k < 6; ++k) v1[k] = ldg64(q1 + (k / 3) * CS + (k % 3) * RL);
#pragma unroll
        for (int k = 0; k < 6; ++k) v2[k] = ldg64(q2 + (k / 3) * CS + (k % 3) * RL);
#pragma unroll
        for (int k = 0; k < 6; ++k) v3[k] = ldg64(q3 + (k / 3) * CS + (k % 3) * RL);
        double x00 = lerp(v0[0], v0[3], tz), y00 = lerp(v0[1], v0[4], tz), z00 = lerp(v0[2], v0[5], tz);
        const double x01 = lerp(v1[0], v1[3], tz), y01 = lerp(v1[1], v1[4], tz), z01 = lerp(v1[2], v1[5], tz);
        x00 = lerp(x00, x01, ty); y00 = lerp(y00, y01, ty); z00 = lerp(z00, z01, ty);
        double x10 = lerp(v2[0], v2[3], tz), y10 = lerp(v2[1], v2[4], tz), z10 = lerp(v2[2], v2[5], tz);
        const double x11 = lerp(v3[0], v3[3], tz), y11 = lerp(v3[1], v3[4], tz), z11 = lerp(v3[2], v3[5], tz);
        x10 = lerp(x10, x11, ty); y10 = lerp(y10, y11, ty); z10 = lerp(z10, z11, ty);
        bx = lerp(x00, x10, tx); by = lerp(y00, y10, tx); bz = lerp(z00, z10, tx);
      }
      const double2 sc = reinterpret_cast<const double2*>(&pc)[0];   // (sin, cos) of the colatitude
      const double2 cs = reinterpret_cast<const double2*>(&pc)[1];   // (cos, sin) of the longitude
      const double bh = bx * cs.x + by * cs.y;                         // horizontal part along the ray's azimuth
      Cell c;
      if ((HAS_NE && !ne_shared) || (HAS_CRE && !cre_shared)) {
        c.base = base; c.sy = sy; c.sx = sx; c.tx = tx; c.ty = ty; c.tz = tz;
      }
      double phi = 0.0;
      if (HAS_NE) {
        double ne812 = g1.y;                               // 0.812 n_e
        if (!ne_shared) {
          const double ne_s = live ? interp_scalar(a.ne + a.ne_row * rr, c) : 0.0;
          dm += ne_s * a.dl_pc;
          ne812 = __dmul_rn(0.812, ne_s);
        }
        const double bpar = -(sc.x * bh + sc.y * bz);      // -(B . n)
        const double f = __dmul_rn(__dmul_rn(ne812, bpar), a.dl_pc);
        phi = carry + 0.5 * f;
        carry += f;
      }
      if (HAS_CRE) {
        double pif = a.pol_frac, expo = a.expo;
        if (VAR_P) { pif = s_var[warp][sl][1]; expo = s_var[warp][sl][2]; }
        double ck = g2.x;                                  // CRE K(p) dl (P3) or CRE K(p)
        if (!cre_shared) {
          const double cre_s = live ? interp_scalar(a.cre + a.cre_row * rr, c) : 0.0;
          const double kp = VAR_P ? s_var[warp][sl][0] : a.sync_k;
          ck = P3 ? cre_s * kp * a.dl_cm : cre_s * kp;
        }
        const double bn = sc.x * bz - sc.y * bh;           // B . north
        const double be = by * cs.x - bx * cs.y;           // B . east
        const double bn2 = bn * bn, be2 = be * be, bp2 = bn2 + be2;
        if (bp2 > 0.0) {
          double em, qc, uc;
          if (P3) {                                  // (p+1)/4 == 1: no pow, no division
            em = ck * bp2;
            const double gq = -ck * pif;
            qc = gq * (bn2 - be2);
            uc = gq * (2.0 * bn * be);
          } else {
            em = ck * pow(bp2, expo) * a.dl_cm;
            const double gq = -em * pif / bp2;
            qc = gq * (bn2 - be2);
            uc = gq * (2.0 * bn * be);
          }
          I += em;
          if (HAS_NE) {
            double sa, ca;
            sincos_cw(phi * two_lam2, sa, ca);              // == 2 phi lam2 exactly
            Q = fma(qc, ca, fma(-uc, sa, Q));
            U = fma(qc, sa, fma(uc, ca, U));
          } else {
            Q += qc;
            U += uc;
          }
        }
      }
    }
    __syncwarp();                                        // phase A of the next chunk reuses s_geo
  }

  // ---- join the segments: the outer one starts at the inner one's total Faraday depth ----------
  if (S > 1) {
#pragma unroll
    for (int o = RL; o < 32; o <<= 1) {
      const double cu = __shfl_down_sync(0xffffffffu, carry, o);
      const double du = __shfl_down_sync(0xffffffffu, dm, o);
      const double iu = __shfl_down_sync(0xffffffffu, I, o);
      const double qu = __shfl_down_sync(0xffffffffu, Q, o);
      const double uu = __shfl_down_sync(0xffffffffu, U, o);
      if (HAS_NE && HAS_CRE) {
        double sa, ca;
        sincos(2.0 * carry * a.lam2, &sa, &ca);
        Q += qu * ca - uu * sa;
        U += qu * sa + uu * ca;
      } else {
        Q += qu; U += uu;
      }
      I += iu; dm += du; carry += cu;
    }
  }
  if (ne_shared) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) dm_geo += __shfl_xor_sync(0xffffffffu, dm_geo, o);
    dm = dm_geo;
  }
  if (seg == 0 && live) {
    const long long o = out_index(a, rr, pix);
    if (HAS_NE) {
      if (a.out_dm) a.out_dm[o] = dm;
      if (a.out_fd) a.out_fd[o] = carry;
    }
    if (HAS_CRE) {
      if (a.out_I) a.out_I[o] = I;
      if (a.out_Q) a.out_Q[o] = Q;
      if (a.out_U) a.out_U[o] = U;
    }
  }
}

template <int RL, bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3>
static int launch_rl_variant(const LosArgs& a, cudaStream_t s) {
  const int n_groups = (a.n_real + RL - 1) / RL;
  dim3 grid((a.m + 7) / 8, n_groups);
  // measured at C3 (64 realisations): RL 32 -> 80 registers (3.40 vs 3.69 ms), RL 16 -> 64 (3.84 vs 4.01)
  int minb = RL == 32 ? 3 : 4;
  if (const char* e = getenv("IMB_LOS_MINB")) minb = atoi(e);
  const bool shared = (!HAS_NE || a.ne_row == 0) && (!HAS_CRE || a.cre_row == 0);
  if (shared) {
    if (minb >= 4) los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, true, 4><<<grid, 256, 0, s>>>(a);
    else los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, true, 3><<<grid, 256, 0, s>>>(a);
  } else {
    if (minb >= 4) los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, false, 4><<<grid, 256, 0, s>>>(a);
    else los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, false, 3><<<grid, 256, 0, s>>>(a);
  }
  return IMB_OK;
}

template <int RL>
static int launch_rl(const LosArgs& a, bool has_ne, bool has_cre, bool var_p, cudaStream_t s) {
  const bool p3 = has_cre && !var_p && a.expo == 1.0;
  if (has_ne && has_cre) {
    if (var_p) return launch_rl_variant<RL, true, true, true, false>(a, s);
    return p3 ? launch_rl_variant<RL, true, true, false, true>(a, s)
              : launch_rl_variant<RL, true, true, false, false>(a, s);
  }
  if (has_ne) return launch_rl_variant<RL, true, false, false, false>(a, s);
  if (var_p) return launch_rl_variant<RL, false, true, true, false>(a, s);
  return p3 ? launch_rl_variant<RL, false, true, false, true>(a, s)
            : launch_rl_variant<RL, false, true, false, false>(a, s);
}

// entry used by imb_los_integrate for b_stride == 96 (RL 32), 48 (RL 16) and 24 (RL 8)
int launch_los_rl(const LosArgs& a, int rl, bool has_ne, bool has_cre, bool var_p, cudaStream_t s) {
  if (a.g.ncell >= (1ll << 31)) return IMB_EUNSUPPORTED;      // 32-bit cell index in RlGeo
  if (rl == 32) return launch_rl<32>(a, has_ne, has_cre, var_p, s);
  if (rl == 16) return launch_rl<16>(a, has_ne, has_cre, var_p, s);
  return launch_rl<8>(a, has_ne, has_cre, var_p, s);
}

}  // namespace imb
