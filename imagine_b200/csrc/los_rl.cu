// Stage 2, realisation-lane kernel: line-of-sight integration over the wide realisation-interleaved
// layout B[group][cell][component][RL] (b_stride 3*RL, RL = 32, 16 or 8 realisations per group).
//
// A BLOCK IS A PIXEL, A WARP IS A REALISATION GROUP, A LANE IS A REALISATION; the warp marches the
// ray sample by sample:
//   * every corner component is one 64-bit load per lane, RL consecutive doubles per request
//     (256 B = 2 L1 wavefronts at RL 32, no padding);
//   * geometry, the in-box test, the shared n_e / CRE interpolants, the trilinear weights and the
//     spectral-index terms do not depend on the realisation: the lanes of ALL the block's warps compute
//     them for 32 G consecutive steps at once (phase A, G = warps per block), park them in shared
//     memory, and every warp's march (phase B) re-reads them as broadcasts.  One pixel per block also
//     gives the block scheduler the finest grain: rays differ 7x in length, and a block of several
//     pixels holds its registers until its longest ray is done (4 pixels per block: 2.31 ms, one:
//     2.05 ms at C3);
//   * THE CELL STAYS RESIDENT IN REGISTERS.  The first version of this kernel re-fetched all 24
//     corner values every step and was bound by the L1 data pipe (ncu r01: 65 GB through L1 for
//     3.4 GB of field, 72 % of the wavefront peak).  At C3 32 % of the in-box steps stay in the cell
//     of the previous step and 47 % cross exactly one face.  The 8 vertices live in 8 register slots
//     addressed by the PARITY of the vertex index per axis (vertex (X,Y,Z) -> slot (X&1, Y&1, Z&1)),
//     so a vertex never moves between slots while the ray walks: crossing a face replaces the four
//     slots of the face left behind and nothing is shuffled.  Phase A hands every step the 8 trilinear
//     weights in slot order (24 FMAs per sample instead of 21 lerps = 42 operations) and an 8-bit mask
//     of the slots to (re)load; the usual masks branch to unpredicated loads of exactly those slots.
//     Average corner loads per step at C3: 24 -> 9.8 (l1tex global-load sectors 2.04e9 -> 0.86e9);
//   * the march is software-pipelined: the loads of step k+1 are issued as soon as step k's
//     interpolation has consumed the slots and fly during step k's physics, which shares a basic block
//     with the interpolation of step k+1;
//   * the ray accumulators (Faraday depth, I, Q, U) are plain registers; the Faraday depth is carried
//     pre-multiplied by 2 lambda^2 (the rotation angle), -Pi(p) is applied once after the march, and
//     the rotation uses short Taylor kernels whenever the whole warp's angle is below 1/8 rad;
//   * RL = 16 / 8: the 2 / 4 lane groups march consecutive segments of the ray; when the segments are
//     joined, each outer one is rotated by the total Faraday depth of those inside it (the rotation
//     is linear in Q+iU).
// Measured and dropped: L1 prefetch of the lines two steps ahead, by CCTL.PF1 or by a load per line
// (2.44 / 2.30 vs 2.31 ms: what the prefetch hides it costs in issue slots); prefetching the weights
// into registers (spills, 5.9 ms); 12 warps / SM at 168 registers (2.70 vs 2.31 ms); first version:
// 128-thread blocks at 72 registers, two realisation groups per warp.  A persistent variant whose blocks
// take adjacent pixels from a per-SM work list (to turn the neighbouring rays' shared vertex lines into L1
// hits: L1 hit rate is 13 % with one-pixel blocks placed by the hardware scheduler): 2.18 vs 2.05 ms at
// C3, 5.79 vs 5.67 ms on a C4 shard -- the atomics, the extra barrier and 60 bytes of spills cost more
// than the L2 round trips they save.  A geometry cache (weights, cell and load mask of every (pixel, step)
// built once per grid / observer / direction set, 80 B per sample, read back by phase A instead of being
// recomputed): bit-identical maps, but 2.23 vs 2.16 ms at 64 realisations, 1.10 vs 1.14 at 32, 0.56 vs 0.55
// at 8 -- phase A is bound by its block-wide barrier and the (n_e, CRE) gathers, not by the arithmetic.
// RL = 8 / 16 (several segments per warp, L1 data pipe at 80 %): the step slot carrying the three cell
// fractions and each lane rebuilding the eight weights itself (24 instead of 64 bytes of shared reads per
// step, 21 more FP64 instructions): 0.494 vs 0.496 ms at RL 8, 0.735 vs 0.712 at RL 16 -- dropped.
// Numeric definition: oracle/los.py.
#include <cstdlib>
#include <type_traits>
#include "los.cuh"

// resident warps per SM the register budget is cut for (16: 128 registers, no spills; 20: 96
// registers and a few spilled words outside the inner dependency chains -- faster)
#ifndef IMB_LOS_WARPS_PER_SM
#define IMB_LOS_WARPS_PER_SM 20
#endif

namespace imb {

struct __align__(16) RlPix {   // per block, in shared memory
  // sky basis by its angles: n = (st cp, st sp, ct), north = (-ct cp, -ct sp, st), east = (-sp, cp, 0)
  double st, ct, cp, sp;
  double n[3];
  int j_lo, j_hi;
};

struct __align__(16) RlGeo {   // per (segment, step slot): everything realisation-independent
  double w[8];                 // trilinear weight of each vertex slot
  double ne, cre;              // -0.812 n_e dl [2 lambda^2]; CRE K(p) [dl]
  int base;                    // cell index of the slot-(0,0,0) vertex (the all-even one)
  int code;                    // bits 0-7: slots to load (bit 4 sx + 2 sy + sz); 8-10: parity of ix, iy, iz;
  double dm;                   // n_e dl of the step (shared-n_e dispersion measure); bit 11 of code: inside the box
  int pad[4];                  // 112 bytes: conflict-free 128-bit stores
};

// three predicated 64-bit loads (the x, y, z components of one vertex slot) in program order
template <int RL>
__device__ __forceinline__ void load_slot(double& x, double& y, double& z, const double* p, int flag) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t"
      "@q ld.global.nc.f64 %0, [%4];\n\t"
      "@q ld.global.nc.f64 %1, [%4+%5];\n\t"
      "@q ld.global.nc.f64 %2, [%4+%6];\n\t}"
      : "+d"(x), "+d"(y), "+d"(z)
      : "r"(flag), "l"(p), "n"(RL * 8), "n"(RL * 16));
}

// shared-memory reads by 32-bit address, volatile: they stay where the march puts them (ahead of the
// arithmetic they feed, so one latency is exposed per step instead of one per use)
template <int OFF>
__device__ __forceinline__ double2 lds_d2(unsigned int addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ int2 lds_i2(unsigned int addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
  return v;
}

// kernel parameters: the LOS arguments plus per-launch constants as constant-bank operands
struct RlArgs {
  LosArgs a;
  double ts[4], tc[5];   // Taylor coefficients of the small-angle rotation
  double inv_dr;
  int sx, sy;            // cell strides of the x and y axes (launch_los_rl guarantees ncell < 2^31)
  int n_groups;
};

// sin and cos for |x| <= 1/8: Taylor to x^9 / x^10 (truncation below 3e-17 relative), Horner form
__device__ __forceinline__ void sincos_small(const RlArgs& ra, double x, double& sn, double& cs) {
  const double z = x * x;
  double ps = fma(ra.ts[3], z, ra.ts[2]), pc = fma(ra.tc[4], z, ra.tc[3]);
  ps = fma(ps, z, ra.ts[1]); pc = fma(pc, z, ra.tc[2]);
  ps = fma(ps, z, ra.ts[0]); pc = fma(pc, z, ra.tc[1]);
  pc = fma(pc, z, ra.tc[0]);
  sn = fma(x * z, ps, x);
  cs = fma(z, pc, 1.0);
}

// sample j of the ray: grid coordinates and the in-box test, bit-identical to
// oracle/los.py:sample_geometry (no fma on this path)
struct RlCell { double px, py, pz, ux, uy, uz; int ix, iy, iz; bool inside; };
__device__ __forceinline__ RlCell rl_cell(const LosArgs& a, const RlPix& pc, int j, bool valid) {
  RlCell c;
  const double r = __dmul_rn((double)j + 0.5, a.dr);
  c.px = __dadd_rn(a.obs[0], __dmul_rn(r, pc.n[0]));
  c.py = __dadd_rn(a.obs[1], __dmul_rn(r, pc.n[1]));
  c.pz = __dadd_rn(a.obs[2], __dmul_rn(r, pc.n[2]));
  c.ux = __dmul_rn(__dsub_rn(c.px, a.g.lo[0]), a.g.inv_step[0]);
  c.uy = __dmul_rn(__dsub_rn(c.py, a.g.lo[1]), a.g.inv_step[1]);
  c.uz = __dmul_rn(__dsub_rn(c.pz, a.g.lo[2]), a.g.inv_step[2]);
  c.inside = valid && c.ux >= 0.0 && c.ux <= (double)(a.g.n[0] - 1) && c.uy >= 0.0 &&
             c.uy <= (double)(a.g.n[1] - 1) && c.uz >= 0.0 && c.uz <= (double)(a.g.n[2] - 1);
  if (!c.inside) c.ux = c.uy = c.uz = 0.0;
  c.ix = min((int)c.ux, a.g.n[0] - 2);
  c.iy = min((int)c.uy, a.g.n[1] - 2);
  c.iz = min((int)c.uz, a.g.n[2] - 2);
  return c;
}

// SHARED: every scalar grid present (n_e, CRE) is one grid for all realisations -- the usual case;
// the per-realisation gathers of RandomThermalElectrons are then compiled out of the march.
// G: warps (realisation groups) per block.
template <int RL, bool HAS_NE, bool HAS_CRE, bool VAR_P, bool P3, bool SHARED, int G>
__global__ void __launch_bounds__(32 * G, IMB_LOS_WARPS_PER_SM / G)
los_rl_kernel(const __grid_constant__ RlArgs ra) {
  const LosArgs& a = ra.a;
  constexpr int S = 32 / RL;                 // ray segments marched side by side by one warp
  constexpr int CH = G * RL;                 // steps per segment and chunk
  constexpr unsigned int FULL = 0xffffffffu;
  constexpr unsigned int GEO = sizeof(RlGeo);
  static_assert(sizeof(RlGeo) == 112, "RlGeo layout is addressed by byte offsets below");
  __shared__ RlPix s_pix;
  // [segment][step]; + 1: the march reads one step slot ahead.  With several segments per warp every
  // segment's slots are 16 bytes further round the banks, so the quarter / half warps of one 128-bit
  // read hit different banks (CH * 112 bytes is a multiple of 128: 4-way conflicts without the pad,
  // 18 M of the 35 M shared-load wavefronts at 8 realisations per group)
  constexpr unsigned int SEG_BYTES = CH * GEO + (S > 1 ? 16 : 0);
  __shared__ __align__(16) unsigned char s_geo_raw[S * SEG_BYTES + GEO];
  auto geo_at = [&](int sg, int i) -> RlGeo& {
    return *reinterpret_cast<RlGeo*>(s_geo_raw + sg * SEG_BYTES + i * GEO);
  };
  __shared__ double s_var[VAR_P ? S * CH : 1][3];
  __shared__ double s_dm[G][32];             // shared-n_e dispersion measure, per warp and lane (see phase B)
  __shared__ int2 s_rng[G];                  // first / last in-box step among those each warp prepared
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = lane / RL, rl = lane % RL;
  const int pix = blockIdx.x;
  const int g = blockIdx.y * G + warp;
  const bool active = g < ra.n_groups;       // an idle warp of the last block still helps with phase A
  const int rr = g * RL + rl;
  const bool live = active && rr < a.n_real;
  RlPix& pc = s_pix;

  if (warp == 0) {
    const double nx = a.dirs[3 * pix], ny = a.dirs[3 * pix + 1], nz = a.dirs[3 * pix + 2];
    // conservative in-box step range (the exact in-box test is still done per step): lane k < 3
    // intersects the ray with the slab of axis k
    double tmin = 0.0, tmax = a.dr * (double)a.n_steps;
    if (lane < 3) {
      const double nk = lane == 0 ? nx : (lane == 1 ? ny : nz);
      const double lo = a.g.lo[lane], hi = lo + (double)(a.g.n[lane] - 1) / a.g.inv_step[lane];
      if (nk != 0.0) {
        const double t1 = (lo - a.obs[lane]) / nk, t2 = (hi - a.obs[lane]) / nk;
        tmin = fmax(tmin, fmin(t1, t2));
        tmax = fmin(tmax, fmax(t1, t2));
      } else if (a.obs[lane] < lo || a.obs[lane] > hi) {
        tmax = -1.0;
      }
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      tmin = fmax(tmin, __shfl_xor_sync(FULL, tmin, o));
      tmax = fmin(tmax, __shfl_xor_sync(FULL, tmax, o));
    }
    if (lane == 0) {
      const double sth = sqrt(nx * nx + ny * ny);
      pc.n[0] = nx; pc.n[1] = ny; pc.n[2] = nz;
      pc.st = sth; pc.ct = nz;
      pc.cp = sth > 0.0 ? nx / sth : 1.0;
      pc.sp = sth > 0.0 ? ny / sth : 0.0;
      int j_lo = 0, j_hi = -1;
      if (tmax >= tmin) {
        j_lo = max(0, (int)floor(tmin * ra.inv_dr - 0.5) - 1);
        j_hi = min(a.n_steps - 1, (int)ceil(tmax * ra.inv_dr - 0.5) + 1);
      }
      pc.j_lo = j_lo; pc.j_hi = j_hi;
    }
  }
  s_dm[warp][lane] = 0.0;
  __syncthreads();

  const int seg_len = (pc.j_hi - pc.j_lo + 1 + S - 1) / S;     // <= 0 when the ray misses the box
  const int sy = ra.sy, sx = ra.sx;
  const bool ne_shared = HAS_NE && (SHARED || a.ne_row == 0), cre_shared = HAS_CRE && (SHARED || a.cre_row == 0);
  constexpr long long CS = 3 * RL;                       // doubles per vertex of the wide layout
  const double* __restrict__ Bg = a.B + (long long)(active ? g : 0) * a.g.ncell * CS + rl;
  // shared-memory byte address of this lane group's (segment's) first step slot
  const unsigned int geo_seg = (unsigned int)__cvta_generic_to_shared(&geo_at(seg, 0));

  const double st = pc.st, ct = pc.ct, cp = pc.cp, sp = pc.sp;
  // the Faraday depth is carried as the rotation angle 2 lambda^2 phi when it rotates anything
  double carry = 0.0, dm = 0.0, I = 0.0, Q = 0.0, U = 0.0;
  // the resident cell: v[slot][component], slot = 4 (X&1) + 2 (Y&1) + (Z&1)
  double v[8][3];
#pragma unroll
  for (int s = 0; s < 8; ++s) v[s][0] = v[s][1] = v[s][2] = 0.0;

  for (int cb = 0; cb < seg_len; cb += CH) {
    // ---- phase A: lane (warp, seg, rl) prepares step cb + warp RL + rl of its segment ------------
    {
      const int istep = warp * RL + rl;                    // step slot within the chunk
      const int off = cb + istep;
      const int j_hi = pc.j_hi;
      const int j = pc.j_lo + seg * seg_len + off;
      const RlCell c = rl_cell(a, pc, j, off < seg_len && j <= j_hi);
      const int ix = c.ix, iy = c.iy, iz = c.iz;
      const bool inside = c.inside;
      Cell cc;
      cc.base = ((long long)ix * a.g.n[1] + iy) * a.g.n[2] + iz;
      cc.sy = sy; cc.sx = sx;
      cc.tx = c.ux - (double)ix; cc.ty = c.uy - (double)iy; cc.tz = c.uz - (double)iz;
      double ne_s = 0.0, cre_s = 0.0;
      if (inside) {
        if (HAS_NE && HAS_CRE && a.necre) {
          interp_pair(a.necre, cc, ne_s, cre_s);
        } else {
          if (ne_shared) ne_s = interp_scalar(a.ne, cc);
          if (cre_shared) cre_s = interp_scalar(a.cre, cc);
        }
      }
      double kp = a.sync_k;
      if (HAS_CRE && VAR_P) {
        const double R = sqrt(c.px * c.px + c.py * c.py);
        const double p = a.alpha - a.beta * R - a.theta * c.pz;
        kp = inside ? sync_coeff_var(a, p) : 0.0;
        s_var[seg * CH + istep][0] = kp;
        s_var[seg * CH + istep][1] = (p + 1.0) / (p + 7.0 / 3.0);
        s_var[seg * CH + istep][2] = (p + 1.0) * 0.25;
      }
      // which slots hold a different vertex than after the previous step of this segment (its cell is
      // recomputed here: cheaper than passing it between lanes and warps)
      const int qx = ix & 1, qy = iy & 1, qz = iz & 1;
      int code = 0;
      if (inside) {
        int cx = 3, cy = 3, cz = 3;                        // bit s: the vertex of parity s changed
        const RlCell pr = rl_cell(a, pc, j - 1, off > 0);
        if (pr.inside) {
          cx = (ix + qx != pr.ix + (pr.ix & 1) ? 1 : 0) | (ix + 1 - qx != pr.ix + 1 - (pr.ix & 1) ? 2 : 0);
          cy = (iy + qy != pr.iy + (pr.iy & 1) ? 1 : 0) | (iy + 1 - qy != pr.iy + 1 - (pr.iy & 1) ? 2 : 0);
          cz = (iz + qz != pr.iz + (pr.iz & 1) ? 1 : 0) | (iz + 1 - qz != pr.iz + 1 - (pr.iz & 1) ? 2 : 0);
        }
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (((cx >> (s >> 2)) | (cy >> ((s >> 1) & 1)) | (cz >> (s & 1))) & 1) code |= 1 << s;
        code |= (qx << 8) | (qy << 9) | (qz << 10) | (1 << 11);
      }
      // the slot of parity s along an axis holds the vertex i (weight 1 - t) when s == i & 1, else i + 1 (t)
      RlGeo& ge = geo_at(seg, istep);
      const double wx[2] = {qx ? cc.tx : 1.0 - cc.tx, qx ? 1.0 - cc.tx : cc.tx};
      const double wy[2] = {qy ? cc.ty : 1.0 - cc.ty, qy ? 1.0 - cc.ty : cc.ty};
      const double wz[2] = {qz ? cc.tz : 1.0 - cc.tz, qz ? 1.0 - cc.tz : cc.tz};
#pragma unroll
      for (int s = 0; s < 8; ++s) ge.w[s] = wx[s >> 2] * wy[(s >> 1) & 1] * wz[s & 1];
      const double dlf = a.dl_pc * ((HAS_NE && HAS_CRE) ? 2.0 * a.lam2 : 1.0);
      // the realisation-independent leading factors are folded in here:
      // -0.812 n_e dl [2 lambda^2] (Faraday; the sign of B_par = -B.n rides along) and CRE K(p) [dl]
      ge.ne = -(__dmul_rn(0.812, ne_s) * dlf);
      ge.cre = P3 ? cre_s * kp * a.dl_cm : cre_s * kp;
      ge.base = (int)cc.base + qx * sx + qy * sy + qz;
      ge.code = code;
      ge.dm = ne_s * a.dl_pc;
      // in-box steps among those this warp prepared, any segment
      const unsigned int mask = __ballot_sync(FULL, inside);
      unsigned int u = 0;
#pragma unroll
      for (int q = 0; q < S; ++q) u |= (mask >> (q * RL)) & ((RL < 32 ? (1u << RL) : 0u) - 1u);
      if (lane == 0) s_rng[warp] = u ? make_int2(warp * RL + __ffs(u) - 1, warp * RL + 31 - __clz(u))
                                     : make_int2(CH, -1);
    }
    __syncthreads();

    // ---- phase B: march the in-box steps of this chunk, one realisation per lane ----------------
    // Software pipeline: the physics of step k-1 and the interpolation of step k are independent, so
    // they sit in ONE basic block; the loads of step k+1 are issued as soon as step k's
    // interpolation has consumed the slots.
    int first = CH, last = -1;
#pragma unroll
    for (int w = 0; w < G; ++w) { const int2 r = s_rng[w]; first = min(first, r.x); last = max(last, r.y); }
    if (ne_shared) {
      // dispersion measure: lane l sums the steps = l (mod RL) of its segment in step order, whatever G
      // is (the maps of a realisation must not depend on how the ensemble is sharded, bit for bit)
      double acc = s_dm[warp][lane];
#pragma unroll
      for (int w = 0; w < G; ++w) acc += geo_at(seg, w * RL + rl).dm;
      s_dm[warp][lane] = acc;
    }
#ifdef IMB_LOS_DBG_ONLYA
    last = -1;
#endif
    if (active && last >= first) {
      // the in-box steps of a ray are contiguous (the box is convex): march first .. last; a step in
      // between that no segment has in the box carries zero factors and an empty load mask
      int n_left = last - first;                              // steps after the current one
      int sl = seg * CH + first;                              // step slot (VAR_P)
      unsigned int ga = geo_seg + (unsigned int)first * GEO;  // shared address of the current RlGeo
      // (re)load the slots whose vertex changed.  S == 1: the mask is warp-uniform; the usual masks (one
      // face crossed, everything) branch to unpredicated loads of exactly those slots -- a predicated-off
      // LDG still takes an issue slot and a pass through the LSU data pipe (ncu: 65 M of them per launch)
      auto fetch_mask = [&](const int base, const int code, auto mask_c) {
        constexpr int M = decltype(mask_c)::value;
        const int dx = (code & 0x100) ? -sx : sx, dy = (code & 0x200) ? -sy : sy, dz = (code & 0x400) ? -1 : 1;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (!(M & (1 << s))) continue;
          const int idx = base + ((s & 4) ? dx : 0) + ((s & 2) ? dy : 0) + ((s & 1) ? dz : 0);
          const double* p = Bg + (long long)idx * CS;
          if (M == 0x1ff) {                        // generic: predicated per slot
            load_slot<RL>(v[s][0], v[s][1], v[s][2], p, code & (1 << s));
          } else {
            v[s][0] = ldg64(p); v[s][1] = ldg64(p + RL); v[s][2] = ldg64(p + 2 * RL);
          }
        }
      };
      auto fetch = [&](const int2 bc) {
        const int base = bc.x, code = bc.y;
#ifdef IMB_LOS_DBG_NOLOAD
        if (base >= 0) return;
#endif
        if (S == 1) {
          const int m = code & 0xff;
          if (m == 0) return;
          if (m == 0x55) fetch_mask(base, code, std::integral_constant<int, 0x55>());        // z face
          else if (m == 0xaa) fetch_mask(base, code, std::integral_constant<int, 0xaa>());
          else if (m == 0x33) fetch_mask(base, code, std::integral_constant<int, 0x33>());   // y face
          else if (m == 0xcc) fetch_mask(base, code, std::integral_constant<int, 0xcc>());
          else if (m == 0x0f) fetch_mask(base, code, std::integral_constant<int, 0x0f>());   // x face
          else if (m == 0xf0) fetch_mask(base, code, std::integral_constant<int, 0xf0>());
          else if (m == 0xff) fetch_mask(base, code, std::integral_constant<int, 0xff>());
          else fetch_mask(base, code, std::integral_constant<int, 0x1ff>());
        } else {
          fetch_mask(base, code, std::integral_constant<int, 0x1ff>());
        }
      };
      // per-realisation scalar grids (RandomThermalElectrons): same slots, same weights
      auto slot_scalar = [&](const double* f, const int2 bc, const double* w) {
        const int dx = (bc.y & 0x100) ? -sx : sx, dy = (bc.y & 0x200) ? -sy : sy, dz = (bc.y & 0x400) ? -1 : 1;
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int idx = bc.x + ((s & 4) ? dx : 0) + ((s & 2) ? dy : 0) + ((s & 1) ? dz : 0);
          acc = fma(w[s], __ldg(f + idx), acc);
        }
        return acc;
      };
      // what a step hands from its interpolation to its physics
      struct Samp { double bx, by, bz, nf, ck, ne_r, cre_r; int sl; };
      // trilinear interpolation of step (ga, bc): one FMA chain per component over the 8 slots
      auto interp = [&](const int2 bc) {
        const double2 w01 = lds_d2<0>(ga), w23 = lds_d2<16>(ga), w45 = lds_d2<32>(ga), w67 = lds_d2<48>(ga);
        const double2 nc = lds_d2<64>(ga);                    // (-0.812 n_e dl [2 lambda^2], CRE K(p) [dl])
        Samp o;
        double e[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          e[k] = w01.x * v[0][k];
          e[k] = fma(w01.y, v[1][k], e[k]);
          e[k] = fma(w23.x, v[2][k], e[k]);
          e[k] = fma(w23.y, v[3][k], e[k]);
          e[k] = fma(w45.x, v[4][k], e[k]);
          e[k] = fma(w45.y, v[5][k], e[k]);
          e[k] = fma(w67.x, v[6][k], e[k]);
          e[k] = fma(w67.y, v[7][k], e[k]);
        }
        o.bx = e[0]; o.by = e[1]; o.bz = e[2];
        o.nf = nc.x; o.ck = nc.y; o.sl = sl;
        o.ne_r = o.cre_r = 0.0;
        if ((HAS_NE && !ne_shared) || (HAS_CRE && !cre_shared)) {
          const bool in_box = live && (S == 1 || ((bc.y >> 11) & 1));    // S > 1: another segment's step
          const double w[8] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y};
          if (HAS_NE && !ne_shared) o.ne_r = in_box ? slot_scalar(a.ne + a.ne_row * rr, bc, w) : 0.0;
          if (HAS_CRE && !cre_shared) o.cre_r = in_box ? slot_scalar(a.cre + a.cre_row * rr, bc, w) : 0.0;
        }
        return o;
      };
      auto physics = [&](const Samp& p) {
#ifdef IMB_LOS_DBG_NOPHYS
        I += p.bx + p.by + p.bz;
        return;
#endif
        const double bh = fma(p.bx, cp, p.by * sp);                     // horizontal part along the ray's azimuth
        double theta = 0.0;
        if (HAS_NE) {
          double nf = p.nf;
          if (!ne_shared) {
            dm += p.ne_r * a.dl_pc;
            nf = -(__dmul_rn(0.812, p.ne_r) * (a.dl_pc * ((HAS_NE && HAS_CRE) ? 2.0 * a.lam2 : 1.0)));
          }
          const double bn_ray = fma(st, bh, ct * p.bz);      // B . n = -B_par
          const double f = nf * bn_ray;
          theta = fma(0.5, f, carry);
          carry += f;
        }
        if (HAS_CRE) {
          double ck = p.ck;                                  // CRE K(p) dl (P3) or CRE K(p)
          if (!cre_shared) {
            const double kp = VAR_P ? s_var[p.sl][0] : a.sync_k;
            ck = P3 ? p.cre_r * kp * a.dl_cm : p.cre_r * kp;
          }
          const double bn = fma(st, p.bz, -(ct * bh));       // B . north
          const double be = fma(p.by, cp, -(p.bx * sp));     // B . east
          const double bn2 = bn * bn;
          const double bp2 = fma(be, be, bn2), dq = fma(-be, be, bn2), du = (bn + bn) * be;
          // (Q + iU) accumulates em/Bperp^2 (B_N + i B_E)^2 e^{i theta} [Pi(p)]; the sign and a constant
          // Pi are applied after the march
          double gq;
          if (P3) {                                          // (p+1)/4 == 1: no pow, no division
            I = fma(ck, bp2, I);
            gq = ck;
          } else {
            const double expo = VAR_P ? s_var[p.sl][2] : a.expo;
            const double em = bp2 > 0.0 ? ck * pow(bp2, expo) * a.dl_cm : 0.0;
            I += em;
            gq = bp2 > 0.0 ? em / bp2 : 0.0;
            if (VAR_P) gq *= s_var[p.sl][1];
          }
          const double qc = gq * dq, uc = gq * du;
          if (HAS_NE) {
            // the short kernels run unconditionally (straight-line code); a warp with any angle above
            // 1/8 rad redoes the step's rotation with the full-range kernel
            double sa, ca;
            sincos_small(ra, theta, sa, ca);
            if (!__all_sync(FULL, fabs(theta) <= 0.125)) sincos_cw(theta, sa, ca);
            Q = fma(qc, ca, fma(-uc, sa, Q));
            U = fma(qc, sa, fma(uc, ca, U));
          } else {
            Q += qc;
            U += uc;
          }
        }
      };
      int2 bc = lds_i2<80>(ga);                               // (base, code) of the current step
      fetch(bc);
      int2 bcn = lds_i2<GEO + 80>(ga);                        // one step ahead
      Samp prev = interp(bc);
      while (n_left > 0) {
        --n_left; ++sl; ga += GEO;
        fetch(bcn);                          // the slots are free: step k+1's loads fly during the physics
        bc = bcn;
        bcn = lds_i2<GEO + 80>(ga);
        physics(prev);
        prev = interp(bc);
      }
      physics(prev);
    }
    __syncthreads();                                     // phase A of the next chunk reuses s_geo
  }

  if (HAS_CRE) {
    const double pf = VAR_P ? -1.0 : -a.pol_frac;
    Q *= pf; U *= pf;
  }
  // ---- join the segments: the outer one starts at the inner one's total Faraday depth ----------
  if (S > 1) {
#pragma unroll
    for (int o = RL; o < 32; o <<= 1) {
      const double cu = __shfl_down_sync(FULL, carry, o);
      const double du = __shfl_down_sync(FULL, dm, o);
      const double iu = __shfl_down_sync(FULL, I, o);
      const double qu = __shfl_down_sync(FULL, Q, o);
      const double uu = __shfl_down_sync(FULL, U, o);
      if (HAS_NE && HAS_CRE) {
        double sa, ca;
        sincos(carry, &sa, &ca);                           // carry is already the rotation angle
        Q += qu * ca - uu * sa;
        U += qu * sa + uu * ca;
      } else {
        Q += qu; U += uu;
      }
      I += iu; dm += du; carry += cu;
    }
  }
  if (ne_shared) {
    double part = s_dm[warp][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) part += __shfl_xor_sync(FULL, part, o);
    dm = part;
  }
  if (seg == 0 && live) {
    const long long o = out_index(a, rr, pix);
    if (HAS_NE) {
      if (a.out_dm) a.out_dm[o] = dm;
      if (a.out_fd) a.out_fd[o] = (HAS_NE && HAS_CRE) ? carry / (2.0 * a.lam2) : carry;
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
  RlArgs ra;
  ra.a = a;
  // Taylor coefficients of sin (x^3 .. x^9) and cos (x^2 .. x^10)
  const double ts[4] = {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0};
  const double tc[5] = {-0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0};
  for (int k = 0; k < 4; ++k) ra.ts[k] = ts[k];
  for (int k = 0; k < 5; ++k) ra.tc[k] = tc[k];
  ra.inv_dr = 1.0 / a.dr;
  ra.sy = a.g.n[2];
  ra.sx = a.g.n[1] * a.g.n[2];
  ra.n_groups = n_groups;
  const bool shared = (!HAS_NE || a.ne_row == 0) && (!HAS_CRE || a.cre_row == 0);
  if (RL == 32 && shared && n_groups >= 2) {
    // two realisation groups of a pixel share one geometry pass
    dim3 grid(a.m, (n_groups + 1) / 2);
    los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, true, (RL == 32 ? 2 : 1)><<<grid, 64, 0, s>>>(ra);
    return IMB_OK;
  }
  dim3 grid(a.m, n_groups);
  if (shared) los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, true, 1><<<grid, 32, 0, s>>>(ra);
  else los_rl_kernel<RL, HAS_NE, HAS_CRE, VAR_P, P3, false, 1><<<grid, 32, 0, s>>>(ra);
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
  if ((a.n_real + rl - 1) / rl > 65535) return IMB_EUNSUPPORTED;   // grid.y
  if (rl == 32) return launch_rl<32>(a, has_ne, has_cre, var_p, s);
  if (rl == 16) return launch_rl<16>(a, has_ne, has_cre, var_p, s);
  return launch_rl<8>(a, has_ne, has_cre, var_p, s);
}

}  // namespace imb
