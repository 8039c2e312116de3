// Shared declarations of the line-of-sight kernels (los.cu, los_il4.cu).
#pragma once
#include "common.cuh"

namespace imb {

struct LosArgs {
  DevGrid g;
  double obs[3];
  double dr, dl_pc, dl_cm, lam2;
  int n_steps, m, n_real;
  // constant-index synchrotron coefficients (beta == theta == 0)
  double sync_k, pol_frac, expo;
  // variable-index: K(p) = kA/(p+1) G(p/4+19/12) G(p/4-1/12) exp(kE0 + p*kE1)
  double alpha, beta, theta, kA, kE0, kE1;
  const double* dirs;
  const double* B;   long long B_row;    // row stride in doubles (0 when rows == 1)
  const double* ne;  long long ne_row;
  const double* cre; long long cre_row;
  const double* necre;                   // packed (n_e, CRE) per vertex when both are shared, else NULL
  double *out_dm, *out_fd, *out_I, *out_Q, *out_U;
  // output addressing: row-major [n_real, m] (out_mp == m, out_dest == 0) or, for a pixel-sharded
  // exchange, [dest][...][n_real][out_mp] with `out_dest` doubles between destination blocks
  int out_mp;
  long long out_dest;
  // fused exchange: destination d's block is written straight into rank d's receive buffer over
  // NVLink; peer_delta[d] = (peer d's buffer base) - (my buffer base), in doubles
  int n_peers;
  long long peer_delta[16];
};

__device__ __forceinline__ long long out_index(const LosArgs& a, int rr, int pix) {
  const int dest = pix / a.out_mp;
  const long long within = (long long)rr * a.out_mp + (pix - dest * a.out_mp);
  if (a.n_peers > 0) return a.peer_delta[dest] + within;
  return (long long)dest * a.out_dest + within;
}

__device__ __forceinline__ double4 ldg256(const double* p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ double lerp(double a, double b, double t) { return fma(t, b - a, a); }

struct Cell {
  long long base;       // linear index of the (ix,iy,iz) corner
  long long sy, sx;     // strides to iy+1, ix+1 (in cells)
  double tx, ty, tz;
};

__device__ __forceinline__ double interp_scalar(const double* __restrict__ f, const Cell& c) {
  const double* p = f + c.base;
  const double c00 = lerp(__ldg(p), __ldg(p + 1), c.tz);
  const double c01 = lerp(__ldg(p + c.sy), __ldg(p + c.sy + 1), c.tz);
  const double c10 = lerp(__ldg(p + c.sx), __ldg(p + c.sx + 1), c.tz);
  const double c11 = lerp(__ldg(p + c.sx + c.sy), __ldg(p + c.sx + c.sy + 1), c.tz);
  return lerp(lerp(c00, c01, c.ty), lerp(c10, c11, c.ty), c.tx);
}

// packed (n_e, CRE) pairs: one 128-bit load per vertex
__device__ __forceinline__ void interp_pair(const double* __restrict__ f, const Cell& c, double& ne,
                                            double& cre) {
  const double2* p = reinterpret_cast<const double2*>(f) + c.base;
  const double2 a0 = __ldg(p), a1 = __ldg(p + 1);
  const double2 b0 = __ldg(p + c.sy), b1 = __ldg(p + c.sy + 1);
  const double2 c0 = __ldg(p + c.sx), c1 = __ldg(p + c.sx + 1);
  const double2 d0 = __ldg(p + c.sx + c.sy), d1 = __ldg(p + c.sx + c.sy + 1);
  ne = lerp(lerp(lerp(a0.x, a1.x, c.tz), lerp(b0.x, b1.x, c.tz), c.ty),
            lerp(lerp(c0.x, c1.x, c.tz), lerp(d0.x, d1.x, c.tz), c.ty), c.tx);
  cre = lerp(lerp(lerp(a0.y, a1.y, c.tz), lerp(b0.y, b1.y, c.tz), c.ty),
             lerp(lerp(c0.y, c1.y, c.tz), lerp(d0.y, d1.y, c.tz), c.ty), c.tx);
}

template <int BS>
__device__ __forceinline__ void zpair(const double* __restrict__ p, double tz, double& x,
                                      double& y, double& z) {
  if (BS == 4) {
    const double4 a = ldg256(p), b = ldg256(p + 4);
    x = lerp(a.x, b.x, tz); y = lerp(a.y, b.y, tz); z = lerp(a.z, b.z, tz);
  } else {
    const double a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
    const double b0 = __ldg(p + 3), b1 = __ldg(p + 4), b2 = __ldg(p + 5);
    x = lerp(a0, b0, tz); y = lerp(a1, b1, tz); z = lerp(a2, b2, tz);
  }
}

template <int BS>
__device__ __forceinline__ void interp_vector(const double* __restrict__ f, const Cell& c,
                                              double& bx, double& by, double& bz) {
  const double* p = f + c.base * BS;
  double x00, y00, z00, x01, y01, z01, x10, y10, z10, x11, y11, z11;
  zpair<BS>(p, c.tz, x00, y00, z00);
  zpair<BS>(p + c.sy * BS, c.tz, x01, y01, z01);
  zpair<BS>(p + c.sx * BS, c.tz, x10, y10, z10);
  zpair<BS>(p + (c.sx + c.sy) * BS, c.tz, x11, y11, z11);
  bx = lerp(lerp(x00, x01, c.ty), lerp(x10, x11, c.ty), c.tx);
  by = lerp(lerp(y00, y01, c.ty), lerp(y10, y11, c.ty), c.tx);
  bz = lerp(lerp(z00, z01, c.ty), lerp(z10, z11, c.ty), c.tx);
}

__device__ __forceinline__ double sync_coeff_var(const LosArgs& a, double p) {
  return a.kA / (p + 1.0) * tgamma(p * 0.25 + 19.0 / 12.0) * tgamma(p * 0.25 - 1.0 / 12.0) *
         exp(a.kE0 + p * a.kE1);
}

}  // namespace imb
