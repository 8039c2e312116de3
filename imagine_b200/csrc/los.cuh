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

// volatile: the compiler keeps these loads in program order, ahead of the arithmetic that follows
__device__ __forceinline__ double ldg64(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
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

// sin and cos of a moderate argument with constant-bank coefficients (kSinC / kCosC): two-constant
// Cody-Waite reduction by pi/2 + the Taylor kernels of common.cuh, <= 2 ulp for |x| < 1e5.  The CUDA
// library sincos() materialises every coefficient through a pair of uniform moves (~34 UMOV per
// call in the ray-march loop, ncu source page); huge or non-finite arguments still go there.
__device__ __forceinline__ void sincos_cw(double x, double& sn, double& cs) {
  if (!(fabs(x) < 1.0e5)) { sincos(x, &sn, &cs); return; }
  const int q = __double2int_rn(x * 0.6366197723675814);
  const double kf = (double)q;
  double r = fma(-kf, 1.5707963267948966, x);
  r = fma(-kf, 6.123233995736766e-17, r);
  const double z = r * r;
  double ps = kSinC[6], pc = kCosC[7];
#pragma unroll
  for (int k = 5; k >= 0; --k) ps = fma(ps, z, kSinC[k]);
#pragma unroll
  for (int k = 6; k >= 0; --k) pc = fma(pc, z, kCosC[k]);
  const double s0 = fma(r * z, ps, r), c0 = fma(z, pc, 1.0);
  const double s1 = (q & 1) ? c0 : s0, c1 = (q & 1) ? s0 : c0;
  sn = (q & 2) ? -s1 : s1;
  cs = ((q + 1) & 2) ? -c1 : c1;
}

__device__ __forceinline__ double sync_coeff_var(const LosArgs& a, double p) {
  return a.kA / (p + 1.0) * tgamma(p * 0.25 + 19.0 / 12.0) * tgamma(p * 0.25 - 1.0 / 12.0) *
         exp(a.kE0 + p * a.kE1);
}

}  // namespace imb
