// Stage 1: grid-field evaluation (analytic models + Philox Gaussian components, fused with
// the same-type sum), HEALPix RING directions, TestSimulator sampling.
//
// Reference arithmetic restated by oracle/fields.py; file:line citations in
// include/imagine_b200.h.  HBM-write-bound streaming kernels: one thread owns a PAIR of
// adjacent cells so that every Philox block (one Box-Muller pair) is consumed whole and
// stores are 16/32-byte vectors, coalesced across the warp.
#include "common.cuh"

namespace imb {

// range-reduction tables of the Box-Muller transform (common.cuh), filled once by the host
__device__ double2 g_normal_tab[kTabSize];

// every block stages the tables in shared memory (random 16-byte reads: conflicts, not cache misses)
__device__ __forceinline__ void stage_normal_tab(double2* s_tab) {
  for (int k = threadIdx.x; k < kTabSize; k += blockDim.x) s_tab[k] = g_normal_tab[k];
  __syncthreads();
}

struct FieldArgs {
  DevGrid g;
  int nfields;
  int rows;                 // 1 or n_real
  int kind[IMB_MAX_FIELDS];
  unsigned int stream[IMB_MAX_FIELDS];
  int seed_row[IMB_MAX_FIELDS];
  double p[IMB_MAX_FIELDS][24];
  int n_real;                       // seeds row length
  const unsigned long long* seeds;  // device [n_seed_rows, n_real]
  // fast path of the magnetic-field kernels: constant fields followed by ONE Gaussian field (or
  // any order when there are just two: the sum commutes).  csum = the constants added in list
  // order on the host (same IEEE additions the generic loop does), gauss = index of the Gaussian.
  int simple, has_const, gauss;
  double csum[3];
  // interleaved output: lanes per group of the BUFFER (>= the RL of the thread mapping) and the first
  // lane this call writes -- several parameter points share one wide buffer (imb_field_eval_lanes)
  int out_rl, lane_off;
};

// the six normals of cell pair q of one realisation: v = [csum +] (z sigma + mu)
__device__ __forceinline__ void simple_pair(const FieldArgs& a, unsigned long long seed, unsigned int q,
                                            const double2* tab, double* v) {
  const double mu = a.p[a.gauss][0], sigma = a.p[a.gauss][1];
  const unsigned int stream = a.stream[a.gauss];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    double z0, z1;
    philox_normal_pair(seed, stream, 3ull * q + b, tab, z0, z1);
    v[2 * b] = __dadd_rn(__dmul_rn(z0, sigma), mu);
    v[2 * b + 1] = __dadd_rn(__dmul_rn(z1, sigma), mu);
  }
  if (a.has_const) {
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = __dadd_rn(a.csum[k % 3], v[k]);
  }
}

__device__ __forceinline__ void cell_xyz(const DevGrid& g, long long c, double& x, double& y,
                                         double& z) {
  const int iz = (int)(c % g.n[2]);
  const long long t = c / g.n[2];
  const int iy = (int)(t % g.n[1]);
  const int ix = (int)(t / g.n[1]);
  x = vertex_coord(g, 0, ix);
  y = vertex_coord(g, 1, iy);
  z = vertex_coord(g, 2, iz);
}

// WMAP-3yr logarithmic spiral arm field (the model hammurabiX's `breg_lsa` implements; the
// reference only routes its four parameters to the binary, fields/hamx/breg_lsa.py:20-31):
//   psi = psi0 + psi1 ln(R/r0), chi = chi0 tanh(z/z0), zero outside r_min < R < r_max,
//   B = b0 (sin psi cos chi R^ + cos psi cos chi phi^ + sin chi z^)
//   p = b0 [uG], psi0, psi1, chi0 [rad], r0, z0, r_min, r_max [kpc]
__device__ __forceinline__ void lsa_field(const double* p, double x, double y, double z, double* b) {
  const double R = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
  if (!(R > p[6]) || !(R < p[7])) { b[0] = b[1] = b[2] = 0.0; return; }
  const double cp = x / R, sp = y / R;
  const double psi = __dadd_rn(p[1], __dmul_rn(p[2], log(R / p[4])));
  const double chi = __dmul_rn(p[3], tanh(z / p[5]));
  const double cc = cos(chi);
  const double br = __dmul_rn(__dmul_rn(p[0], sin(psi)), cc);
  const double bphi = __dmul_rn(__dmul_rn(p[0], cos(psi)), cc);
  b[0] = __dsub_rn(__dmul_rn(br, cp), __dmul_rn(bphi, sp));
  b[1] = __dadd_rn(__dmul_rn(br, sp), __dmul_rn(bphi, cp));
  b[2] = __dmul_rn(p[0], sin(chi));
}

// cell pair of a position-dependent regular model
template <typename F>
__device__ __forceinline__ void model_pair(const DevGrid& g, const double* p, long long c0, bool has2, double* t,
                                           F model) {
  double x, y, z;
  cell_xyz(g, c0, x, y, z);
  model(p, x, y, z, t);
  if (has2) {
    cell_xyz(g, c0 + 1, x, y, z);
    model(p, x, y, z, t + 3);
  } else {
    t[3] = t[4] = t[5] = 0.0;
  }
}

// Jansson & Farrar (2012) regular field (the model hammurabiX's `breg_jf12` implements; the reference
// only routes its 20 parameters to the binary, fields/hamx/breg_jf12.py:17-39).  Same operations as
// oracle/fields.py:jf12_B.  p = b_arm1..7, b_ring, h_disk, w_disk, Bn, Bs, rn, rs, wh, z0, B0_X,
// Xtheta [deg], rpc_X, r0_X.
__device__ __forceinline__ double jf12_logistic(double v, double h, double w) {
  return 1.0 / (1.0 + exp(-2.0 * (fabs(v) - h) / w));
}

__device__ __noinline__ void jf12_field(const double* p, double x, double y, double z, double* b) {
  const double rc[8] = {5.1, 6.3, 7.1, 8.3, 9.8, 11.4, 12.7, 15.5};
  const double fa[8] = {0.130, 0.165, 0.094, 0.122, 0.13, 0.118, 0.084, 0.156};
  const double R = sqrt(x * x + y * y), rsph = sqrt(R * R + z * z);
  b[0] = b[1] = b[2] = 0.0;
  if (!(rsph >= 1.0) || !(R <= 20.0) || !(R > 0.0)) return;
  const double cp = x / R, sp = y / R, phi = atan2(y, x);
  const double pitch = 11.5 * 0.017453292519943295;
  const double sin_i = sin(pitch), cos_i = cos(pitch), tan_c = tan(1.5707963267948966 - pitch);
  double barm[8], s8 = 0.0;
  for (int i = 0; i < 7; ++i) { barm[i] = p[i]; s8 += fa[i] * p[i]; }
  barm[7] = -s8 / fa[7];
  const double lfd = jf12_logistic(z, p[8], p[9]);
  double r_negx = R * exp(-(phi - 3.141592653589793) / tan_c);
  if (r_negx > rc[7]) r_negx = R * exp(-(phi + 3.141592653589793) / tan_c);
  if (r_negx > rc[7]) r_negx = R * exp(-(phi + 3.0 * 3.141592653589793) / tan_c);
  double arm = 0.0;
  for (int i = 7; i >= 0; --i) if (r_negx < rc[i]) arm = barm[i];
  double bdr = 0.0, bdp = 0.0;
  if (R > 5.0) {
    const double bm = arm * (5.0 / R);
    bdr = bm * sin_i; bdp = bm * cos_i;
  } else if (R > 3.0) {
    bdp = p[7];
  }
  bdr *= 1.0 - lfd; bdp *= 1.0 - lfd;
  const double btor = exp(-fabs(z) / p[15]) * lfd *
                      (z >= 0.0 ? p[10] * (1.0 - jf12_logistic(R, p[12], p[14]))
                                : p[11] * (1.0 - jf12_logistic(R, p[13], p[14])));
  const double xth = p[17] * 0.017453292519943295, tan_x = tan(xth);
  const double rc_bound = p[18] + fabs(z) / tan_x;
  const bool outer = R >= rc_bound;
  const double rp = outer ? R - fabs(z) / tan_x : R * p[18] / rc_bound;
  const double ratio = rp / R;
  const double bxm = p[16] * exp(-rp / p[19]) * (outer ? ratio : ratio * ratio);
  double thx = xth;
  if (!outer) thx = (R - rp > 0.0) ? atan2(fabs(z), R - rp) : 1.5707963267948966;
  const double bxr = bxm * cos(thx) * (z >= 0.0 ? 1.0 : -1.0), bxz = bxm * sin(thx);
  const double br = bdr + bxr, bphi = bdp + btor;
  b[0] = br * cp - bphi * sp;
  b[1] = br * sp + bphi * cp;
  b[2] = bxz;
}

__device__ __forceinline__ void lsa_pair(const DevGrid& g, const double* p, long long c0, bool has2,
                                         double* t) {
  double x, y, z;
  cell_xyz(g, c0, x, y, z);
  lsa_field(p, x, y, z, t);
  if (has2) {
    cell_xyz(g, c0 + 1, x, y, z);
    lsa_field(p, x, y, z, t + 3);
  } else {
    t[3] = t[4] = t[5] = 0.0;
  }
}

// ---- magnetic field: [rows, ncell, BS] ------------------------------------------------
template <int BS, bool VEC>
__global__ void __launch_bounds__(256) field_b_kernel(const __grid_constant__ FieldArgs a,
                                                       double* __restrict__ out) {
  __shared__ double2 s_tab[kTabSize];
  stage_normal_tab(s_tab);
  const long long npair = (a.g.ncell + 1) >> 1;
  const long long total = npair * a.rows;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(w / npair);
    const long long q = w - (long long)r * npair;
    double v[6];
    bool first = true;
    for (int f = 0; f < a.nfields; ++f) {
      double t[6];
      if (a.kind[f] == IMB_FIELD_CONST_B) {
#pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = a.p[f][k % 3];
      } else if (a.kind[f] == IMB_FIELD_LSA_B) {
        lsa_pair(a.g, a.p[f], 2 * q, (2 * q + 1) < a.g.ncell, t);
      } else if (a.kind[f] == IMB_FIELD_JF12_B) {
        model_pair(a.g, a.p[f], 2 * q, (2 * q + 1) < a.g.ncell, t, jf12_field);
      } else {  // IMB_FIELD_GAUSS_B: flat normal index n = 3*cell + comp, block n >> 1
        const unsigned long long seed = a.seeds[(size_t)a.seed_row[f] * a.n_real + r];
        const double mu = a.p[f][0], sigma = a.p[f][1];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          double z0, z1;
          philox_normal_pair(seed, a.stream[f], (unsigned long long)(3 * q + b), s_tab, z0, z1);
          t[2 * b] = __dadd_rn(__dmul_rn(z0, sigma), mu);
          t[2 * b + 1] = __dadd_rn(__dmul_rn(z1, sigma), mu);
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = first ? t[k] : __dadd_rn(v[k], t[k]);
      first = false;
    }
    const long long c0 = 2 * q;
    const bool has2 = (c0 + 1) < a.g.ncell;
    double* o = out + ((long long)r * a.g.ncell + c0) * BS;
    if (VEC && has2) {
      if (BS == 3) {
        double2* o2 = reinterpret_cast<double2*>(o);
        o2[0] = make_double2(v[0], v[1]);
        o2[1] = make_double2(v[2], v[3]);
        o2[2] = make_double2(v[4], v[5]);
      } else {
        double4* o4 = reinterpret_cast<double4*>(o);
        o4[0] = make_double4(v[0], v[1], v[2], 0.0);
        o4[1] = make_double4(v[3], v[4], v[5], 0.0);
      }
    } else {
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
      if (BS == 4) o[3] = 0.0;
      if (has2) {
        o[BS] = v[3]; o[BS + 1] = v[4]; o[BS + 2] = v[5];
        if (BS == 4) o[BS + 3] = 0.0;
      }
    }
  }
}

// the same-type sum of all magnetic fields for cell pair q of realisation r (any field list)
__device__ __forceinline__ void generic_pair(const FieldArgs& a, unsigned int q, int r, const double2* tab,
                                             double* v) {
  bool first = true;
  for (int f = 0; f < a.nfields; ++f) {
    double t[6];
    if (a.kind[f] == IMB_FIELD_CONST_B) {
#pragma unroll
      for (int k = 0; k < 6; ++k) t[k] = a.p[f][k % 3];
    } else if (a.kind[f] == IMB_FIELD_LSA_B) {
      lsa_pair(a.g, a.p[f], 2ll * q, (2ll * q + 1) < a.g.ncell, t);
    } else if (a.kind[f] == IMB_FIELD_JF12_B) {
      model_pair(a.g, a.p[f], 2ll * q, (2ll * q + 1) < a.g.ncell, t, jf12_field);
    } else {
      const unsigned long long seed = a.seeds[(size_t)a.seed_row[f] * a.n_real + r];
      const double mu = a.p[f][0], sigma = a.p[f][1];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double z0, z1;
        philox_normal_pair(seed, a.stream[f], 3ull * q + b, tab, z0, z1);
        t[2 * b] = __dadd_rn(__dmul_rn(z0, sigma), mu);
        t[2 * b + 1] = __dadd_rn(__dmul_rn(z1, sigma), mu);
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = first ? t[k] : __dadd_rn(v[k], t[k]);
    first = false;
  }
}

// ---- magnetic field, realisation-interleaved layouts ---------------------------------------------
// Thread = (pair of cells, realisation), realisation fastest; the grid stride is a multiple of the
// interleave, so realisation and seed are fixed per thread.  Layouts (group = r / RL):
//   RL 32 | 16 | 8 (b_stride 3 RL): out[group][cell][component][r % RL]: one coalesced 64-bit load per
//                               corner component for the lane-per-realisation LOS kernel (los_rl.cu)
// (Measured and dropped: an 8-wide layout storing (B(cell), B(cell + 1 in z)) pairs so that a z-pair
// of corners is one 128-bit load -- 4.95 vs 5.03 ms, twice the memory.)
// SIMPLE: the list is constants + ONE Gaussian field (the usual case): its own instantiation, so that the
// register budget of the hot kernel is not set by the analytic models of the generic path
// WIDE: the buffer's lane count and this call's first lane come from the arguments (imb_field_eval_lanes);
// otherwise they are the compile-time RL and 0 (runtime strides cost the issue-bound kernel 4 %)
template <int RL, bool SIMPLE, bool WIDE = false>
__global__ void __launch_bounds__(256) field_b_il_kernel(const __grid_constant__ FieldArgs a,
                                                         double* __restrict__ out) {
  __shared__ double2 s_tab[kTabSize];
  stage_normal_tab(s_tab);
  const unsigned int npair_rl = (unsigned int)((a.g.ncell + 1) >> 1) * (unsigned int)RL;
  const int g = blockIdx.y;
  const int rl = (int)(threadIdx.x % RL), r = g * RL + rl;
  const bool live = r < a.rows;
  constexpr bool simple = SIMPLE;
  const unsigned long long seed =
      (simple && live) ? a.seeds[(size_t)a.seed_row[a.gauss] * a.n_real + r] : 0ull;
  for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < npair_rl; w += gridDim.x * blockDim.x) {
    const unsigned int q = w / RL;
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (live) {
      if (SIMPLE) simple_pair(a, seed, q, s_tab, v);
      else generic_pair(a, q, r, s_tab, v);
    }
    const long long c0 = 2ll * q;
    const bool has2 = c0 + 1 < a.g.ncell;
    const int orl = WIDE ? a.out_rl : RL;
    double* o = out + ((long long)g * a.g.ncell + c0) * (3 * orl) + (WIDE ? a.lane_off : 0) + rl;
    o[0] = v[0]; o[orl] = v[1]; o[2 * orl] = v[2];
    if (has2) { o[3 * orl] = v[3]; o[4 * orl] = v[4]; o[5 * orl] = v[5]; }
  }
}

// ---- scalar fields (thermal electrons, cosmic-ray electrons): [rows, ncell] -----------
__device__ __forceinline__ double scalar_model(const FieldArgs& a, int f, double x, double y,
                                               double z) {
  const double* p = a.p[f];
  switch (a.kind[f]) {
    case IMB_FIELD_CONST_NE:
    case IMB_FIELD_CRE_CONST:
      return p[0];
    case IMB_FIELD_EXP_NE: {
      const double R = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
      return __dmul_rn(__dmul_rn(p[0], exp(-R / p[1])), exp(-fabs(z / p[2])));
    }
    case IMB_FIELD_COS_NE: {
      const double fx = __dadd_rn(1.0, cos(__dadd_rn(__dmul_rn(p[1], x), p[2])));
      const double fy = __dadd_rn(1.0, cos(__dadd_rn(__dmul_rn(p[3], y), p[4])));
      const double fz = __dadd_rn(1.0, cos(__dadd_rn(__dmul_rn(p[5], z), p[6])));
      return __dmul_rn(__dmul_rn(__dmul_rn(p[0], fx), fy), fz);
    }
    case IMB_FIELD_CRE_ANALYTIC: {
      const double R = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
      const double ch = cosh(z / p[2]);
      return __dmul_rn(__dmul_rn(p[0], exp(-(R - p[3]) / p[1])), 1.0 / __dmul_rn(ch, ch));
    }
    default:
      return 0.0;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) field_scalar_kernel(const __grid_constant__ FieldArgs a,
                                                            double* __restrict__ out) {
  __shared__ double2 s_tab[kTabSize];
  stage_normal_tab(s_tab);
  const long long npair = (a.g.ncell + 1) >> 1;
  const long long total = npair * a.rows;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(w / npair);
    const long long q = w - (long long)r * npair;
    const long long c0 = 2 * q;
    const bool has2 = (c0 + 1) < a.g.ncell;
    double x0, y0, z0, x1 = 0, y1 = 0, z1 = 0;
    cell_xyz(a.g, c0, x0, y0, z0);
    if (has2) cell_xyz(a.g, c0 + 1, x1, y1, z1);
    double v0 = 0, v1 = 0;
    bool first = true;
    for (int f = 0; f < a.nfields; ++f) {
      double t0, t1;
      if (a.kind[f] == IMB_FIELD_RANDOM_NE) {
        double n0, n1;
        philox_normal_pair(a.seeds[(size_t)a.seed_row[f] * a.n_real + r], a.stream[f],
                           (unsigned long long)q, s_tab, n0, n1);
        t0 = __dadd_rn(__dmul_rn(n0, a.p[f][1]), a.p[f][0]);
        t1 = __dadd_rn(__dmul_rn(n1, a.p[f][1]), a.p[f][0]);
        const double mn = a.p[f][2];
        if (isfinite(mn)) {
          if (t0 < mn) t0 = mn;
          if (t1 < mn) t1 = mn;
        }
      } else {
        t0 = scalar_model(a, f, x0, y0, z0);
        t1 = has2 ? scalar_model(a, f, x1, y1, z1) : 0.0;
      }
      v0 = first ? t0 : __dadd_rn(v0, t0);
      v1 = first ? t1 : __dadd_rn(v1, t1);
      first = false;
    }
    double* o = out + (long long)r * a.g.ncell + c0;
    if (VEC && has2) {
      *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
    } else {
      o[0] = v0;
      if (has2) o[1] = v1;
    }
  }
}

// ---- HEALPix RING pixel centres -> unit vectors (oracle/healpix.py) -------------------
__global__ void healpix_ring_dirs_kernel(int nside, double* __restrict__ dirs) {
  const long long npix = 12ll * nside * nside;
  const long long ncap = 2ll * nside * (nside - 1);
  const double pi = 3.141592653589793;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    double z, sth, phi;
    if (p < ncap) {
      long long i = (1 + (long long)floor(sqrt(1.0 + 2.0 * (double)p))) >> 1;
      if (2 * i * (i - 1) > p) --i;
      if (2 * (i + 1) * i <= p) ++i;
      const long long j = p + 1 - 2 * i * (i - 1);
      const double tmp = __dmul_rn((double)(i * i), 4.0 / (double)npix);
      z = 1.0 - tmp;
      sth = sqrt(__dmul_rn(tmp, 2.0 - tmp));
      phi = __dmul_rn((double)j - 0.5, pi / (2.0 * (double)i));
    } else if (p < npix - ncap) {
      const long long q = p - ncap;
      const long long i = q / (4 * nside) + nside;
      const long long j = q % (4 * nside) + 1;
      const double fodd = ((i + nside) & 1) ? 1.0 : 0.5;
      z = __dmul_rn((double)(2 * nside - i), 2.0 / (3.0 * (double)nside));
      sth = sqrt(__dmul_rn(1.0 - z, 1.0 + z));
      phi = __dmul_rn((double)j - fodd, pi / (2.0 * (double)nside));
    } else {
      const long long q = npix - p;
      long long i = (1 + (long long)floor(sqrt(2.0 * (double)q - 1.0))) >> 1;
      if (2 * i * (i - 1) >= q) --i;
      if (2 * (i + 1) * i < q) ++i;
      const long long j = 4 * i + 1 - (q - 2 * i * (i - 1));
      const double tmp = __dmul_rn((double)(i * i), 4.0 / (double)npix);
      z = -(1.0 - tmp);
      sth = sqrt(__dmul_rn(tmp, 2.0 - tmp));
      phi = __dmul_rn((double)j - 0.5, pi / (2.0 * (double)i));
    }
    double s, c;
    sincos(phi, &s, &c);
    dirs[3 * p + 0] = __dmul_rn(sth, c);
    dirs[3 * p + 1] = __dmul_rn(sth, s);
    dirs[3 * p + 2] = z;
  }
}

// ---- TestSimulator: nearest-vertex sampling of B[axis]*ne ---------------------------------
__device__ __forceinline__ int nearest_vertex(const DevGrid& g, int axis, double p, int* oob) {
  const int n = g.n[axis];
  const double c_first = g.lo[axis];
  const double c_last = vertex_coord(g, axis, n - 1);
  if (!(p >= c_first) || !(p <= c_last)) {  // also catches NaN
    *oob = 1;
    return 0;
  }
  if (n == 1) return 0;
  // i = searchsorted(c, p, 'right') - 1 clipped to [0, n-2], on the exact vertex values
  int i = (int)floor((p - g.lo[axis]) * g.inv_step[axis]);
  i = max(0, min(i, n - 2));
  while (i > 0 && vertex_coord(g, axis, i) > p) --i;
  while (i < n - 2 && vertex_coord(g, axis, i + 1) <= p) ++i;
  const double ci = vertex_coord(g, axis, i), cn = vertex_coord(g, axis, i + 1);
  const double yi = __ddiv_rn(__dsub_rn(p, ci), __dsub_rn(cn, ci));
  return yi <= 0.5 ? i : i + 1;
}

__global__ void test_simulate_kernel(DevGrid g, const double* __restrict__ B, int rows_B, int bs,
                                     const double* __restrict__ ne, int rows_ne, int n_real,
                                     int b_axis, const double* __restrict__ pts, int P,
                                     double scale, double* __restrict__ out, int* oob_flag) {
  const long long total = (long long)n_real * P;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(w / P), p = (int)(w % P);
    int oob = 0;
    const int ix = nearest_vertex(g, 0, pts[3 * p + 0], &oob);
    const int iy = nearest_vertex(g, 1, pts[3 * p + 1], &oob);
    const int iz = nearest_vertex(g, 2, pts[3 * p + 2], &oob);
    if (oob) {
      atomicOr(oob_flag, 1);
      out[w] = nan("");
      continue;
    }
    const long long c = ((long long)ix * g.n[1] + iy) * g.n[2] + iz;
    const double b = B[((long long)(rows_B > 1 ? r : 0) * g.ncell + c) * bs + b_axis];
    const double e = ne[(long long)(rows_ne > 1 ? r : 0) * g.ncell + c];
    out[w] = __dmul_rn(__dmul_rn(b, e), scale);
  }
}

__global__ void pol_derive_kernel(const double* __restrict__ Q, const double* __restrict__ U,
                                  long long n, double* __restrict__ PI, double* __restrict__ PA) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double q = Q[i], u = U[i];
    if (PI) PI[i] = sqrt(__dadd_rn(__dmul_rn(q, q), __dmul_rn(u, u)));
    if (PA) PA[i] = 0.5 * atan2(u, q);
  }
}

// ---- host side --------------------------------------------------------------------------
int make_dev_grid(const imb_grid* grid, DevGrid* g) {
  if (!grid) return IMB_EINVAL;
  g->ncell = 1;
  for (int a = 0; a < 3; ++a) {
    if (grid->n[a] < 1) return IMB_EINVAL;
    g->n[a] = grid->n[a];
    g->lo[a] = grid->lo[a];
    const double span = grid->hi[a] - grid->lo[a];
    g->step[a] = grid->n[a] != 1 ? span / (double)(grid->n[a] - 1) : 1.0;
    g->inv_step[a] = (grid->n[a] != 1 && span != 0.0) ? (double)(grid->n[a] - 1) / span : 0.0;
    g->ncell *= grid->n[a];
  }
  return IMB_OK;
}

static int field_type(int kind) {
  switch (kind) {
    case IMB_FIELD_CONST_B: case IMB_FIELD_GAUSS_B: case IMB_FIELD_LSA_B: case IMB_FIELD_JF12_B: return 0;
    case IMB_FIELD_CONST_NE: case IMB_FIELD_EXP_NE: case IMB_FIELD_COS_NE:
    case IMB_FIELD_RANDOM_NE: return 1;
    case IMB_FIELD_CRE_ANALYTIC: case IMB_FIELD_CRE_CONST: return 2;
    default: return -1;
  }
}
static bool is_stochastic(int kind) {
  return kind == IMB_FIELD_GAUSS_B || kind == IMB_FIELD_RANDOM_NE;
}

// (1/c_i, ln c_i) with c_i := 1 / fl(1 / (1 + i/128)) and (cos, sin)(pi j / 128), rounded from long double
static int upload_normal_tab() {
  static bool ready = false;
  if (ready) return IMB_OK;
  static double2 h[kTabSize];
  for (int i = 0; i <= 128; ++i) {
    const double inv_c = (double)(1.0L / (1.0L + (long double)i / 128.0L));
    h[kTabLog + i].x = inv_c;
    h[kTabLog + i].y = (double)(-logl((long double)inv_c));
  }
  h[kTabLog].y = 0.0;
  h[kTabLog + 128].y = kLn2;                    // cancels e ln 2 exactly for u next to 1
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int j = 0; j <= 256; ++j) {
    h[kTabTrig + j].x = (double)cosl(pi * (long double)j / 128.0L);
    h[kTabTrig + j].y = (double)sinl(pi * (long double)j / 128.0L);
  }
  for (int j = 0; j <= 256; j += 64) {          // exact values on the axes
    const int q = j / 64;
    h[kTabTrig + j].x = (q % 2) ? 0.0 : ((q % 4) ? -1.0 : 1.0);
    h[kTabTrig + j].y = (q % 2) ? ((q % 4 == 1) ? 1.0 : -1.0) : 0.0;
  }
  IMB_CUDA(cudaMemcpyToSymbol(g_normal_tab, h, sizeof(h)));
  ready = true;
  return IMB_OK;
}

int grid_blocks(long long work, int threads, int per_sm) {
  const long long need = (work + threads - 1) / threads;
  const long long cap = (long long)state().sm_count * per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace imb

using namespace imb;

extern "C" {

int imb_field_rows(const imb_field_desc* fields, int nfields, int type, int n_real) {
  int rows = 0;
  for (int f = 0; f < nfields; ++f) {
    if (field_type(fields[f].kind) != type) continue;
    if (is_stochastic(fields[f].kind)) return n_real;
    rows = 1;
  }
  return rows;
}

static int field_eval_impl(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                           const uint64_t* seeds, int n_seed_rows, int n_real, double* out_B, int rows_B,
                           int b_stride, int out_b_stride, int lane_offset, double* out_ne, int rows_ne,
                           double* out_cre, int rows_cre, void* stream_);

int imb_field_eval(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                   const uint64_t* seeds, int n_seed_rows, int n_real, double* out_B, int rows_B,
                   int b_stride, double* out_ne, int rows_ne, double* out_cre, int rows_cre,
                   void* stream_) {
  return field_eval_impl(grid, fields, nfields, seeds, n_seed_rows, n_real, out_B, rows_B, b_stride, b_stride, 0,
                         out_ne, rows_ne, out_cre, rows_cre, stream_);
}

int imb_field_eval_lanes(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                         const uint64_t* seeds, int n_seed_rows, int n_real, double* out_B, int b_stride,
                         int out_b_stride, int lane_offset, void* stream_) {
  if (!out_B || (b_stride != 24 && b_stride != 48 && b_stride != 96)) return IMB_EINVAL;
  if ((out_b_stride != 48 && out_b_stride != 96) || out_b_stride < b_stride) return IMB_EINVAL;
  const int rl = b_stride / 3, orl = out_b_stride / 3;
  // ONE group of the thread mapping, landing inside one group of the buffer
  if (n_real < 1 || n_real > rl || lane_offset < 0 || lane_offset % 8 != 0 || lane_offset + rl > orl)
    return IMB_EINVAL;
  return field_eval_impl(grid, fields, nfields, seeds, n_seed_rows, n_real, out_B, n_real, b_stride, out_b_stride,
                         lane_offset, nullptr, 0, nullptr, 0, stream_);
}

static int field_eval_impl(const imb_grid* grid, const imb_field_desc* fields, int nfields,
                           const uint64_t* seeds, int n_seed_rows, int n_real, double* out_B, int rows_B,
                           int b_stride, int out_b_stride, int lane_offset, double* out_ne, int rows_ne,
                           double* out_cre, int rows_cre, void* stream_) {
  IMB_TRY(check_init());
  IMB_TRY(upload_normal_tab());
  cudaStream_t stream = (cudaStream_t)stream_;
  DevGrid g;
  IMB_TRY(make_dev_grid(grid, &g));
  if (!fields || nfields < 0 || n_real < 1) return IMB_EINVAL;
  if (out_B && b_stride != 3 && b_stride != 4 && b_stride != 24 && b_stride != 48 && b_stride != 96)
    return IMB_EINVAL;
  for (int f = 0; f < nfields; ++f)
    if (field_type(fields[f].kind) < 0) return IMB_EINVAL;

  // seeds -> device scratch (pageable H2D is staged before the call returns)
  if (n_seed_rows < 0) return IMB_EINVAL;
  const size_t n_seeds = (size_t)(n_seed_rows > 0 ? n_seed_rows : 1) * (size_t)n_real;
  // four rotating slots: a later call (possibly on another stream) never overwrites the seeds a
  // kernel still in flight is reading
  const size_t slot_bytes = align_up(sizeof(unsigned long long) * n_seeds, 256);
  IMB_TRY(ensure_buffer(&state().seeds, &state().seeds_bytes, 4 * slot_bytes));
  unsigned long long* d_seeds = reinterpret_cast<unsigned long long*>(
      static_cast<char*>(state().seeds) +
      (state().seeds_slot++ % 4) * (state().seeds_bytes / 4 / 256 * 256));   // capacity-based stride
  bool any_stochastic = false;
  for (int f = 0; f < nfields; ++f) {
    if (!is_stochastic(fields[f].kind)) continue;
    any_stochastic = true;
    if (fields[f].seed_row < 0 || fields[f].seed_row >= n_seed_rows) return IMB_EINVAL;
  }
  if (any_stochastic) {
    if (!seeds) return IMB_EINVAL;
    IMB_CUDA(cudaMemcpyAsync(d_seeds, seeds, sizeof(unsigned long long) * n_seeds,
                             cudaMemcpyHostToDevice, stream));
  }

  double* outs[3] = {out_B, out_ne, out_cre};
  const int rows[3] = {rows_B, rows_ne, rows_cre};
  for (int type = 0; type < 3; ++type) {
    if (!outs[type]) continue;
    const int need = imb_field_rows(fields, nfields, type, n_real);
    if (need == 0) return IMB_EINVAL;               // output requested, no field of that type
    if (rows[type] != need) return IMB_EINVAL;
    FieldArgs a;
    a.g = g;
    a.rows = need;
    a.seeds = d_seeds;
    a.n_real = n_real;
    a.nfields = 0;
    for (int f = 0; f < nfields; ++f) {
      if (field_type(fields[f].kind) != type) continue;
      if (a.nfields == IMB_MAX_FIELDS) return IMB_EUNSUPPORTED;
      a.kind[a.nfields] = fields[f].kind;
      a.stream[a.nfields] = (unsigned int)fields[f].stream;
      a.seed_row[a.nfields] = fields[f].seed_row;
      for (int k = 0; k < 24; ++k) a.p[a.nfields][k] = fields[f].p[k];
      a.nfields++;
    }
    a.simple = a.has_const = a.gauss = 0;
    a.csum[0] = a.csum[1] = a.csum[2] = 0.0;
    a.out_rl = out_b_stride / 3;
    a.lane_off = lane_offset;
    if (type == 0) {
      int n_gauss = 0, n_const = 0;
      for (int f = 0; f < a.nfields; ++f) {
        if (a.kind[f] == IMB_FIELD_GAUSS_B) { n_gauss++; a.gauss = f; }
        if (a.kind[f] == IMB_FIELD_CONST_B) {
          for (int k = 0; k < 3; ++k) a.csum[k] = n_const ? a.csum[k] + a.p[f][k] : a.p[f][k];
          n_const++;
        }
      }
      a.has_const = n_const > 0;
      a.simple = n_gauss == 1 && n_gauss + n_const == a.nfields &&
                 (a.gauss == a.nfields - 1 || a.nfields == 2);
    }
    const long long work = ((g.ncell + 1) / 2) * need;
    const int blocks = grid_blocks(work, 256, 8);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(outs[type]);
    if (type == 0) {
      const long long row_elems = g.ncell * b_stride;
      if (b_stride >= 24) {
        // realisation-interleaved layouts (see field_b_il_kernel); only meaningful per realisation
        const int rl = b_stride / 3;
        if (need != n_real || addr % 32 != 0) return IMB_EINVAL;
        if ((g.ncell + 1) / 2 * rl >= (1ll << 32)) return IMB_EUNSUPPORTED;   // 32-bit work index
        const int n_groups = (need + rl - 1) / rl;
        int bx = grid_blocks(((g.ncell + 1) / 2) * rl, 256, 8) / n_groups;
        if (bx < 1) bx = 1;
        const dim3 grid(bx, n_groups);
        const bool wide = out_b_stride != b_stride || lane_offset != 0;
        if (wide) {
          if (rl == 32) return IMB_EINVAL;
          if (a.simple) {
            if (rl == 16) field_b_il_kernel<16, true, true><<<grid, 256, 0, stream>>>(a, outs[type]);
            else field_b_il_kernel<8, true, true><<<grid, 256, 0, stream>>>(a, outs[type]);
          } else {
            if (rl == 16) field_b_il_kernel<16, false, true><<<grid, 256, 0, stream>>>(a, outs[type]);
            else field_b_il_kernel<8, false, true><<<grid, 256, 0, stream>>>(a, outs[type]);
          }
        } else if (a.simple) {
          if (rl == 32) field_b_il_kernel<32, true><<<grid, 256, 0, stream>>>(a, outs[type]);
          else if (rl == 16) field_b_il_kernel<16, true><<<grid, 256, 0, stream>>>(a, outs[type]);
          else field_b_il_kernel<8, true><<<grid, 256, 0, stream>>>(a, outs[type]);
        } else {
          if (rl == 32) field_b_il_kernel<32, false><<<grid, 256, 0, stream>>>(a, outs[type]);
          else if (rl == 16) field_b_il_kernel<16, false><<<grid, 256, 0, stream>>>(a, outs[type]);
          else field_b_il_kernel<8, false><<<grid, 256, 0, stream>>>(a, outs[type]);
        }
      } else if (b_stride == 3) {
        const bool vec = (addr % 16 == 0) && (need == 1 || row_elems % 2 == 0);
        if (vec) field_b_kernel<3, true><<<blocks, 256, 0, stream>>>(a, outs[type]);
        else field_b_kernel<3, false><<<blocks, 256, 0, stream>>>(a, outs[type]);
      } else {
        const bool vec = (addr % 32 == 0);
        if (vec) field_b_kernel<4, true><<<blocks, 256, 0, stream>>>(a, outs[type]);
        else field_b_kernel<4, false><<<blocks, 256, 0, stream>>>(a, outs[type]);
      }
    } else {
      const bool vec = (addr % 16 == 0) && (need == 1 || g.ncell % 2 == 0);
      if (vec) field_scalar_kernel<true><<<blocks, 256, 0, stream>>>(a, outs[type]);
      else field_scalar_kernel<false><<<blocks, 256, 0, stream>>>(a, outs[type]);
    }
    IMB_LAUNCHED();
  }
  return IMB_OK;
}

int imb_healpix_ring_dirs(int nside, double* dirs, void* stream) {
  IMB_TRY(check_init());
  if (nside < 1 || nside > 8192 || !dirs) return IMB_EINVAL;
  const long long npix = 12ll * nside * nside;
  healpix_ring_dirs_kernel<<<grid_blocks(npix, 256, 8), 256, 0, (cudaStream_t)stream>>>(nside, dirs);
  IMB_LAUNCHED();
  return IMB_OK;
}

int imb_test_simulate(const imb_grid* grid, const double* B, int rows_B, int b_stride,
                      const double* ne, int rows_ne, int n_real, int b_axis,
                      const double* points, int n_points, double scale, double* out,
                      void* stream_) {
  IMB_TRY(check_init());
  cudaStream_t stream = (cudaStream_t)stream_;
  DevGrid g;
  IMB_TRY(make_dev_grid(grid, &g));
  if (!B || !ne || !points || !out || n_points < 0 || n_real < 1) return IMB_EINVAL;
  if (b_axis < 0 || b_axis > 2 || (b_stride != 3 && b_stride != 4)) return IMB_EINVAL;
  if ((rows_B != 1 && rows_B != n_real) || (rows_ne != 1 && rows_ne != n_real)) return IMB_EINVAL;
  if (n_points == 0) return IMB_OK;
  IMB_TRY(ensure_workspace(256));
  int* flag = static_cast<int*>(state().ws);
  IMB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
  const long long work = (long long)n_real * n_points;
  test_simulate_kernel<<<grid_blocks(work, 128, 16), 128, 0, stream>>>(
      g, B, rows_B, b_stride, ne, rows_ne, n_real, b_axis, points, n_points, scale, out, flag);
  IMB_LAUNCHED();
  int* h = static_cast<int*>(state().pinned);
  IMB_CUDA(cudaMemcpyAsync(h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
  IMB_CUDA(cudaStreamSynchronize(stream));
  return *h ? IMB_EBOUNDS : IMB_OK;
}

int imb_pol_derive(const double* Q, const double* U, int64_t count, double* out_PI,
                   double* out_PA, void* stream) {
  IMB_TRY(check_init());
  if (!Q || !U || count < 0) return IMB_EINVAL;
  if (count == 0) return IMB_OK;
  pol_derive_kernel<<<grid_blocks(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(Q, U, count,
                                                                               out_PI, out_PA);
  IMB_LAUNCHED();
  return IMB_OK;
}

}  // extern "C"
