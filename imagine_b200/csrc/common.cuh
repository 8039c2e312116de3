// Shared device/host helpers for libimagine_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/imagine_b200.h"

namespace imb {

// ---- library state (api.cu) ---------------------------------------------------------
struct State {
  bool initialised = false;
  int device = -1;
  int sm_count = 148;
  void* ws = nullptr;          // scratch workspace (device)
  size_t ws_bytes = 0;
  // dedicated small buffers, so that field evaluation and LOS integration may run on different
  // streams without touching the likelihood scratch: seed slots (rotating) and packed (n_e, CRE)
  void* seeds = nullptr;
  size_t seeds_bytes = 0;
  unsigned seeds_slot = 0;
  void* aux = nullptr;
  size_t aux_bytes = 0;
  void* pinned = nullptr;      // small pinned host buffer for scalar read-back
  size_t pinned_bytes = 0;
  long long launches = 0;
  // every call that carves the scratch workspace bumps ws_epoch; the phased (sharded) Gram route records the
  // epoch of its phase 1 and refuses to continue if anything else used the workspace in between
  unsigned long long ws_epoch = 0, phase_epoch = 0;
  char cuda_err[256] = {0};
};
State& state();
int set_cuda_error(cudaError_t e, const char* where);
int ensure_workspace(size_t bytes);   // grows state().ws; returns IMB_* code
int ensure_buffer(void** buf, size_t* cap, size_t bytes);   // same for a dedicated buffer
int check_init();

#define IMB_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t _e = (call);                                               \
    if (_e != cudaSuccess) return imb::set_cuda_error(_e, #call);          \
  } while (0)

#define IMB_LAUNCHED()                                                     \
  do {                                                                     \
    imb::state().launches++;                                               \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return imb::set_cuda_error(_e, "kernel launch");\
  } while (0)

#define IMB_TRY(expr)                          \
  do {                                         \
    int _rc = (expr);                          \
    if (_rc != IMB_OK) return _rc;             \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve typed pointers out of the workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};

// ---- device helpers --------------------------------------------------------------------
#ifdef __CUDACC__

// grid geometry as the kernels see it
struct DevGrid {
  double lo[3];
  double step[3];      // (hi-lo)/(n-1), or 1.0 when n == 1   (numpy mgrid semantics)
  double inv_step[3];  // (n-1)/(hi-lo)
  int n[3];
  long long ncell;
};

__device__ __forceinline__ double vertex_coord(const DevGrid& g, int axis, int i) {
  // np.mgrid: index*step + start with two roundings (no fma) -> bit-identical coordinates
  return __dadd_rn(__dmul_rn((double)i, g.step[axis]), g.lo[axis]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Neumaier-compensated accumulator: keeps long reductions within ~1 ulp of exact
struct Comp {
  double s = 0.0, c = 0.0;
  __device__ __forceinline__ void add(double x) {
    double t = __dadd_rn(s, x);
    if (fabs(s) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(s, t), x));
    else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), s));
    s = t;
  }
  __device__ __forceinline__ double value() const { return s + c; }
};

// block-wide sum of a double; result valid in thread 0.  `smem` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (warp == 0) {
    r = lane < nw ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// ---- Philox4x32-10 (Salmon et al. 2011); oracle/fields.py:philox4x32_10 --------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo) {
  const unsigned long long k = ((unsigned long long)hi << 20) | (unsigned long long)(lo >> 12);
  return ((double)k + 0.5) * 2.220446049250313e-16;  // 2^-52
}

// ---- Box-Muller pair of block `blk` of stream `stream` under `seed` ------------------------
// z0 + i z1 = sqrt(-2 ln u1) exp(2 pi i u2) with u = (k + 1/2) 2^-52 from 52 random bits each
// (oracle/fields.py:philox_normal_blocks).  log and sin/cos are evaluated directly on the integer
// random bits with TABLE-DRIVEN range reduction (a 129-entry table of (1/c, ln c), c = 1 + i/128,
// and a 257-entry table of (cos, sin)(pi j/128), 6 KB, staged in shared memory by every block):
//   ln u   = e ln 2 + ln c_i + log1p(m/c_i - 1),  |m/c_i - 1| <= 2^-8   -> degree-7 series
//   angle  = pi j/128 + x, |x| <= pi/256                                  -> sin to x^5, cos to x^6,
//            then one rotation by the tabulated (cos, sin)
// |error| <= 4e-16 absolute on ln u (relative 2^-52 next to u = 1, where c = 2 and e = -1 cancel
// exactly), <= 3e-16 on sin/cos.  The field kernel is instruction-issue bound: this is ~45 instead
// of ~95 instructions for the two transcendentals of a pair (the first version used a division and
// degree-9 / degree-13 polynomials; the CUDA library's log + sincospi take ~2x that again).
__device__ __constant__ double kSinC[7] = {-0.16666666666666666, 0.008333333333333333,
    -0.0001984126984126984, 2.7557319223985893e-06, -2.505210838544172e-08,
    1.6059043836821613e-10, -7.647163731819816e-13};
__device__ __constant__ double kCosC[8] = {-0.5, 0.041666666666666664, -0.001388888888888889,
    2.48015873015873e-05, -2.755731922398589e-07, 2.08767569878681e-09,
    -1.1470745597729725e-11, 4.779477332387385e-14};

constexpr int kTabLog = 0;       // 129 entries (1/c_i, ln c_i)
constexpr int kTabTrig = 129;    // 257 entries (cos, sin)(pi j / 128)
constexpr int kTabSize = 386;
constexpr double kLn2 = 0.6931471805599453;

// ln(K * 2^-53) for an odd 53-bit integer K (so the argument is in (0, 1)).  Exponent and mantissa
// come from the exact conversion (double)K (one I2F instead of count-leading-zeros + 64-bit shifts).
// The coefficients of r^4 .. r^7 are cut to their upper 32 bits so that they encode as FP64
// immediates (no constant materialisation in the loop); with |r| <= 2^-8 that costs < 2e-19.
__device__ __forceinline__ double log_u53(unsigned long long K, const double2* __restrict__ tab) {
  const double dk = (double)K;                                 // exact, K < 2^53
  const unsigned int hi = (unsigned int)__double2hiint(dk);
  const unsigned int fh = hi & 0x000FFFFFu;
  const int i = ((int)(fh >> 12) + 1) >> 1;                    // round((m - 1) 128), 0..128
  const double m = __hiloint2double((int)(fh | 0x3FF00000u), __double2loint(dk));       // [1, 2)
  const double2 t = tab[kTabLog + i];
  const double r = fma(m, t.x, -1.0);                          // m / c_i - 1 (c_i := 1 / t.x exactly)
  double q = __hiloint2double(0x3FC24924, 0);                  // 1/7
  q = fma(q, r, __hiloint2double((int)0xBFC55555, 0));         // -1/6
  q = fma(q, r, __hiloint2double(0x3FC99999, 0));              // 1/5
  q = fma(q, r, -0.25);
  q = fma(q, r, 0.3333333333333333);
  q = fma(q, r, -0.5);
  const double lp = fma(r * r, q, r);                          // log1p(r)
  return fma((double)((int)(hi >> 20) - 1076), kLn2, t.y) + lp;   // (e, i) = (-1, 128): exactly 0 + lp
}

// (cos, sin)(pi * K * 2^-52) for an odd 53-bit integer K (angle in (0, 2 pi))
__device__ __forceinline__ void sincos_u53(unsigned long long K, const double2* __restrict__ tab,
                                           double& sn, double& cs) {
  const int j = (int)((K + (1ull << 44)) >> 45);               // round(t 128), t = K 2^-52 in (0, 2)
  const long long d = (long long)K - ((long long)j << 45);     // |d| <= 2^44, exact
  // (double)d by the 2^52 + 2^51 offset trick (two integer adds and one DADD)
  const double df = __longlong_as_double(0x4338000000000000ll + d) - 6755399441055744.0;
  const double x = df * 6.975736996017264e-16;                 // pi 2^-52
  const double2 cj = tab[kTabTrig + j];
  const double z = x * x;
  // 1/120 and -1/720 as 32-bit-immediate constants (|x| <= pi/256: < 2e-18)
  const double sl = fma(x * z, fma(z, __hiloint2double(0x3F811111, 0), -0.16666666666666666), x);
  const double cl = fma(z, fma(z, fma(z, __hiloint2double((int)0xBF56C16C, 0), 0.041666666666666664), -0.5), 1.0);
  cs = fma(cj.x, cl, -(cj.y * sl));
  sn = fma(cj.y, cl, cj.x * sl);
}

// sqrt(w) for a positive normal w (here w = -2 ln u in [2e-16, 74]): hardware reciprocal-square-root
// seed (2^-22), one coupled Goldschmidt step (2^-45), one residual correction; within 1 ulp, and no
// special-case branch (the compiler's IEEE sequence carries one around every call)
__device__ __forceinline__ double sqrt_pos(double w) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
  double g = w * y, h = 0.5 * y;
  const double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  return fma(fma(-g, g, w), h, g);
}

__device__ __forceinline__ void philox_normal_pair(unsigned long long seed, uint32_t stream,
                                                   unsigned long long blk, const double2* __restrict__ tab,
                                                   double& z0, double& z1) {
  const uint4 x = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const unsigned long long k1 = ((unsigned long long)x.x << 20) | (unsigned long long)(x.y >> 12);
  const unsigned long long k2 = ((unsigned long long)x.z << 20) | (unsigned long long)(x.w >> 12);
  const double r = sqrt_pos(-2.0 * log_u53(2 * k1 + 1, tab));  // u1 = (2 k1 + 1) 2^-53
  double s, c;
  sincos_u53(2 * k2 + 1, tab, s, c);                           // 2 u2 = (2 k2 + 1) 2^-52
  z0 = r * c;
  z1 = r * s;
}

#endif  // __CUDACC__
}  // namespace imb
