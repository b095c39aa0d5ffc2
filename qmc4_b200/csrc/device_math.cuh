// device_math.cuh — scalar fp64 building blocks of the pair loop, Philox4x32-10 and warp helpers.
// Everything here is hand-scheduled for the sm_100a FP64 pipe (64 DFMA lanes/SM): each helper
// documents how many FP64-pipe instructions it issues, because that pipe is the roofline of the
// whole path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmc {

#define QMC_DEVICE __device__ __forceinline__
#ifdef QMC_EMULATE   /* tests/emu host build: libstdc++ uses the token __noinline__ itself */
#define QMC_NOINLINE __attribute__((noinline))
#else
#define QMC_NOINLINE __noinline__
#endif

// ---------------------------------------------------------------------------------------------
// Minimum image, exactly the reference's  t - std::round(t*Linv)*L  (dmc/geometry.h:50),
// std::round = half away from zero.  FP64 pipe: 2 (DFMA.RZ + DFMA); FRND.TRUNC runs on the
// conversion unit and the copysign is one LOP3.
//   trunc(RZ(|t|*Linv + 0.5)) == round(|t|*Linv)   [same construction CUDA's round() uses:
//   DADD.RZ + FRND.TRUNC], with the product folded into the add; round() is odd so the sign of t
//   is re-applied to the result.  It can differ from the
//   reference only when fl(t*Linv) is an exact tie but t*Linv is not (or vice versa); for
//   power-of-two boxes the product is exact and ties (lattice configurations) round identically.
QMC_DEVICE double min_image(double t, double L, double Linv)
{
  // work on |t| (free operand modifier on the FP64 pipe) and put the sign back with one LOP3:
  //   m = trunc(RZ(|t|*Linv + 0.5)) = round-half-away(|t|/L);   d = sign(t) * (|t| - m*L)
  const double a = fabs(t);
  const double m = trunc(__fma_rz(a, Linv, 0.5));
  const double d = fma(-m, L, a);
  const int hi = __double2hiint(d) ^ (__double2hiint(t) & 0x80000000);
  return __hiloint2double(hi, __double2loint(d));
}

// |minimum image| when only r^2 is needed (histogram pass): the sign fix-up is skipped
QMC_DEVICE double min_image_abs(double t, double L, double Linv)
{
  const double a = fabs(t);
  const double m = trunc(__fma_rz(a, Linv, 0.5));
  return fma(-m, L, a);
}

// ---------------------------------------------------------------------------------------------
// exp(a*g) for a*g <= 0 given the reduced-argument constants of ExpScaled (below).
// FP64 pipe: 14 (1 DFMA magic, 1 DADD, 2 DFMA Cody-Waite, 10 DFMA Horner); the 2^n scaling is
// integer arithmetic on the exponent field.  Degree-10 interpolant on |f| <= ln2/2: relative
// error 1.2e-15 (oracle tolerance is 1e-10).  Underflow (n < -1000) flushes to 0.
struct ExpScaled {
  double k;          // a*log2(e)
  double chi, clo;   // ln2/a split hi/lo so that n*chi is exact
  double c[11];      // c[j] = T_j * a^j, T = degree-10 fit of exp on [-ln2/2, ln2/2]
};

constexpr double EXP_T[11] = {1.0, 1.0, 0.5, 0.16666666666648303, 0.041666666666651364,
                              0.008333333353717156, 0.0013888888905871347, 0.0001984120875699232,
                              2.4801536409064087e-05, 2.7625102005388108e-06, 2.761379555451986e-07};

QMC_DEVICE double exp_scaled(double g, const ExpScaled &E)
{
  const double MAGIC = 6755399441055744.0;          // 2^52 + 2^51
  double t = fma(g, E.k, MAGIC);
  int n = __double2loint(t);
  double nd = t - MAGIC;
  double f = fma(nd, -E.chi, g);
  f = fma(nd, -E.clo, f);
  double p = E.c[10];
#pragma unroll
  for (int j = 9; j >= 1; --j) p = fma(p, f, E.c[j]);
  p = fma(p, f, 1.0);
  int hi = __double2hiint(p) + (n << 20);
  double r = __hiloint2double(hi, __double2loint(p));
  return (n < -1000) ? 0.0 : r;
}

// plain exp(x) for x <= 0 with the same polynomial (used by the generic functors' hot cases)
QMC_DEVICE double exp_neg(double x)
{
  const double MAGIC = 6755399441055744.0;
  double t = fma(x, 1.4426950408889634, MAGIC);
  int n = __double2loint(t);
  double nd = t - MAGIC;
  double f = fma(nd, -6.93147180369123816490e-01, x);
  f = fma(nd, -1.90821492927058770002e-10, f);
  double p = EXP_T[10];
#pragma unroll
  for (int j = 9; j >= 1; --j) p = fma(p, f, EXP_T[j]);
  p = fma(p, f, 1.0);
  int hi = __double2hiint(p) + (n << 20);
  double r = __hiloint2double(hi, __double2loint(p));
  return (n < -1000) ? 0.0 : r;
}

// 1/sqrt(x) to ~1 ulp for normal positive x without the library's special-case branch:
// MUFU.RSQ64H seed + one Halley-type step (5 FP64).  x must be > 0 and normal.
QMC_DEVICE double rsqrt_fast(double x)
{
  double y = __hiloint2double(0, 0);
#ifdef QMC_EMULATE   /* tests/emu: host build of the kernels, test infrastructure only */
  y = (double)(1.0f / sqrtf((float)x));
#else
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
  double y2 = y * y;
  double e = fma(-x, y2, 1.0);                 // 1 - x*y^2
  double h = fma(e, 0.375, 0.5);               // 1/2 + 3/8 e
  double ye = y * e;
  return fma(h, ye, y);                        // y*(1 + e/2 + 3/8 e^2)
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — counter-based, so every (walker slot, step, chunk)
// gets its stream with no state in HBM.
struct u32x4 { uint32_t x, y, z, w; };

QMC_DEVICE u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
{
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r)
    {
      uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
      uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
      u32x4 n;
      n.x = hi1 ^ c.y ^ k0;
      n.y = lo1;
      n.z = hi0 ^ c.w ^ k1;
      n.w = lo0;
      c = n;
      k0 += W0;
      k1 += W1;
    }
  return c;
}

// uniform in (0,1) with 53 random bits
QMC_DEVICE double u01_53(uint32_t a, uint32_t b)
{
  uint64_t m = ((uint64_t)a << 21) ^ (uint64_t)b;    // 53 bits
  m &= ((1ull << 53) - 1);
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

// Two N(0,1) deviates from two 32-bit words (Box-Muller).  The transcendental part runs on the
// MUFU unit (lg2, sin, cos, sqrt: ~2^-21 absolute error), the FP64 pipe only sees the two
// conversions: the deviates only seed the proposal R' (dmc/moves/vmcMoves.cpp:31); the Green
// function ratio is computed from the fp64 positions actually proposed, so detailed balance does not
// depend on the deviates' precision.  (x+0.5)*2^-32 in float is exact for small x, so the radial
// tail reaches sqrt(2*33*ln2) = 6.8 sigma.
QMC_DEVICE void normal_pair(uint32_t a, uint32_t b, double &z0, double &z1)
{
  float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  u1 = fminf(u1, 1.0f);
  const float r = sqrtf(-1.3862943611198906f * __log2f(u1));     /* -2 ln u = -2 ln2 * lg2 u */
  float s, c;
  __sincosf(6.283185307179586f * u2 - 3.141592653589793f, &s, &c);   /* angle in [-pi, pi): best MUFU accuracy */
  z0 = (double)(r * c);
  z1 = (double)(r * s);
}

// ---------------------------------------------------------------------------------------------
// exp(x) K_0(x) and exp(x) K_1(x), x > 0, for the dipolar_rep Jastrow (jastrowDipolar.h:31-62 takes them from
// boost::math::cyl_bessel_k).  K_nu(x) = int_0^inf exp(-x cosh t) cosh(nu t) dt by the trapezoidal rule: the integrand is
// analytic and decays doubly exponentially, so the rule converges geometrically (step min(0.2, 0.5/sqrt(x)), <= 70 nodes,
// 4e-15 relative against scipy.special.kve on 1e-4..700); cosh t - 1 = 2 sinh^2(t/2) avoids cancellation.  Cold path.
QMC_DEVICE void bessel_k01_scaled(double x, double &k0, double &k1)
{
  const double h = fmin(0.2, 0.5 / sqrt(x));
  double s0 = 0.5, s1 = 0.5;
  for (int n = 1; n < 4096; ++n)
    {
      const double t = n * h, sh = sinh(0.5 * t), arg = 2.0 * x * sh * sh;
      if (arg > 45.0) break;
      const double e = exp(-arg);
      s0 += e;
      s1 += e * cosh(t);
    }
  k0 = h * s0;
  k1 = h * s1;
}

// ---------------------------------------------------------------------------------------------
QMC_DEVICE double shfl_d(double v, int src)
{
  int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
  int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
  return __hiloint2double(hi, lo);
}
QMC_DEVICE double shfl_xor_d(double v, int m)
{
  int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), m);
  int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), m);
  return __hiloint2double(hi, lo);
}
QMC_DEVICE double warp_sum(double v)
{
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}
QMC_DEVICE int warp_sum_i(int v)
{
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// barrier over one walker group (a set of whole warps inside the CTA); id 1..15
QMC_DEVICE void group_sync(int bar_id, int nthreads)
{
#ifdef QMC_EMULATE
  emu::barrier(bar_id, nthreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
#endif
}

} // namespace qmc
