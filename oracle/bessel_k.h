/* oracle/bessel_k.h — TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Modified Bessel functions of the second kind K_n(x), integer order, for the `dipolar_rep` Jastrow
 * (dmc/wavefunction/jastrows/jastrowDipolar.h:31-62), which the reference takes from Boost.Math
 * (boost::math::cyl_bessel_k; Boost is not vendored under /root/reference, version unpinned:
 * find_package(Boost REQUIRED), cmake/init.cmake:11) and this image does not have.
 *
 * Algorithm (published, restated here): K_nu(x) = int_0^inf exp(-x cosh t) cosh(nu t) dt evaluated for nu = 0, 1 with
 * the trapezoidal rule, which converges geometrically for this analytic, doubly-exponentially decaying integrand
 * (step min(0.2, 0.5/sqrt(x)), at most 70 nodes); cosh t - 1 = 2 sinh^2(t/2) keeps the exponentially scaled form
 * exp(x) K_nu(x) free of cancellation.  Higher orders by the forward recurrence K_{n+1} = K_{n-1} + (2n/x) K_n, which is
 * what Boost's bessel_kn does for integer orders; K_{-n} = K_n.
 * PINNED against scipy.special.kv/kve (AMOS) to 4e-15 relative on 1e-4 <= x <= 700 (tests/test_oracle.py) — not against
 * Boost itself: for this one Jastrow, parity with the reference's binary is "unpinned". */
#ifndef QMC4_ORACLE_BESSEL_K_H
#define QMC4_ORACLE_BESSEL_K_H
#include <math.h>

static inline void orc_bessel_k01_scaled(double x, double *k0, double *k1)
{ /* exp(x) K_0(x), exp(x) K_1(x), x > 0 */
  double h = 0.5 / sqrt(x);
  if (h > 0.2) h = 0.2;
  double s0 = 0.5, s1 = 0.5;
  for (int n = 1; n < 4096; n++)
    {
      const double t = n * h, sh = sinh(0.5 * t), arg = 2 * x * sh * sh;
      if (arg > 45.) break;
      const double e = exp(-arg);
      s0 += e;
      s1 += e * cosh(t);
    }
  *k0 = h * s0;
  *k1 = h * s1;
}

static inline double orc_cyl_bessel_k(int n, double x)
{
  double k0, k1;
  if (n < 0) n = -n;
  orc_bessel_k01_scaled(x, &k0, &k1);
  const double ex = exp(-x);
  k0 *= ex;
  k1 *= ex;
  if (n == 0) return k0;
  for (int m = 1; m < n; m++)
    {
      const double k2 = k0 + (2. * m / x) * k1;
      k0 = k1;
      k1 = k2;
    }
  return k1;
}
#endif
