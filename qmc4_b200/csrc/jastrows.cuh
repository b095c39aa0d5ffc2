// jastrows.cuh — device versions of the reference's radial Jastrow functors u(r), u'(r), u''(r)
// (dmc/wavefunction/jastrows/*.h, CRTP base jastrow.h:30) and of its pair-potential functors
// (dmc/potentialFunctors.h, dmc/potential.cpp).  One switch replaces the template zoo of
// factory.h:50-60; the two functors of the BASELINE configs (gaussian, gaussiangaussian) also
// have r^2-only fast forms inside the pair loop (evaluate.cuh) that never take a sqrt.
//
// Parameter slots RadialFn::p[] per kind (filled by derive_radial() in capi.cu, which follows
// the reference constructors line by line):
//   GAUSSIAN        p0=alpha
//   GAUSSGAUSS      p0=alpha(=1/(2 r0^2)) p1=u0
//   SQUAREWELL      p0=R0 p1=Rm p2=halfLBox p3=K0 p4=aInverse p5=alpha p6=lBox p7=B p8=C p9=longDistanceConstant
//   HARDSPHEREGAUSS p0=Rm p1=a p2=d p3=alpha p4=logC
//   BSPLINE         p0=inverseStepSize p1=cutOff p2=valueRight p3=derivativeRight  (+ coefficients)
//   POSCHTELLER     p0=Rm p1=halfLBox p2=K0 p3=alpha p4=lBox p5=C p6=longDistanceConstant
//   HARDSPHERE      p0=a p1=cutOff p2=A p3=k
//   SQUAREWELL_BOUND p0=R0 p1=Rm p2=halfLBox p3=K0 p4=logA p5=expCoeff p6=alpha p7=lBox p8=B p9=C p10=longDistanceConstant
//   DELTA_PHONONS   p0=k p1=delta p2=beta p3=z p4=cut_off p5=k2(=pi/(2 cut_off))
//   DELTA_IN_TRAP   p0=a
//   DIPOLAR_REP     p0=D(=sqrt(input D)) p1=halfLBox(=cut_off) p2=matching_point p3=alpha p4=k(=pi/(2 cut_off)) p5=logC
//   DELTA_BOUND_STATE_PHONONS p0=xI p1=k p2=logA p3=beta p4=k2(=pi/(2 cut_off))
#pragma once
#include "plan.h"
#include "device_math.cuh"

namespace qmc {

// u, u', u'' at distance r.  `coef` = spline coefficient pool (global memory, read-only).
static __device__ QMC_NOINLINE void eval_radial(const RadialFn &J, const double *__restrict__ coef, double r,
                                         double &d0, double &d1, double &d2)
{
  const double *p = J.p;
  switch (J.kind)
    {
    case J_GAUSSIAN: /* jastrowGaussian.h:23-25 */
      d0 = -p[0] * r * r;
      d1 = -2. * p[0] * r;
      d2 = -2. * p[0];
      return;
    case J_GAUSSGAUSS: /* jastrowGaussGauss.h:19-21 */
      {
        double e = exp(-r * r * p[0]);
        d0 = p[1] * e;
        d1 = -2 * p[1] * p[0] * r * e;
        d2 = -2 * p[0] * p[1] * (1 - 2 * p[0] * r * r) * e;
        return;
      }
    case J_SQUAREWELL: /* jastrowSquareWell.h:20-62 */
      {
        const double R0 = p[0], Rm = p[1], half = p[2], K0 = p[3], aInv = p[4], alpha = p[5], lBox = p[6], B = p[7], C = p[8];
        if (r <= R0)
          {
            double s = sin(K0 * r);
            d0 = log(s / r);
            d1 = K0 / tan(K0 * r) - 1 / r;
            d2 = -K0 * K0 / (s * s) + 1 / (r * r);
          }
        else if (r <= Rm)
          {
            d0 = log(1 / r - aInv);
            d1 = 1 / (r * (-1 + r * aInv));
            d2 = (-2 * r * aInv + 1) / (r * r * (r * aInv - 1) * (r * aInv - 1));
          }
        else if (r <= half)
          {
            double e1 = exp(-alpha * r), e2 = exp(-alpha * (lBox - r));
            double tmp = e1 + e2, tmp1 = -e1 + e2, den = B + C * tmp;
            d0 = log(den);
            d1 = alpha * C * tmp1 / den;
            d2 = alpha * alpha * C * (tmp / den - C * tmp1 * tmp1 / (den * den));
          }
        else
          {
            d0 = p[9];
            d1 = 0;
            d2 = 0;
          }
        return;
      }
    case J_SQUAREWELL_BOUND: /* jastrowSquareWell.h:80-150 */
      {
        const double R0 = p[0], Rm = p[1], half = p[2], K0 = p[3], logA = p[4], ec = p[5], alpha = p[6], lBox = p[7], B = p[8], C = p[9];
        if (r <= R0)
          {
            double s = sin(K0 * r);
            d0 = log(s / r) + logA;
            d1 = K0 / tan(K0 * r) - 1 / r;
            d2 = -K0 * K0 / (s * s) + 1 / (r * r);
          }
        else if (r <= Rm)
          {
            d0 = -ec * r - log(r);
            d1 = -ec - 1. / r;
            d2 = 1. / (r * r);
          }
        else if (r <= half)
          {
            double e1 = exp(-alpha * r), e2 = exp(-alpha * (lBox - r));
            double tmp = e1 + e2, tmp1 = -e1 + e2, den = B + C * tmp;
            d0 = log(den);
            d1 = alpha * C * tmp1 / den;
            d2 = alpha * alpha * C * (tmp / den - C * tmp1 * tmp1 / (den * den));
          }
        else
          {
            d0 = p[10];
            d1 = 0;
            d2 = 0;
          }
        return;
      }
    case J_HARDSPHEREGAUSS: /* jastrowHardSphereGauss.h:13-23 */
      {
        const double Rm = p[0], a = p[1], d = p[2], alpha = p[3], logC = p[4];
        if (r <= Rm)
          {
            d0 = log(1 - a / r);
            d1 = a / (r * (r - a));
            double t = r * (r - a);
            d2 = -a / (t * t) * (2 * r - a);
          }
        else
          {
            d0 = logC - alpha * (r - d) * (r - d);
            d1 = -2 * alpha * (r - d);
            d2 = -2 * alpha;
          }
        return;
      }
    case J_HARDSPHERE: /* jastrowHardSphereGauss.h:39-60; d0 is 1 (sic) beyond the cut-off */
      {
        const double a = p[0], cut = p[1], A = p[2], k = p[3];
        if (r < cut)
          {
            double s = sin(k * (r - a));
            d0 = log(A * s) - log(r);
            d1 = k / tan(k * (r - a)) - 1 / r;
            d2 = -k * k * 1. / (s * s) + 1 / (r * r);
          }
        else
          {
            d0 = 1;
            d1 = 0;
            d2 = 0;
          }
        return;
      }
    case J_POSCHTELLER: /* jastrowPoschTeller.h:19-57 */
      {
        const double Rm = p[0], half = p[1], K0 = p[2], alpha = p[3], lBox = p[4], C = p[5];
        if (r <= Rm)
          {
            double sh = sinh(2 * K0 * r);
            d0 = log(tanh(K0 * r) / r);
            d1 = 2 * K0 / sh - 1 / r;
            d2 = -4 * K0 * K0 * cosh(2 * K0 * r) / (sh * sh) + 1. / (r * r);
          }
        else if (r <= half)
          {
            double e1 = exp(-alpha * r), e2 = exp(-alpha * (lBox - r));
            double tmp = e1 + e2, tmp1 = -e1 + e2;
            d0 = log(C * tmp);
            d1 = alpha * tmp1 / tmp;
            d2 = alpha * alpha * (1. - tmp1 * tmp1 / (tmp * tmp));
          }
        else
          {
            d0 = p[6];
            d1 = 0;
            d2 = 0;
          }
        return;
      }
    case J_DELTA_PHONONS: /* jastrow_delta.h:14-17,50-53 */
      {
        const double k = p[0], delta = p[1], beta = p[2], z = p[3], cut = p[4], k2 = p[5];
        if (r < z)
          {
            double s = sin(k * r + delta);
            d0 = log(s);
            d1 = k / tan(k * r + delta);
            d2 = -k * k / (s * s);
          }
        else if (r > cut)
          {
            d0 = log(sin(k * cut + delta));      /* sic: the inner branch evaluated at the cut-off */
            d1 = 0;
            d2 = 0;
          }
        else
          {
            double s = sin(k2 * r);
            d0 = log(pow(s, beta));
            d1 = k2 * beta / tan(k2 * r);
            d2 = -k2 * k2 * beta / (s * s);
          }
        return;
      }
    case J_DIPOLAR_REP: /* jastrowDipolar.h:21-62; K_2, K_3 by the forward recurrence, K_{-1} = K_1 */
      {
        const double D = p[0], half = p[1], mp = p[2], alpha = p[3], k = p[4], logC = p[5];
        if (r < mp)
          {
            const double a = sqrt(r), y = 2 * D / a;
            double K0, K1;
            bessel_k01_scaled(y, K0, K1);
            const double ex = exp(-y);               /* the reference works with the unscaled functions: same underflow */
            K0 *= ex;
            K1 *= ex;
            const double K2 = K0 + (2. / y) * K1, K3 = K1 + (4. / y) * K2;
            d0 = log(a * K1) + logC;
            const double Jp = -(K2 + K0) / 2. / K1;
            d1 = 0.5 / r * (1 - Jp * y);
            const double Jp2 = (K3 + K1 + 2 * K1) / 4. / K1;
            d2 = (-1 + y * Jp * 3 / 2. + 0.5 * y * y * (Jp2 - Jp * Jp)) / (2 * r * r);
          }
        else if (r <= half)
          {
            const double s = sin(k * r);
            d0 = alpha * log(s);
            d1 = alpha * k / tan(k * r);
            d2 = -alpha * k * k / (s * s);
          }
        else d0 = d1 = d2 = 0;
        return;
      }
    case J_DELTA_IN_TRAP: /* jastrow_delta.h:64-66 */
      d0 = log(r + p[0]);
      d1 = 1. / (r + p[0]);
      d2 = -(1. / ((r + p[0]) * (r + p[0])));
      return;
    case J_DELTA_BOUND_STATE_PHONONS: /* jastrow_delta.h:81-83 */
      {
        const double xI = p[0], k = p[1], logA = p[2], beta = p[3], k2 = p[4];
        if (r < xI)
          {
            d0 = -k * r;
            d1 = -k;
            d2 = 0;
          }
        else
          {
            double s = sin(k2 * r);
            d0 = logA + log(pow(s, beta));
            d1 = beta * k2 / tan(k2 * r);
            d2 = -k2 * beta * k2 / (s * s);
          }
        return;
      }
    case J_BSPLINE: /* jastrowSpline.cpp:80-110: cubic B-spline, linear extrapolation beyond cutOff */
      {
        const double inv = p[0], cut = p[1];
        if (r < cut)
          {
            double x = r * inv;
            int i = (int)floor(x);
            x -= i;
            const double *c = coef + J.coef_off + i;
            double c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
            double x2 = x * x, x3 = x2 * x;
            /* A*X, Ad1*X, Ad2*X with X=(1,x,x^2,x^3) (jastrowSpline.cpp:25-76) */
            double a0 = 1. / 6 - 0.5 * x + 0.5 * x2 - 1. / 6 * x3;
            double a1 = 2. / 3 - x2 + 0.5 * x3;
            double a2 = 1. / 6 + 0.5 * x + 0.5 * x2 - 0.5 * x3;
            double a3 = 1. / 6 * x3;
            double b0 = -0.5 + x - 0.5 * x2;
            double b1 = -2 * x + 1.5 * x2;
            double b2 = 0.5 + x - 1.5 * x2;
            double b3 = 0.5 * x2;
            double e0 = 1 - x, e1 = -2 + 3 * x, e2 = 1 - 3 * x, e3 = x;
            d0 = c0 * a0 + c1 * a1 + c2 * a2 + c3 * a3;
            d1 = (c0 * b0 + c1 * b1 + c2 * b2 + c3 * b3) * inv;
            d2 = (c0 * e0 + c1 * e1 + c2 * e2 + c3 * e3) * inv * inv;
          }
        else
          {
            d0 = p[2] + (r - cut) * p[3];
            d1 = p[3];
            d2 = 0;
          }
        return;
      }
    default:
      d0 = d1 = d2 = 0;
      return;
    }
}

// pair potential at distance r for the kinds that need r itself (generic path)
static __device__ QMC_NOINLINE double eval_pair_potential(int vkind, double v_val, double r)
{
  switch (vkind)
    {
    case V_DIPOLAR: /* potential.cpp:72-98 with k=0 images */
      return v_val / (r * r * r);
    case V_POSCHTELLER: /* potential.cpp:100-114 */
      {
        double c = cosh(r * v_val);
        return -2 * v_val * v_val / (c * c);
      }
    default:
      return 0.;
    }
}

} // namespace qmc
