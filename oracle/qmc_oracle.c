/* qmc_oracle.c — CPU restatement of the reference's DMC walker-propagation path, plain C11.
 *
 * TEST INFRASTRUCTURE ONLY (see qmc_oracle.h).  Every function follows the reference routine it
 * cites (paths under /root/reference), in the reference's loop order and with the reference's
 * formulas (u, u', u'' evaluated separately, distance table built first, division by the norm in
 * the gradient, ...), so that it agrees with oracle/_ref to rounding.  Compiled with
 * -ffp-contract=off: the reference's Release build targets baseline x86-64, which has no FMA.
 * Status: pinned — tests/test_oracle.py checks it against oracle/_ref (the reference itself) and
 * against the golden vectors in tests/golden/ produced by the reference.
 */
#include "qmc_oracle.h"
#include "bessel_k.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(const char *m)
{
  snprintf(g_err, sizeof g_err, "%s", m);
  return 1;
}

/* ---- geometryPBC (dmc/geometry.h:50): t - round(t/L)*L, std::round = half away from zero ---- */
static double pbc_difference(const orc_model *m, double t, int d) { return t - round(t * (1. / m->lBox[d])) * m->lBox[d]; }

/* ---- radial functors (dmc/wavefunction/jastrows/) ----------------------------------------------- */
typedef struct {
  int kind;
  const double *p;
  int np;
  /* derived */
  double a[16];
} radial;

static int radial_init(radial *J, int kind, const double *p, int np)
{
  J->kind = kind;
  J->p = p;
  J->np = np;
  memset(J->a, 0, sizeof J->a);
  switch (kind)
    {
    case ORC_J_GAUSSIAN: J->a[0] = p[0]; return 0;
    case ORC_J_GAUSSGAUSS: /* jastrowGaussGauss.cpp:5-10 */
      J->a[0] = 1. / (2 * p[0] * p[0]);
      J->a[1] = p[1];
      return 0;
    case ORC_J_SQUAREWELL: /* jastrowSquareWell.cpp:5-44 */
      {
        double V0 = p[0], R0 = p[1], Rm = p[2], alpha = p[3], aInv = p[4], lBox = p[5] * 2;
        double K0 = sqrt(V0), half = lBox / 2.;
        double A = sin(K0 * R0) / (1 - R0 * aInv);
        double C = -A / (Rm * Rm * alpha * (exp(-alpha * (lBox - Rm)) - exp(-alpha * Rm)));
        double B = A * (1 / Rm - aInv) - C * (exp(-alpha * Rm) + exp(-alpha * (lBox - Rm)));
        if (B + C * 2 * exp(-alpha * half) < 0) return fail("Jastrow cannot be negative");
        J->a[0] = R0; J->a[1] = Rm; J->a[2] = half; J->a[3] = K0; J->a[4] = aInv; J->a[5] = alpha; J->a[6] = lBox;
        J->a[7] = B; J->a[8] = C; J->a[9] = log(B + C * 2 * exp(-alpha * half));
        return 0;
      }
    case ORC_J_SQUAREWELL_BOUND: /* jastrowSquareWell.cpp:46-95 */
      {
        double V0 = p[0], R0 = p[1], Rm = p[2], lBox = p[3] * 2., eb = p[5], alpha = p[6];
        double mm = 0.5, K0 = sqrt(2 * mm * (V0 - eb)), half = lBox / 2., ec = sqrt(2 * mm * eb);
        double A = exp(-ec * R0) / sin(K0 * R0);
        double C = exp(-ec * Rm) * (1. / (Rm * Rm) + ec / Rm) / (alpha * (exp(-alpha * Rm) - exp(-alpha * (lBox - Rm))));
        double B = -C * (exp(-alpha * Rm) + exp(-alpha * (lBox - Rm))) + exp(-ec * Rm) / Rm;
        if (B + C * 2 * exp(-alpha * half) < 0) return fail("Jastrow cannot be negative");
        J->a[0] = R0; J->a[1] = Rm; J->a[2] = half; J->a[3] = K0; J->a[4] = log(A); J->a[5] = ec; J->a[6] = alpha;
        J->a[7] = lBox; J->a[8] = B; J->a[9] = C; J->a[10] = log(B + C * 2 * exp(-alpha * half));
        return 0;
      }
    case ORC_J_HARDSPHEREGAUSS: /* jastrowHardSphereGauss.cpp:4-10 */
      {
        double a = p[0], Rm = p[1], d = p[2];
        double alpha = a / (Rm * (Rm - a) * 2 * (d - Rm));
        double C = (1 - a / Rm) * exp(alpha * (Rm - d) * (Rm - d));
        J->a[0] = Rm; J->a[1] = a; J->a[2] = d; J->a[3] = alpha; J->a[4] = log(C);
        return 0;
      }
    case ORC_J_HARDSPHERE: /* jastrowHardSphereGauss.cpp:18-26 */
      J->a[0] = p[0]; J->a[1] = p[2]; J->a[2] = p[2] / sin(p[1] * (p[2] - p[0])); J->a[3] = p[1];
      return 0;
    case ORC_J_POSCHTELLER: /* jastrowPoschTeller.cpp:4-17 */
      {
        double R0 = p[0], C = p[1], alpha = p[2], Rm = p[3], lBox = p[4] * 2.;
        J->a[0] = Rm; J->a[1] = lBox / 2.; J->a[2] = 1 / R0; J->a[3] = alpha; J->a[4] = lBox; J->a[5] = C;
        J->a[6] = 2 * C * exp(-alpha * (lBox / 2.));
        return 0;
      }
    case ORC_J_DELTA_PHONONS: /* jastrow_delta.cpp:5-39; p = {k,delta,beta,z,cut_off} */
      for (int i = 0; i < 5; i++) J->a[i] = p[i];
      J->a[5] = M_PI / (2 * p[4]);
      return 0;
    case ORC_J_DELTA_IN_TRAP: J->a[0] = p[0]; return 0;      /* jastrow_delta.cpp:71-74 */
    case ORC_J_DELTA_BOUND_STATE_PHONONS: /* jastrow_delta.cpp:76-86; p = {xI,k,A,beta,cut_off} */
      J->a[0] = p[0]; J->a[1] = p[1]; J->a[2] = log(p[2]); J->a[3] = p[3]; J->a[4] = M_PI / (p[4] * 2);
      return 0;
    case ORC_J_DIPOLAR_REP: /* jastrowDipolar.cpp:3-27; p = {D,cut_off,matching_point,alpha,C} as in the input */
      J->a[0] = sqrt(p[0]);            /* D = sqrt(j["D"]) */
      J->a[1] = p[1] * 2;              /* lBox = cut_off*2 */
      J->a[2] = p[2];                  /* matching_point */
      J->a[3] = p[3];                  /* alpha */
      J->a[4] = M_PI / J->a[1];        /* k = pi/lBox */
      J->a[5] = log(p[4]);             /* logC */
      return 0;
    case ORC_J_BSPLINE: /* jastrowSpline.cpp:20-78; p = {stepSize,valueRight,derivativeRight,c...} */
      J->a[0] = 1. / p[0];
      J->a[1] = (np - 3 - 3) * p[0];
      J->a[2] = p[1];
      J->a[3] = p[2];
      return 0;
    }
  return fail("unknown radial functor");
}

/* jastrow<T>::evaluateDerivatives (jastrow.h:30): d0, d1, d2 evaluated one after the other */
static void radial_eval(const radial *J, double x, double *d0, double *d1, double *d2)
{
  const double *a = J->a;
  switch (J->kind)
    {
    case ORC_J_GAUSSIAN: /* jastrowGaussian.h:23-25 */
      *d0 = -a[0] * x * x;
      *d1 = -2. * a[0] * x;
      *d2 = -2 * a[0];
      return;
    case ORC_J_GAUSSGAUSS: /* jastrowGaussGauss.h:19-21 */
      *d0 = a[1] * exp(-x * x * a[0]);
      *d1 = -2 * a[1] * a[0] * x * exp(-x * x * a[0]);
      *d2 = -2 * a[0] * a[1] * (1 - 2 * a[0] * x * x) * exp(-x * x * a[0]);
      return;
    case ORC_J_SQUAREWELL: /* jastrowSquareWell.h:20-62 */
      {
        double R0 = a[0], Rm = a[1], half = a[2], K0 = a[3], aInv = a[4], alpha = a[5], lBox = a[6], B = a[7], C = a[8];
        if (x <= R0)
          {
            *d0 = log(sin(K0 * x) / x);
            *d1 = (K0 / tan(K0 * x) - 1 / x);
            *d2 = (-K0 * K0 / pow(sin(K0 * x), 2) + 1 / (x * x));
          }
        else if (x <= Rm)
          {
            *d0 = log(1 / x - aInv);
            *d1 = 1 / (x * (-1 + x * aInv));
            *d2 = (-2 * x * aInv + 1) / (x * x * (x * aInv - 1) * (x * aInv - 1));
          }
        else if (x <= half)
          {
            double tmp = (exp(-alpha * x) + exp(-alpha * (lBox - x)));
            double tmp1 = (-exp(-alpha * x) + exp(-alpha * (lBox - x)));
            *d0 = log(B + C * (exp(-alpha * x) + exp(-alpha * (lBox - x))));
            *d1 = alpha * C * (-exp(-alpha * x) + exp(-alpha * (lBox - x))) / (B + C * (exp(-alpha * x) + exp(-alpha * (lBox - x))));
            *d2 = alpha * alpha * C * ((tmp) / (B + C * tmp) - C * tmp1 * tmp1 / pow(B + C * tmp, 2));
          }
        else
          {
            *d0 = a[9];
            *d1 = 0;
            *d2 = 0;
          }
        return;
      }
    case ORC_J_SQUAREWELL_BOUND: /* jastrowSquareWell.h:80-150 */
      {
        double R0 = a[0], Rm = a[1], half = a[2], K0 = a[3], logA = a[4], ec = a[5], alpha = a[6], lBox = a[7], B = a[8], C = a[9];
        if (x <= R0)
          {
            *d0 = log(sin(K0 * x) / x) + logA;
            *d1 = (K0 / tan(K0 * x) - 1 / x);
            *d2 = (-K0 * K0 / pow(sin(K0 * x), 2) + 1 / (x * x));
          }
        else if (x <= Rm)
          {
            *d0 = -ec * x - log(x);
            *d1 = -ec - 1. / x;
            *d2 = 1. / (x * x);
          }
        else if (x <= half)
          {
            double tmp = (exp(-alpha * x) + exp(-alpha * (lBox - x)));
            double tmp1 = (-exp(-alpha * x) + exp(-alpha * (lBox - x)));
            *d0 = log(B + C * tmp);
            *d1 = alpha * C * tmp1 / (B + C * tmp);
            *d2 = alpha * alpha * C * ((tmp) / (B + C * tmp) - C * tmp1 * tmp1 / pow(B + C * tmp, 2));
          }
        else
          {
            *d0 = a[10];
            *d1 = 0;
            *d2 = 0;
          }
        return;
      }
    case ORC_J_HARDSPHEREGAUSS: /* jastrowHardSphereGauss.h:13-23 */
      {
        double Rm = a[0], aa = a[1], d = a[2], alpha = a[3], logC = a[4];
        if (x <= Rm)
          {
            *d0 = log(1 - aa / x);
            *d1 = aa / (x * (x - aa));
            *d2 = -aa / pow(x * (x - aa), 2) * (2 * x - aa);
          }
        else
          {
            *d0 = logC - alpha * (x - d) * (x - d);
            *d1 = -2 * alpha * (x - d);
            *d2 = -2 * alpha;
          }
        return;
      }
    case ORC_J_HARDSPHERE: /* jastrowHardSphereGauss.h:39-60 (d0 = 1 beyond the cut-off, sic) */
      {
        double aa = a[0], cut = a[1], A = a[2], k = a[3];
        if (x < cut)
          {
            *d0 = log(A * sin(k * (x - aa))) - log(x);
            *d1 = k / tan(k * (x - aa)) - 1 / x;
            *d2 = -k * k * 1. / pow(sin(k * (x - aa)), 2) + 1 / (x * x);
          }
        else
          {
            *d0 = 1;
            *d1 = 0;
            *d2 = 0;
          }
        return;
      }
    case ORC_J_POSCHTELLER: /* jastrowPoschTeller.h:19-57 */
      {
        double Rm = a[0], half = a[1], K0 = a[2], alpha = a[3], lBox = a[4], C = a[5];
        if (x <= Rm)
          {
            *d0 = log(tanh(K0 * x) / x);
            *d1 = (2 * K0 / sinh(2 * K0 * x) - 1 / (x));
            *d2 = (-4 * K0 * K0 * cosh(2 * K0 * x) / (sinh(2 * K0 * x) * sinh(2 * K0 * x)) + 1. / (x * x));
          }
        else if (x <= half)
          {
            double tmp = (exp(-alpha * x) + exp(-alpha * (lBox - x)));
            double tmp1 = (-exp(-alpha * x) + exp(-alpha * (lBox - x)));
            *d0 = log(C * tmp);
            *d1 = alpha * tmp1 / tmp;
            *d2 = alpha * alpha * (1. - tmp1 * tmp1 / pow(tmp, 2));
          }
        else
          {
            *d0 = a[6];
            *d1 = 0;
            *d2 = 0;
          }
        return;
      }
    case ORC_J_DELTA_PHONONS: /* jastrow_delta.h:14-17 (d0 = log(d0Exp), :50-53) */
      {
        double k = a[0], delta = a[1], beta = a[2], z = a[3], cut = a[4], k2 = a[5];
        double e = (x < z) ? sin(k * x + delta) : x > cut ? sin(k * cut + delta) : pow(sin(k2 * x), beta);
        *d0 = log(e);
        *d1 = (x < z) ? k / tan(k * x + delta) : x > cut ? 0 : k2 * beta / tan(k2 * x);
        *d2 = (x < z) ? -k * k / pow(sin(k * x + delta), 2) : x > cut ? 0 : -k2 * k2 * beta / pow(sin(k2 * x), 2);
        return;
      }
    case ORC_J_DELTA_IN_TRAP: /* jastrow_delta.h:64-66 */
      *d0 = log(x + a[0]);
      *d1 = 1. / (x + a[0]);
      *d2 = -(1. / pow(x + a[0], 2));
      return;
    case ORC_J_DELTA_BOUND_STATE_PHONONS: /* jastrow_delta.h:81-83 */
      {
        double xI = a[0], k = a[1], logA = a[2], beta = a[3], k2 = a[4];
        *d0 = (x < xI) ? -k * x : logA + log(pow(sin(k2 * x), beta));
        *d1 = (x < xI) ? -k : beta * k2 / tan(k2 * x);
        *d2 = (x < xI) ? 0 : -k2 * beta * k2 / pow(sin(k2 * x), 2);
        return;
      }
    case ORC_J_DIPOLAR_REP: /* jastrowDipolar.h:21-62 */
      {
        const double D = a[0], lBox = a[1], mp = a[2], alpha = a[3], k = a[4], logC = a[5];
        if (x < mp)
          {
            double sq = sqrt(x), y = 2 * D / sq;
            double J1 = orc_cyl_bessel_k(1, y), K0 = orc_cyl_bessel_k(0, y), K2 = orc_cyl_bessel_k(2, y);
            double K3 = orc_cyl_bessel_k(3, y), Km1 = orc_cyl_bessel_k(-1, y);
            *d0 = log(sq * J1) + logC;
            double Jp = -(K2 + K0) / 2.;
            Jp /= J1;
            *d1 = 0.5 / x * (1 - Jp * y);
            double Jp2 = (K3 + Km1 + 2 * J1) / 4.;
            Jp2 = Jp2 / J1;
            *d2 = (-1 + y * Jp * 3 / 2. + 0.5 * y * y * (Jp2 - Jp * Jp)) / (2 * x * x);
          }
        else if (x <= lBox / 2.)
          {
            *d0 = alpha * log(sin(k * x));
            *d1 = alpha * k * 1. / tan(k * x);
            *d2 = -alpha * k * k / pow(sin(k * x), 2);
          }
        else *d0 = *d1 = *d2 = 0;
        return;
      }
    case ORC_J_BSPLINE: /* jastrowSpline.cpp:80-110 */
      {
        double inv = a[0], cut = a[1];
        const double *c = J->p + 3;
        if (x < cut)
          {
            x = x * inv;
            int i = (int)floor(x);
            x -= i;
            double X[4] = {1, x, x * x, x * x * x};
            /* A = M^T (jastrowSpline.cpp:33-46), Ad1, Ad2 as written (:47-66) */
            static const double M[4][4] = {{1 / 6., 2. / 3, 1. / 6, 0}, {-0.5, 0, 0.5, 0}, {0.5, -1, 0.5, 0}, {-1. / 6, 0.5, -0.5, 1. / 6}};
            static const double A1[4][4] = {{-0.5, 1, -0.5, 0}, {0, -2, 3. / 2, 0}, {0.5, 1, -3. / 2, 0}, {0, 0, 0.5, 0}};
            static const double A2[4][4] = {{1, -1, 0, 0}, {-2, 3., 0, 0}, {1, -3, 0, 0}, {0, 1, 0, 0}};
            double AX[4], AX1[4], AX2[4];
            for (int r = 0; r < 4; r++)
              {
                AX[r] = AX1[r] = AX2[r] = 0;
                for (int cc = 0; cc < 4; cc++)
                  {
                    AX[r] += M[cc][r] * X[cc];
                    AX1[r] += A1[r][cc] * X[cc];
                    AX2[r] += A2[r][cc] * X[cc];
                  }
              }
            *d0 = c[i] * AX[0] + c[i + 1] * AX[1] + c[i + 2] * AX[2] + c[i + 3] * AX[3];
            *d1 = (c[i] * AX1[0] + c[i + 1] * AX1[1] + c[i + 2] * AX1[2] + c[i + 3] * AX1[3]) * inv;
            *d2 = (c[i] * AX2[0] + c[i + 1] * AX2[1] + c[i + 2] * AX2[2] + c[i + 3] * AX2[3]) * inv * inv;
          }
        else
          {
            *d0 = a[2] + (x - cut) * a[3];
            *d1 = a[3];
            *d2 = 0;
          }
        return;
      }
    }
  *d0 = *d1 = *d2 = 0;
}

/* jastrow<T>::print (dmc/wavefunction/jastrows/jastrow.h:42-63): rows {x, d0, d1, d2} at the bin centres */
int orc_tabulate(int kind, const double *p, int np, double minx, double maxx, int n, double *out)
{
  radial J;
  if (radial_init(&J, kind, p, np)) return 1;
  double deltax = (maxx - minx) / n;
  for (int i = 0; i < n; i++)
    {
      double x = minx + (i + 0.5) * deltax;
      out[4 * i] = x;
      radial_eval(&J, x, out + 4 * i + 1, out + 4 * i + 2, out + 4 * i + 3);
    }
  return 0;
}

/* ---- walker with its distance tables (dmc/walkers.h:20-78, dmc/tableDistances.cpp:81-108) ------- */
typedef struct {
  int a, b;      /* b = -1: one-body table of set a */
  int n;
  double *diff;  /* column-major n x 3: diff[d*n + k] */
  double *norm;
} table;

typedef struct {
  const orc_model *m;
  const double *x[8];   /* positions per set: x[s][d*N+i] */
  int off[8];
  int dofs;
  table tab[16];
  int ntab;
} walker_t;

static void walker_init(walker_t *w, const orc_model *m, const double *R)
{
  w->m = m;
  int off = 0;
  for (int s = 0; s < m->nSpecies; s++)
    {
      w->x[s] = R + off;
      w->off[s] = off;
      off += 3 * m->N[s];
    }
  w->dofs = off;
  w->ntab = 0;
}
static void walker_free(walker_t *w)
{
  for (int t = 0; t < w->ntab; t++)
    {
      free(w->tab[t].diff);
      free(w->tab[t].norm);
    }
  w->ntab = 0;
}

static void table_norms(table *T)
{ /* norm() (geometry.cpp:27-44) */
  for (int i = 0; i < T->n; i++)
    {
      double tmp = 0;
      for (int d = 0; d < 3; d++) tmp += T->diff[d * T->n + i] * T->diff[d * T->n + i];
      T->norm[i] = sqrt(tmp);
    }
}

static const table *get_table(walker_t *w, int a, int b)
{
  for (int t = 0; t < w->ntab; t++)
    if (w->tab[t].a == a && w->tab[t].b == b) return &w->tab[t];
  const orc_model *m = w->m;
  table *T = &w->tab[w->ntab++];
  T->a = a;
  T->b = b;
  if (b < 0)
    { /* differencesOneBody with x = {0,0,0} (geometry.cpp:9-25, tableDistances.cpp:86) */
      int N = m->N[a];
      T->n = N;
      T->diff = malloc(sizeof(double) * 3 * (N > 0 ? N : 1));
      T->norm = malloc(sizeof(double) * (N > 0 ? N : 1));
      for (int i = 0; i < N; i++)
        for (int d = 0; d < 3; d++) T->diff[d * N + i] = pbc_difference(m, w->x[a][d * N + i] - 0., d);
    }
  else if (a == b)
    { /* differencesTwoBody(data) (geometry.cpp:47-65): k = i(i-1)/2 + j, r_i - r_j, i > j */
      int N = m->N[a], P = N * (N - 1) / 2;
      T->n = P;
      T->diff = malloc(sizeof(double) * 3 * (P > 0 ? P : 1));
      T->norm = malloc(sizeof(double) * (P > 0 ? P : 1));
      int k = 0;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < i; j++)
          {
            for (int d = 0; d < 3; d++) T->diff[d * P + k] = pbc_difference(m, w->x[a][d * N + i] - w->x[a][d * N + j], d);
            k++;
          }
    }
  else
    { /* differencesTwoBody(data1,data2) (geometry.cpp:100-122): k = i*N2 + j */
      int N1 = m->N[a], N2 = m->N[b], P = N1 * N2;
      T->n = P;
      T->diff = malloc(sizeof(double) * 3 * (P > 0 ? P : 1));
      T->norm = malloc(sizeof(double) * (P > 0 ? P : 1));
      int k = 0;
      for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++)
          {
            for (int d = 0; d < 3; d++) T->diff[d * P + k] = pbc_difference(m, w->x[a][d * N1 + i] - w->x[b][d * N2 + j], d);
            k++;
          }
    }
  table_norms(T);
  return T;
}

/* LOOP3B (dmc/tools3B.h:4-22): k outer, j<k, i<j ; ji=j(j-1)/2+i, kj=k(k-1)/2+j, ki=k(k-1)/2+i */
#define ORC_LOOP3B(N, BODY)                         \
  {                                                 \
    int kj = 0, ki2 = 0, ki = 0;                    \
    for (int k = 0; k < (N); k++)                   \
      {                                             \
        int ji = 0;                                 \
        for (int j = 0; j < k; j++, kj++)           \
          {                                         \
            ki = ki2;                               \
            for (int i = 0; i < j; i++, ji++, ki++) \
              {                                     \
                BODY                                \
              }                                     \
          }                                         \
        ki2 += k;                                   \
      }                                             \
  }

/* LOOP3B_DIS (dmc/tools3B.h:25-43): i<N1, j<N2, k<N3 ; ij=i*N2+j, jk=j*N3+k, ik=i*N3+k */
#define ORC_LOOP3B_DIS(N1, N2, N3, BODY)            \
  {                                                 \
    int ij = 0, ik2 = 0, ik = 0;                    \
    for (int i = 0; i < (N1); i++)                  \
      {                                             \
        int jk = 0;                                 \
        for (int j = 0; j < (N2); j++, ij++)        \
          {                                         \
            ik = ik2;                               \
            for (int k = 0; k < (N3); k++, jk++, ik++) \
              {                                     \
                BODY                                \
              }                                     \
          }                                         \
        ik2 += (N3);                                \
      }                                             \
  }

/* ---- evaluation: productWavefunction::evaluateDerivatives + sumPotentials + energy ------------- */
typedef struct {
  double logpsi, lap, V, E;
  int valid;
} eval_out;

static int evaluate(const orc_model *m, const double *R, double *grad, eval_out *out)
{
  walker_t w;
  walker_init(&w, m, R);
  for (int k = 0; k < w.dofs; k++) grad[k] = 0;
  double waveValue = 0, laplacian = 0, V = 0;
  int valid = 1;
  const int D = 3;
  /* satisfyConstraints first (walkers.cpp:51): hardSphereConstraintThreeBodyUnDis (jastrowWavefunctionThreeBody.cpp:15-36) */
  for (int t = 0; t < m->nTerms && valid; t++)
    {
      const int *ti = m->ti + 8 * t;
      if (ti[0] != 1 || ti[1] != ORC_WF_JASTROW3B_UNDIS) continue;
      const double *p = m->tp + ti[7];
      double hs = p[ti[6] - 1];
      if (hs < 0) continue;
      int A = ti[2], N = m->N[A];
      const table *T = get_table(&w, A, A);
      ORC_LOOP3B(N, double rji = T->norm[ji]; double rkj = T->norm[kj]; double rki = T->norm[ki];
                 double Rr = sqrt(rji * rji + rki * rki + rkj * rkj); if (Rr <= hs) valid = 0;)
    }
  /* hardSphereConstraintThreeBodyDis (jastrowWavefunctionThreeBody.cpp:54-80) */
  for (int t = 0; t < m->nTerms && valid; t++)
    {
      const int *ti = m->ti + 8 * t;
      if (ti[0] != 1 || ti[1] != ORC_WF_JASTROW3B_DIS) continue;
      const double *p = m->tp + ti[7];
      double hs = p[ti[6] - 1];
      if (hs < 0) continue;
      int A = ti[2], B = ti[3], C = ti[4];
      const table *TAB = get_table(&w, A, B), *TBC = get_table(&w, B, C), *TAC = get_table(&w, A, C);
      ORC_LOOP3B_DIS(m->N[A], m->N[B], m->N[C], double rij = TAB->norm[ij]; double rjk = TBC->norm[jk]; double rik = TAC->norm[ik];
                     double Rr = sqrt(rij * rij + rjk * rjk + rik * rik); if (Rr <= hs) valid = 0;)
    }
  if (valid)
    {
      /* potentials (potential.cpp:40-48) are summed BEFORE the kinetic part (energy.cpp:30-35) */
      for (int t = 0; t < m->nTerms; t++)
        {
          const int *ti = m->ti + 8 * t;
          if (ti[0] != 2) continue;
          const double *p = m->tp + ti[7];
          double sum = 0;
          switch (ti[1])
            {
            case ORC_POT_HARMONIC1B: /* potential.cpp:21-32 */
              {
                const table *T = get_table(&w, ti[2], -1);
                for (int i = 0; i < T->n; i++) sum += 0.5 * p[0] * p[0] * T->norm[i] * T->norm[i];
                break;
              }
            case ORC_POT_SQUAREWELL2B: /* potential.cpp:57-65: params {V0,R0} */
              {
                const table *T = get_table(&w, ti[2], ti[3]);
                long cnt = 0;
                for (int i = 0; i < T->n; i++) cnt += T->norm[i] <= p[1];
                sum = -cnt * p[0];
                break;
              }
            case ORC_POT_DIPOLAR2B: /* potential.cpp:72-98 with k=0 */
              {
                const table *T = get_table(&w, ti[2], ti[3]);
                for (int i = 0; i < T->n; i++)
                  {
                    double xx = T->norm[i];
                    sum += p[0] / (xx * xx * xx);
                  }
                break;
              }
            case ORC_POT_POSCHTELLER2B: /* potential.cpp:100-114, alpha = 1/R0 */
              {
                const table *T = get_table(&w, ti[2], ti[3]);
                double alpha = 1 / p[0];
                for (int i = 0; i < T->n; i++)
                  {
                    double tmp = cosh(T->norm[i] * alpha);
                    sum -= 2 * alpha * alpha / (tmp * tmp);
                  }
                break;
              }
            case ORC_POT_TWOBODY_SQUAREWELL: /* potentialTwoBody.h:23-33 + potentialFunctors.h:9 ; params {R0,V0} */
            case ORC_POT_TWOBODY_BARRIER:    /* potentialFunctors.h:26 */
              {
                const table *T = get_table(&w, ti[2], ti[3]);
                double v = ti[1] == ORC_POT_TWOBODY_SQUAREWELL ? -p[1] : p[1];
                for (int i = 0; i < T->n; i++) sum += T->norm[i] <= p[0] ? v : 0;
                break;
              }
            case ORC_POT_THREEBODY_UNDIS_SQUAREWELL: /* potentialThreeBody.h:23-44 */
            case ORC_POT_THREEBODY_UNDIS_BARRIER:
              {
                int A = ti[2], N = m->N[A];
                const table *T = get_table(&w, A, A);
                double v = ti[1] == ORC_POT_THREEBODY_UNDIS_SQUAREWELL ? -p[1] : p[1];
                ORC_LOOP3B(N, double rji = T->norm[ji]; double rkj = T->norm[kj]; double rki = T->norm[ki];
                           double Rr = sqrt(rji * rji + rki * rki + rkj * rkj); sum += Rr <= p[0] ? v : 0;)
                break;
              }
            case ORC_POT_THREEBODY_DIS_SQUAREWELL: /* potentialThreeBody.h:82-108 */
            case ORC_POT_THREEBODY_DIS_BARRIER:
              {
                int A = ti[2], B = ti[3], C = ti[4];
                const table *TAB = get_table(&w, A, B), *TBC = get_table(&w, B, C), *TAC = get_table(&w, A, C);
                double v = ti[1] == ORC_POT_THREEBODY_DIS_SQUAREWELL ? -p[1] : p[1];
                ORC_LOOP3B_DIS(m->N[A], m->N[B], m->N[C], double rij = TAB->norm[ij]; double rjk = TBC->norm[jk];
                               double rik = TAC->norm[ik]; double Rr = sqrt(rij * rij + rjk * rjk + rik * rik);
                               sum += Rr <= p[0] ? v : 0;)
                break;
              }
            default:
              walker_free(&w);
              return fail("unknown potential kind");
            }
          V += sum;
        }
      /* wavefunction terms in order (productWavefunction.cpp:34-37) */
      for (int t = 0; t < m->nTerms; t++)
        {
          const int *ti = m->ti + 8 * t;
          if (ti[0] != 1) continue;
          radial J;
          int np = ti[6] - ((ti[1] == ORC_WF_JASTROW3B_UNDIS || ti[1] == ORC_WF_JASTROW3B_DIS) ? 1 : 0);
          if (radial_init(&J, ti[5], m->tp + ti[7], np))
            {
              walker_free(&w);
              return 1;
            }
          double tmp, tmp1, tmp2;
          if (ti[1] == ORC_WF_JASTROW1B)
            { /* jastrowWavefunctionOneBody.h:53-85 */
              int A = ti[2], N = m->N[A];
              const table *T = get_table(&w, A, -1);
              double *g = grad + w.off[A];
              for (int i = 0; i < N; i++)
                {
                  double d = T->norm[i];
                  radial_eval(&J, d, &tmp, &tmp1, &tmp2);
                  laplacian += tmp2 + (D - 1) * tmp1 / d;
                  waveValue += tmp;
                  for (int id = 0; id < D; id++) g[id * N + i] += T->diff[id * T->n + i] / d * tmp1;
                }
            }
          else if (ti[1] == ORC_WF_JASTROW2B_UNDIS)
            { /* jastrowWavefunctionTwoBody.h:124-164 */
              int A = ti[2], N = m->N[A];
              const table *T = get_table(&w, A, A);
              double *g = grad + w.off[A];
              int k = 0;
              for (int i = 0; i < N; i++)
                for (int j = 0; j < i; j++)
                  {
                    double d = T->norm[k];
                    radial_eval(&J, d, &tmp, &tmp1, &tmp2);
                    laplacian += 2 * (tmp2 + (D - 1) * tmp1 / d);
                    waveValue += tmp;
                    for (int id = 0; id < D; id++)
                      {
                        g[id * N + i] += T->diff[id * T->n + k] / d * tmp1;
                        g[id * N + j] -= T->diff[id * T->n + k] / d * tmp1;
                      }
                    k++;
                  }
            }
          else if (ti[1] == ORC_WF_JASTROW2B_DIS)
            { /* jastrowWavefunctionTwoBody.h:33-76 */
              int A = ti[2], B = ti[3], NA = m->N[A], NB = m->N[B];
              const table *T = get_table(&w, A, B);
              double *gA = grad + w.off[A], *gB = grad + w.off[B];
              int k = 0;
              for (int i = 0; i < NA; i++)
                for (int j = 0; j < NB; j++)
                  {
                    double d = T->norm[k];
                    radial_eval(&J, d, &tmp, &tmp1, &tmp2);
                    laplacian += 2 * (tmp2 + (D - 1) * tmp1 / d);
                    waveValue += tmp;
                    for (int id = 0; id < D; id++)
                      {
                        gA[id * NA + i] += T->diff[id * T->n + k] / d * tmp1;
                        gB[id * NB + j] -= T->diff[id * T->n + k] / d * tmp1;
                      }
                    k++;
                  }
            }
          else if (ti[1] == ORC_WF_JASTROW3B_UNDIS)
            { /* jastrowWavefunctionThreeBody.h:91-137 */
              int A = ti[2], N = m->N[A];
              const table *T = get_table(&w, A, A);
              double *g = grad + w.off[A];
              int P = T->n;
              ORC_LOOP3B(N, double rji = T->norm[ji]; double rkj = T->norm[kj]; double rki = T->norm[ki];
                         double Rr = sqrt(rji * rji + rki * rki + rkj * rkj); radial_eval(&J, Rr, &tmp, &tmp1, &tmp2);
                         for (int id = 0; id < D; id++) {
                           g[id * N + i] -= tmp1 * (T->diff[id * P + ji] + T->diff[id * P + ki]) / Rr;
                           g[id * N + j] += tmp1 * (T->diff[id * P + ji] - T->diff[id * P + kj]) / Rr;
                           g[id * N + k] += tmp1 * (T->diff[id * P + ki] + T->diff[id * P + kj]) / Rr;
                         } laplacian += tmp2 * 3;
                         laplacian += 3 * tmp1 / Rr * (2 * D - 1); waveValue += tmp;)
            }
          else if (ti[1] == ORC_WF_JASTROW3B_DIS)
            { /* jastrowWavefunctionThreeBody.h:222-281 */
              int A = ti[2], B = ti[3], C = ti[4], NA = m->N[A], NB = m->N[B], NC = m->N[C];
              const table *TAB = get_table(&w, A, B), *TBC = get_table(&w, B, C), *TAC = get_table(&w, A, C);
              double *gA = grad + w.off[A], *gB = grad + w.off[B], *gC = grad + w.off[C];
              ORC_LOOP3B_DIS(NA, NB, NC, double rij = TAB->norm[ij]; double rjk = TBC->norm[jk]; double rik = TAC->norm[ik];
                             double Rr = sqrt(rij * rij + rjk * rjk + rik * rik); radial_eval(&J, Rr, &tmp, &tmp1, &tmp2);
                             for (int id = 0; id < D; id++) {
                               gA[id * NA + i] += tmp1 * (TAB->diff[id * TAB->n + ij] + TAC->diff[id * TAC->n + ik]) / Rr;
                               gB[id * NB + j] += tmp1 * (-TAB->diff[id * TAB->n + ij] + TBC->diff[id * TBC->n + jk]) / Rr;
                               gC[id * NC + k] -= tmp1 * (TAC->diff[id * TAC->n + ik] + TBC->diff[id * TBC->n + jk]) / Rr;
                             } laplacian += tmp2 * 3;
                             laplacian += 3 * tmp1 / Rr * (2 * D - 1); waveValue += tmp;)
            }
          else
            {
              walker_free(&w);
              return fail("unknown wavefunction kind");
            }
        }
    }
  double ef = 0;
  for (int k = 0; k < w.dofs; k++) ef += grad[k] * grad[k];      /* energy.cpp:15-19 */
  out->logpsi = waveValue;
  out->lap = laplacian;
  out->V = V;
  out->E = V + (-0.5 * (ef + laplacian));                         /* energy.cpp:27,33-34 */
  out->valid = valid;
  walker_free(&w);
  return 0;
}

int orc_eval_model(const orc_model *m, const double *R, int nW, double *logpsi, double *grad, double *lap, double *V,
                   double *E, int *valid)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  double *g = malloc(sizeof(double) * dofs);
  for (int w = 0; w < nW; w++)
    {
      eval_out o;
      if (evaluate(m, R + (size_t)w * dofs, g, &o))
        {
          free(g);
          return 1;
        }
      if (valid) valid[w] = o.valid;
      if (logpsi) logpsi[w] = o.valid ? o.logpsi : 0;
      if (lap) lap[w] = o.valid ? o.lap : 0;
      if (V) V[w] = o.valid ? o.V : 0;
      if (E) E[w] = o.valid ? o.E : 0;
      if (grad)
        for (int k = 0; k < dofs; k++) grad[(size_t)w * dofs + k] = o.valid ? g[k] : 0;
    }
  free(g);
  return 0;
}

/* driftDiffusionFirstOrder::transitionProbabilityRatio (dmcMoves.cpp:32-57) */
static double green_ratio(int dofs, double timeStep, const double *Rn, const double *Fn, const double *Ro, const double *Fo)
{
  double q = 0;
  for (int k = 0; k < dofs; k++)
    {
      double a = (timeStep * 1. / 2 * (Fo[k] - Fn[k]) - Rn[k] + Ro[k]);
      double b = (Fo[k] + Fn[k]);
      q += a * b;
    }
  return q;
}

int orc_green_ratio_model(const orc_model *m, double timeStep, int nW, const double *Rnew, const double *Fnew,
                          const double *Rold, const double *Fold, double *q)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  for (int w = 0; w < nW; w++)
    q[w] = green_ratio(dofs, timeStep, Rnew + (size_t)w * dofs, Fnew + (size_t)w * dofs, Rold + (size_t)w * dofs, Fold + (size_t)w * dofs);
  return 0;
}

/* driftDiffusionFirstOrder::move (dmcMoves.cpp:12-30) + gaussianMover::move (vmcMoves.cpp:11-35) with the
 * deviates supplied in draw order (set-major, particle-major, dimension-minor) */
static void drift_diffuse(const orc_model *m, double timeStep, const double *Ro, const double *Fo, const double *xi, double *Rn)
{
  int off = 0, kd = 0;
  for (int s = 0; s < m->nSpecies; s++)
    {
      int N = m->N[s];
      for (int k = 0; k < 3 * N; k++) Rn[off + k] = Ro[off + k] + Fo[off + k] * timeStep;
      for (int i = 0; i < N; i++)
        for (int d = 0; d < 3; d++) Rn[off + d * N + i] = Rn[off + d * N + i] + xi[kd++] * sqrt(timeStep);
      off += 3 * N;
    }
}

int orc_drift_diffuse_model(const orc_model *m, double timeStep, int nW, const double *Rold, const double *Fold,
                            const double *xi, double *Rnew)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  for (int w = 0; w < nW; w++)
    drift_diffuse(m, timeStep, Rold + (size_t)w * dofs, Fold + (size_t)w * dofs, xi + (size_t)w * dofs, Rnew + (size_t)w * dofs);
  return 0;
}

/* dmcDriver::update (dmcDriver.cpp:56-84) with metropolis::acceptLog (metropolis.cpp:5-29) */
int orc_update_model(const orc_model *m, double timeStep, int nW, const double *Rold, const double *xi, const double *u_m,
                     double *Rout, double *Fout, double *logpsi_out, double *E_out, int *accepted)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  double *Fo = malloc(sizeof(double) * dofs), *Fn = malloc(sizeof(double) * dofs), *Rn = malloc(sizeof(double) * dofs);
  int rc = 0;
  for (int w = 0; w < nW && !rc; w++)
    {
      const double *Ro = Rold + (size_t)w * dofs;
      eval_out eo, en;
      rc = evaluate(m, Ro, Fo, &eo);
      if (rc) break;
      drift_diffuse(m, timeStep, Ro, Fo, xi + (size_t)w * dofs, Rn);
      rc = evaluate(m, Rn, Fn, &en);
      if (rc) break;
      int acc = 0;
      if (en.valid)
        {
          double q = green_ratio(dofs, timeStep, Rn, Fn, Ro, Fo);
          double ratioLog = 2 * (en.logpsi - eo.logpsi) + q;
          acc = (ratioLog > 0.) ? 1 : (ratioLog > log(u_m[w]));
        }
      accepted[w] = acc;
      for (int k = 0; k < dofs; k++)
        {
          Rout[(size_t)w * dofs + k] = acc ? Rn[k] : Ro[k];
          Fout[(size_t)w * dofs + k] = acc ? Fn[k] : Fo[k];
        }
      logpsi_out[w] = acc ? en.logpsi : eo.logpsi;
      E_out[w] = acc ? en.E : eo.E;
    }
  free(Fo);
  free(Fn);
  free(Rn);
  return rc;
}

/* One vmcDriver::step (dmc/vmcDriver.cpp:55-84) per chain given the random numbers it consumes: gaussianMover
 * x' = x + xi*sqrt(sigma) in draw order (vmcMoves.cpp:22-33), update (:38-52: constraints, log psi), acceptLog on
 * 2*(log psi' - log psi) (:68), invalid = rejected (:73-77), reject restores the old walker (:79-82).  E_out is the
 * `energy` observable the driver accumulates on the current walker (energy.cpp:28-35). */
int orc_vmc_step_model(const orc_model *m, double sigma, int nW, const double *Rold, const double *xi, const double *u_m,
                       double *Rout, double *logpsi_out, double *E_out, int *accepted)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  double *Fo = malloc(sizeof(double) * dofs), *Fn = malloc(sizeof(double) * dofs), *Rn = malloc(sizeof(double) * dofs);
  int rc = 0;
  for (int w = 0; w < nW && !rc; w++)
    {
      const double *Ro = Rold + (size_t)w * dofs, *x = xi + (size_t)w * dofs;
      eval_out eo, en;
      rc = evaluate(m, Ro, Fo, &eo);
      if (rc) break;
      int off = 0, kk = 0;
      for (int s = 0; s < m->nSpecies; s++)
        {
          int N = m->N[s];
          for (int i = 0; i < N; i++)
            for (int d = 0; d < 3; d++) Rn[off + d * N + i] = Ro[off + d * N + i] + x[kk++] * sqrt(sigma);
          off += 3 * N;
        }
      rc = evaluate(m, Rn, Fn, &en);
      if (rc) break;
      int acc = 0;
      if (en.valid)
        {
          double ratioLog = 2 * (en.logpsi - eo.logpsi);
          acc = (ratioLog > 0.) ? 1 : (ratioLog > log(u_m[w]));
        }
      accepted[w] = acc;
      for (int k = 0; k < dofs; k++) Rout[(size_t)w * dofs + k] = acc ? Rn[k] : Ro[k];
      logpsi_out[w] = acc ? en.logpsi : eo.logpsi;
      E_out[w] = acc ? en.E : eo.E;
    }
  free(Fo);
  free(Fn);
  free(Rn);
  return rc;
}

/* ---- estimators --------------------------------------------------------------------------------- */
static void hist_accumulate(double *sums, double minx, double dxinv, double a, double x)
{ /* histogramAccumulator::accumulate (accumulators.h:117).  The reference has no lower bound: for x < minx - dx it
   * writes before the array (undefined behaviour); such samples are dropped here and by the CUDA kernels.  minx - dx < x <
   * minx truncates to bin 0 exactly as the reference's int conversion does. */
  int i = (int)((x - minx) * dxinv);
  if (i < 0) return;
  sums[i] += a;
}

int orc_histogram_model(const orc_model *m, const double *R, int nW, int iEst, double *sums, int bins, double *weight)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  const int *ti = NULL;
  int seen = 0;
  for (int t = 0; t < m->nTerms; t++)
    if (m->ti[8 * t] == 3)
      {
        if (seen == iEst) ti = m->ti + 8 * t;
        seen++;
      }
  if (!ti) return fail("estimator index out of range");
  if (ti[5] != bins) return fail("bins mismatch");
  const double *p = m->tp + ti[7];     /* {minx,maxx} or {maxq,lBox,channel} */
  for (int i = 0; i < bins; i++) sums[i] = 0;
  double wsum = 0;
  double *g = malloc(sizeof(double) * dofs);
  for (int iw = 0; iw < nW; iw++)
    {
      const double *Rw = R + (size_t)iw * dofs;
      eval_out o;
      if (evaluate(m, Rw, g, &o))
        {
          free(g);
          return 1;
        }
      if (!o.valid) continue;
      walker_t w;
      walker_init(&w, m, Rw);
      if (ti[1] == ORC_EST_PAIRCORRELATION)
        { /* pairCorrelation.cpp:17-59 */
          double minx = p[0], maxx = p[1], dx = (maxx - minx) / bins, dxinv = 1. / dx;
          int A = ti[2], B = ti[3];
          double nf = pow(m->lBox[0], 3) / (dx * 4 * M_PI * m->N[A] * m->N[B]);
          nf = (A == B) ? 2 * nf : nf;
          const table *T = get_table(&w, A, B);
          for (int i = 0; i < T->n; i++)
            if (T->norm[i] < maxx) hist_accumulate(sums, minx, dxinv, nf / (T->norm[i] * T->norm[i]), T->norm[i]);
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_TRIMERCORRELATION_UNDIS)
        { /* trimerCorrelation.cpp:82-143 */
          double minx = p[0], maxx = p[1], dx = (maxx - minx) / bins, dxinv = 1. / dx;
          int A = ti[2], N = m->N[A];
          double NA = N;
          double nf = 1 / (dx * 4 * M_PI * NA * NA * NA);
          const table *T = get_table(&w, A, A);
          ORC_LOOP3B(N, double rji = T->norm[ji]; double rkj = T->norm[kj]; double rki = T->norm[ki];
                     double radius = sqrt(rji * rji + rki * rki + rkj * rkj);
                     if (radius < maxx) hist_accumulate(sums, minx, dxinv, nf / (radius * radius), radius);)
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_TRIMERCORRELATION)
        { /* trimerCorrelation.cpp:9-78: set C holds ONE particle; NC = getN(setB) in the normalisation (sic, :64) */
          double minx = p[0], maxx = p[1], dx = (maxx - minx) / bins, dxinv = 1. / dx;
          int A = ti[2], B = ti[3], C = ti[4];
          if (m->N[C] != 1)
            {
              walker_free(&w);
              free(g);
              return fail("trimerCorrelation: set C must hold one particle");
            }
          double NA = m->N[A], NB = m->N[B], NC = m->N[B];
          double nf = 1 / (dx * 4 * M_PI * NA * NB * NC);
          const table *T1 = get_table(&w, A, C), *T2 = get_table(&w, B, C), *T3 = get_table(&w, A, B);
          int N1 = T1->n, N2 = T2->n;
          for (int i = 0; i < N1; i++)
            for (int j = 0; j < N2; j++)
              {
                double d3 = T3->norm[i * N2 + j];
                double radius = sqrt(T1->norm[i] * T1->norm[i] + T2->norm[j] * T2->norm[j] + d3 * d3);
                if (radius < maxx) hist_accumulate(sums, minx, dxinv, nf / (radius * radius), radius);
              }
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_TRIMERCORRELATION_DIS)
        { /* trimerCorrelation.cpp:163-227 */
          double minx = p[0], maxx = p[1], dx = (maxx - minx) / bins, dxinv = 1. / dx;
          int A = ti[2], B = ti[3], C = ti[4];
          double NA = m->N[A], NB = m->N[B], NC = m->N[C];
          double nf = 1 / (dx * 4 * M_PI * NA * NB * NC);
          const table *TAB = get_table(&w, A, B), *TBC = get_table(&w, B, C), *TAC = get_table(&w, A, C);
          ORC_LOOP3B_DIS(m->N[A], m->N[B], m->N[C], double rij = TAB->norm[ij]; double rjk = TBC->norm[jk]; double rik = TAC->norm[ik];
                         double radius = sqrt(rij * rij + rik * rik + rjk * rjk);
                         if (radius < maxx) hist_accumulate(sums, minx, dxinv, nf / (radius * radius), radius);)
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_TRIMERHYPERRADIUS)
        { /* hyperRadius.cpp:17-43 (scalar: sums[0] += value, realScalarEstimator / scalarAccumulator accumulators.h:27) */
          int A = ti[2], B = ti[3], C = ti[4];
          if (m->N[C] != 1)
            {
              walker_free(&w);
              free(g);
              return fail("trimerhyperRadius: set C must hold one particle");
            }
          const table *T1 = get_table(&w, A, C), *T2 = get_table(&w, B, C), *T3 = get_table(&w, A, B);
          int N1 = T1->n, N2 = T2->n;
          double radius = 0;
          for (int i = 0; i < N1; i++)
            for (int j = 0; j < N2; j++)
              {
                double d3 = T3->norm[i * N2 + j];
                radius += sqrt(T1->norm[i] * T1->norm[i] + T2->norm[j] * T2->norm[j] + d3 * d3);
              }
          sums[0] += radius / (N1 * N2);
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_CENTEROFMASSSQUARED)
        { /* centerOfMassSquared.cpp:8-18: dis = norm(differencesOneBody(positions,0)); cm = dis.sum()^2 / dis.size() */
          const table *T = get_table(&w, ti[2], -1);
          double tot = 0;
          for (int i = 0; i < T->n; i++) tot += T->norm[i];
          sums[0] += tot * tot / T->n;
          wsum += 1;
        }
      else if (ti[1] == ORC_EST_STRUCTUREFACTOR)
        { /* structureFactor.cpp:8-40,56-90 applied to coordinate 0 (the reference builds it for
             DIMENSIONS==1 only; SURVEY.md §8(c) defines the 3-D estimator as the same formula on x) */
          double qMax = p[0], lb = p[1];
          int channel = (int)p[2];
          double deltaQ = (int)(qMax / (2 * M_PI * bins / lb)) * 2 * M_PI / lb;
          if (deltaQ == 0) deltaQ = 2 * M_PI / lb;
          int A = ti[2], B = ti[3], NA = m->N[A], NB = m->N[B];
          for (int j = 0; j < bins; j++)
            {
              double qj = (j == 0) ? 2 * M_PI / lb : 2 * M_PI / lb + j * deltaQ;
              double re1 = 0, im1 = 0, re2 = 0, im2 = 0;
              for (int i = 0; i < NA; i++)
                {
                  re1 += cos(qj * w.x[A][i]);
                  im1 += sin(qj * w.x[A][i]);
                }
              if (A != B)
                {
                  for (int i = 0; i < NB; i++)
                    {
                      re2 += cos(qj * w.x[B][i]);
                      im2 += sin(qj * w.x[B][i]);
                    }
                  double re = re1 + channel * re2, im = im1 + channel * im2;
                  sums[j] += (re * re + im * im) / (NA + NB);
                }
              else sums[j] += (re1 * re1 + im1 * im1) / (NA);
            }
          wsum += 1;
        }
      else
        {
          walker_free(&w);
          free(g);
          return fail("unknown estimator kind");
        }
      walker_free(&w);
    }
  free(g);
  *weight = wsum;
  return 0;
}

/* ---- forward walking (dmc/correlationEstimator.cpp) ----------------------------------------------
 * Restatement over a given genealogy (same arguments as ref_forward_walking in ref_shim.cpp).
 * Per walker: a ring data[recordSteps][bins], a time index and a filling flag
 * (realHistogramStorer::reserve, :109-115).  Every sampling event each walker first stores its own
 * histogram, sums/weight with weight 1 (store, :87-107), then the estimator adds the record
 * wrapIndex(i - forwardWalkingSteps - 1) of every walker whose ring has wrapped (accumulate, :128-152;
 * wrapIndex: tools.cpp:163-166).  Cloning copies the whole walker (walkers.cpp:139-179). */
typedef struct { double *data; int i, filling; } fw_walker;

int orc_forward_walking_model(const orc_model *m, const double *R, const int *n, const int *anc, int K, int iEst,
                              int recordSteps, int fwSteps, double *sums, int bins, double *weight)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  if (fwSteps >= recordSteps) return fail("forwardWalkingSteps < recordSteps (assert, correlationEstimator.cpp:137)");
  for (int b = 0; b < bins; b++) sums[b] = 0;
  *weight = 0;
  fw_walker *pop = NULL, *next = NULL;
  int npop = 0, rc = 0;
  size_t off = 0;
  const size_t ring = (size_t)recordSteps * bins;
  double *tmp = malloc(sizeof(double) * bins);
  for (int e = 0; e < K && !rc; e++)
    {
      next = malloc(sizeof(fw_walker) * (n[e] > 0 ? n[e] : 1));
      for (int i = 0; i < n[e]; i++)
        {
          next[i].data = malloc(sizeof(double) * ring);
          if (e == 0)
            {
              for (size_t k = 0; k < ring; k++) next[i].data[k] = 0;
              next[i].i = 0;
              next[i].filling = 1;
            }
          else
            {
              const fw_walker *src = pop + anc[off + i];
              memcpy(next[i].data, src->data, sizeof(double) * ring);
              next[i].i = src->i;
              next[i].filling = src->filling;
            }
        }
      for (int i = 0; i < n[e] && !rc; i++)
        {
          fw_walker *w = next + i;
          double wt = 0;
          rc = orc_histogram_model(m, R + (off + i) * dofs, 1, iEst, tmp, bins, &wt);      /* tmpAcc.clear(); ob->accumulate(w,psi,tmpAcc) */
          if (rc) break;
          for (int j = 0; j < bins; j++) w->data[(size_t)w->i * bins + j] = tmp[j] / wt;
          w->i = (w->i + 1) % recordSteps;
          if (w->i == 0) w->filling = 0;
          {
            int t = w->i - fwSteps - 1;
            int j = (t >= 0) ? t % recordSteps : recordSteps + t;
            if (!w->filling)
              {
                for (int b = 0; b < bins; b++) sums[b] += w->data[(size_t)j * bins + b];
                *weight += 1;
              }
          }
        }
      for (int i = 0; i < npop; i++) free(pop[i].data);
      free(pop);
      pop = next;
      npop = n[e];
      off += n[e];
    }
  for (int i = 0; i < npop; i++) free(pop[i].data);
  free(pop);
  free(tmp);
  return rc;
}

/* exported for the scipy pin of oracle/bessel_k.h */
double orc_bessel_k(int n, double x) { return orc_cyl_bessel_k(n, x); }
void orc_bessel_k01e(double x, double *k0, double *k1) { orc_bessel_k01_scaled(x, k0, k1); }

/* ---- forceEnergy (dmc/energy.cpp:37-61) -------------------------------------------------------------
 * evaluateDerivatives, then ef = sum over sets of (grad.array()*grad.array()).sum() (Eigen reduces a column-major
 * N x 3 array; the order of that reduction is Eigen's, so agreement with the reference is to rounding, not bit-equal),
 * no phase gradients for Jastrow-only wavefunctions; returns 0.5*ef + V. */
int orc_force_energy_model(const orc_model *m, const double *R, int nW, double *out)
{
  int dofs = 0;
  for (int s = 0; s < m->nSpecies; s++) dofs += 3 * m->N[s];
  double *grad = malloc(sizeof(double) * (size_t)nW * dofs), *V = malloc(sizeof(double) * nW);
  int *valid = malloc(sizeof(int) * nW);
  int rc = orc_eval_model(m, R, nW, NULL, grad, NULL, V, NULL, valid);
  for (int w = 0; w < nW && !rc; w++)
    {
      if (!valid[w]) { rc = fail("invalid configuration"); break; }
      double ef = 0;
      size_t off = 0;
      for (int s = 0; s < m->nSpecies; s++)
        {
          double tmp = 0;
          for (int k = 0; k < 3 * m->N[s]; k++) tmp += grad[(size_t)w * dofs + off + k] * grad[(size_t)w * dofs + off + k];
          ef += tmp;
          off += 3 * m->N[s];
        }
      out[w] = 0.5 * ef + V[w];
    }
  free(grad);
  free(V);
  free(valid);
  return rc;
}

/* ---- superfluidFraction (dmc/correlationEstimator.cpp:187-357) ------------------------------------
 * Same genealogy arguments.  storedSets[nStored]: the sets the centre-of-mass storer keeps, in order of first
 * appearance over all superfluidFraction measurements (createSuperfluidFractionStorer, :363-410); recordSteps =
 * the first measurement's "bins" (:199-207); idxA/idxB: positions of the estimator's two sets inside storedSets
 * (estimatorsFactory.cpp:86-108).  Per walker: data[recordSteps][3*nStored], a time index (the filling flag is
 * written, never read).  store (:248-268): record i = {sum over particles of each coordinate of each stored set}
 * — plain coordinates, no minimum image; i = (i+1) % recordSteps.  accumulate (:322-357): recordSize =
 * data.size()/bins, j = wrapIndex(i-1, bins), for i' = 0..j: W = sum_d (A_i' - A_j)(B_i' - B_j); sums[j-i'] += W;
 * weights[j-i'] += 1. */
typedef struct { double *data; int i; } sfl_walker;

int orc_superfluid_fraction_model(const orc_model *m, const double *R, const int *n, const int *anc, int K,
                                  const int *storedSets, int nStored, int recordSteps, int idxA, int idxB,
                                  double *sums, double *weights, int bins)
{
  int dofs = 0, offs[64];
  if (m->nSpecies > 64) return fail("too many species");
  for (int s = 0; s < m->nSpecies; s++) { offs[s] = dofs; dofs += 3 * m->N[s]; }
  const int width = 3 * nStored;
  const size_t ring = (size_t)recordSteps * width;
  const size_t recordSize = ring / bins;                                   /* :333 */
  for (int b = 0; b < bins; b++) sums[b] = weights[b] = 0;
  sfl_walker *pop = NULL, *next = NULL;
  int npop = 0;
  size_t off = 0;
  for (int e = 0; e < K; e++)
    {
      next = malloc(sizeof(sfl_walker) * (n[e] > 0 ? n[e] : 1));
      for (int i = 0; i < n[e]; i++)
        {
          next[i].data = malloc(sizeof(double) * ring);
          if (e == 0)
            {
              for (size_t k = 0; k < ring; k++) next[i].data[k] = 0;     /* reserve, :209-215 */
              next[i].i = 0;
            }
          else
            {
              const sfl_walker *src = pop + anc[off + i];
              memcpy(next[i].data, src->data, sizeof(double) * ring);
              next[i].i = src->i;
            }
        }
      for (int i = 0; i < n[e]; i++)
        {
          sfl_walker *w = next + i;
          const double *x = R + (off + i) * dofs;
          for (int k = 0; k < nStored; k++)
            {
              const int s = storedSets[k], N = m->N[s];
              double xcm[3] = {0, 0, 0};
              for (int p = 0; p < N; p++)
                for (int d = 0; d < 3; d++) xcm[d] += x[offs[s] + d * N + p];
              for (int d = 0; d < 3; d++) w->data[(size_t)w->i * width + 3 * k + d] = xcm[d];
            }
          w->i = (w->i + 1) % recordSteps;
          {
            int t = w->i - 1;
            int j = (t >= 0) ? t % bins : bins + t;                        /* wrapIndex, tools.cpp:163-166 */
            for (int ip = 0; ip <= j; ip++)
              {
                double W = 0;
                for (int d = 0; d < 3; d++)
                  {
                    double x1i = w->data[ip * recordSize + 3 * idxA + d], x1j = w->data[j * recordSize + 3 * idxA + d];
                    double x2i = w->data[ip * recordSize + 3 * idxB + d], x2j = w->data[j * recordSize + 3 * idxB + d];
                    W += (x1i - x1j) * (x2i - x2j);
                  }
                sums[j - ip] += W;
                weights[j - ip] += 1;
              }
          }
        }
      for (int i = 0; i < npop; i++) free(pop[i].data);
      free(pop);
      pop = next;
      npop = n[e];
      off += n[e];
    }
  for (int i = 0; i < npop; i++) free(pop[i].data);
  free(pop);
  return 0;
}

/* ---- branching (dmc/branching.cpp) ---------------------------------------------------------------- */
double orc_energy_shift(const double *E, int n, double timeStep, double meanWalkers_, double deltaWalkers_)
{ /* setEnergyShift (branching.cpp:65-97); meanWalkers/deltaWalkers are size_t members (branching.h:23-24) */
  size_t meanWalkers = (size_t)meanWalkers_, deltaWalkers = (size_t)deltaWalkers_;
  double energyShift = 0;
  for (int i = 0; i < n; i++) energyShift += E[i];
  int population = n;
  energyShift /= population;
  if ((size_t)population > (meanWalkers + deltaWalkers)) energyShift += 1. / timeStep * log(meanWalkers * 1. / (meanWalkers + deltaWalkers));
  else if ((size_t)population < (meanWalkers - deltaWalkers)) energyShift += 1. / timeStep * log(meanWalkers * 1. / (meanWalkers - deltaWalkers));
  return energyShift;
}

int orc_n_descendants(double Enew, double Eold, double energyShift, double timeStep, double u)
{ /* branching.cpp:11-14 */
  return (int)(exp(-timeStep * (0.5 * (Enew + Eold) - energyShift)) + u);
}

int orc_branch(const int *nDesc_in, const double *E_in, int n, double *Eafter, int capOut)
{ /* branchingControl::branch (branching.cpp:16-63): swap-with-last kill, push_back copies */
  int cap = capOut;
  int *nd = malloc(sizeof(int) * (cap > n ? cap : n));
  double *E = Eafter;
  if (n > cap)
    {
      free(nd);
      return -1;
    }
  int newPop = 0;
  for (int i = 0; i < n; i++)
    {
      nd[i] = nDesc_in[i];
      E[i] = E_in[i];
      newPop += nd[i];
    }
  if (newPop == 0) nd[0] = 1;
  int size = n, j1 = 0;
  while (j1 < size)
    {
      int c = nd[j1];
      if (c == 0)
        {
          double tmp = E[j1];
          E[j1] = E[size - 1];
          E[size - 1] = tmp;
          nd[j1] = nd[size - 1];
          size--;
        }
      else
        {
          for (int k = 0; k < c - 1; k++)
            {
              if (size >= cap)
                {
                  free(nd);
                  return -1;
                }
              E[size] = E[j1];
              nd[size] = 1;
              size++;
            }
          nd[j1] = 1;
          j1++;
        }
    }
  free(nd);
  return size;
}

int orc_plan(const int *populations_in, int nRanks, int *sources, int *amounts)
{ /* determineLoadBalanceComunicationsAliasMethod (ptools.cpp:93-161) */
  int *populations = malloc(sizeof(int) * nRanks), *permutations = malloc(sizeof(int) * nRanks);
  int size = 0;
  for (int i = 0; i < nRanks; i++)
    {
      populations[i] = populations_in[i];
      permutations[i] = i;
      size += populations[i];
      sources[i] = 0;
      amounts[i] = 0;
    }
  int k = size / nRanks, r = size % nRanks;
  int i = 0;
  while ((i < nRanks) && (populations[i] <= k)) i += 1;
  int j = i + 1;
  while (j < nRanks)
    {
      if (populations[j] <= k)
        {
          int t = populations[i];
          populations[i] = populations[j];
          populations[j] = t;
          t = permutations[i];
          permutations[i] = permutations[j];
          permutations[j] = t;
          i += 1;
        }
      j += 1;
    }
  j = 0;
  while ((i > j) && (i < nRanks))
    {
      int J = permutations[j], I = permutations[i];
      sources[J] = I;
      amounts[J] = (k + (j < r ? 1 : 0)) - populations[j];
      populations[i] = populations[i] - amounts[J];
      if (populations[i] < (k + (i < r ? 1 : 0))) ++i;
      j++;
    }
  free(populations);
  free(permutations);
  return 0;
}
