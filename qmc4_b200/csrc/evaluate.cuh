// evaluate.cuh — evaluation of one walker by one "walker group" (1..8 whole warps of a CTA):
// log psi, grad log psi, laplacian(log psi), V and the estimator histograms in ONE pass over
// the minimum-image pairs.  It fuses what the reference does in four O(N^2) passes
// (SURVEY.md §3.3): geometryPBC::differencesTwoBody (geometry.cpp:47-65) + norm (:27-44) +
// potentialTwoBody::operator() (potentialTwoBody.h:23-33) + jastrowTwoBody*::accumulateDerivatives
// (jastrowWavefunctionTwoBody.h:33-76,124-164), plus pairCorrelation::accumulate
// (pairCorrelation.cpp:17-39) on sampling steps.  No distance table is ever written to memory.
//
// Pair-loop organisation (DESIGN.md "pair tiles"): particles are cut in blocks of 32.  A warp
// owns a tile (row block, column block); lane l keeps row particle l in registers and at step k
// meets column particle (l+k) mod n, read from shared memory.  The gradient on the ROW particle
// accumulates in the lane's registers; the gradient on the COLUMN particle accumulates in a
// "visitor" triple that is rotated one lane per step with warp shuffles, so each unordered pair
// is evaluated exactly once and nothing is scattered with atomics.  Diagonal tiles use the
// cyclic schedule k = 1..n/2 (last step half-populated when n is even).  Each warp adds its tile
// results into a private per-warp accumulator in shared memory; the group reduces them in a
// fixed order, so results are bitwise reproducible run to run.
#pragma once
#include "plan.h"
#include "device_math.cuh"
#include "jastrows.cuh"

namespace qmc {

struct Group {
  int gtid;      // thread index inside the group
  int gsize;     // threads in the group (32*nwarps)
  int gwarp;     // warp index inside the group
  int nwarps;
  int lane;
  int bar_id;    // named barrier of this group
  QMC_DEVICE void sync() const
  {
    if (nwarps == 1) __syncwarp();
    else group_sync(bar_id, gsize);
  }
};

struct Scratch {
  double *R;      // [dofs]   configuration being evaluated (walker layout)
  double *acc;    // [nwarps][dofs] private gradient accumulators; acc[0] holds the total after evaluate
  double *red;    // [nwarps*8] scalar reduction slots
  double *tab;    // [2*P] pair table (r^2, G) for triplet passes, or nullptr
  double *aux;    // [nwarps*2*bins] structure-factor partial sums, or nullptr
};

struct EvalResult {
  double logpsi, lap, V;
  int valid;
};

struct HistCtx {
  double *base;   // this CTA's histogram replica (global memory) or nullptr
  double w;       // weight of this configuration
};

// ---------------------------------------------------------------------------------------------
// exp(-alpha*r2) with the scaled-argument constants stored in RadialFn::p[2..14]
// (p2 = -alpha*log2e, p3/p4 = hi/lo of ln2/(-alpha), p5..p14 = T_j*(-alpha)^j, j=1..10).
QMC_DEVICE double exp_gg(double r2, const double *p)
{
  const double MAGIC = 6755399441055744.0;
  double t = fma(r2, p[2], MAGIC);
  int n = __double2loint(t);
  double nd = t - MAGIC;
  double f = fma(nd, -p[3], r2);
  f = fma(nd, -p[4], f);
  double q = p[14];
#pragma unroll
  for (int j = 13; j >= 5; --j) q = fma(q, f, p[j]);
  q = fma(q, f, 1.0);
  int hi = __double2hiint(q) + (n << 20);
  double r = __hiloint2double(hi, __double2loint(q));
  return (n < -1000) ? 0.0 : r;
}

// One fused pair pass: Jastrow JK + pair potential VK (+ g(r) when HIST).
// pair() returns the coefficient g with   grad_row += s*g*d,  grad_col -= s*g*d   (s = grad_scale()).
template <int JK, int VK, bool HIST> struct PairOp {
  const PairPass &P;
  const double *coef;
  const Estimator *est;
  double *hist;
  double hw;
  double s0 = 0, s1 = 0, vs = 0;
  int cnt = 0;

  QMC_DEVICE PairOp(const PairPass &P_, const double *coef_, const Plan &plan, const HistCtx &h) : P(P_), coef(coef_)
  {
    est = nullptr;
    hist = nullptr;
    hw = 0;
    if (HIST && P.hist >= 0 && h.base)
      {
        est = &plan.est[P.hist];
        hist = h.base + est->offset;
        hw = h.w * est->norm;
      }
  }

  QMC_DEVICE double pair(double r2, bool act)
  {
    double g = 0;
    const int jk = (JK == J_GENERIC) ? P.J.kind : JK;
    double r = 0;
    bool have_r = false;
    if (JK == J_GAUSSGAUSS)
      { /* u=u0 e, u'/r = -2 u0 a e, lap term = e*(8 a^2 u0 r2 - 12 a u0): sums of e and e*r2 */
        double e = exp_gg(r2, P.J.p);
        e = act ? e : 0.0;
        s0 += e;
        s1 = fma(e, r2, s1);
        g = e;
      }
    else if (JK == J_GAUSSIAN)
      { /* u=-a r2, u'/r=-2a, u''=-2a */
        g = act ? 1.0 : 0.0;
        s0 = fma(g, r2, s0);
        cnt += 0;
      }
    else if (jk != J_NONE)
      {
        r = sqrt(r2);
        have_r = true;
        double d0, d1, d2;
        eval_radial(P.J, coef, r, d0, d1, d2);
        double gr = d1 / r;
        if (act)
          {
            s0 += d0;
            s1 += 2 * (d2 + 2 * gr);     /* jastrowWavefunctionTwoBody.h:150, D=3 */
            g = gr;
          }
      }
    const int vk = (VK == V_GENERIC) ? P.vkind : VK;
    if (VK == V_STEP) { cnt += (act && r2 <= P.v_thr2) ? 1 : 0; }
    else if (vk == V_STEP) { cnt += (act && r2 <= P.v_thr2) ? 1 : 0; }
    else if (vk != V_NONE)
      {
        if (!have_r) { r = sqrt(r2); have_r = true; }
        double v = eval_pair_potential(vk, P.v_val, r);
        if (act) vs += v;
      }
    if (HIST)
      {
        if (hist)
          { /* pairCorrelation.cpp:24-35 + histogramAccumulator::accumulate (accumulators.h:117) */
            double y = rsqrt_fast(r2);
            double rr = r2 * y;
            if (act && rr < est->maxx)
              {
                int b = (int)((rr - est->minx) * est->dxinv);
                if (b >= 0) atomicAdd(hist + min(b, est->bins - 1), hw * y * y);
              }
          }
      }
    return g;
  }

  QMC_DEVICE double grad_scale() const
  {
    if (JK == J_GAUSSGAUSS) return -2.0 * P.J.p[1] * P.J.p[0];
    if (JK == J_GAUSSIAN) return -2.0 * P.J.p[0];
    return 1.0;
  }
  // contributions of this pass to (log psi, laplacian, V); per-thread partials
  QMC_DEVICE void finish(double &lp, double &lap, double &V) const
  {
    if (JK == J_GAUSSGAUSS)
      {
        const double a = P.J.p[0], u0 = P.J.p[1];
        lp += u0 * s0;
        lap += -12.0 * a * u0 * s0 + 8.0 * a * a * u0 * s1;
      }
    else if (JK == J_GAUSSIAN)
      {
        const double a = P.J.p[0];
        lp += -a * s0;
        lap += 0;   /* -12a per pair: added analytically by the caller (pair count) */
      }
    else
      {
        lp += s0;
        lap += s1;
      }
    V += vs + P.v_val * (double)cnt * ((VK == V_STEP || (VK == V_GENERIC && P.vkind == V_STEP)) ? 1.0 : 0.0);
  }
};

struct Box {
  double Lx, Ly, Lz, ix, iy, iz;
};

// ---------------------------------------------------------------------------------------------
// Diagonal tile: all unordered pairs inside one block of n <= 32 particles (cyclic schedule).
// xs: set base in shared memory with dimension stride ns; accw: this warp's private accumulator.
template <class Op>
QMC_DEVICE void pair_tile_diag(Op &op, const Box &box, const double *xs, int ns, int base, int n, int lane, double *accw)
{
  const bool active = lane < n;
  const int p = active ? lane : 0;
  const double xi = xs[base + p], yi = xs[ns + base + p], zi = xs[2 * ns + base + p];
  double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0;
  const int K = n >> 1;
  const int src = (lane + 1 < n) ? lane + 1 : 0;
  for (int k = 1; k <= K; ++k)
    {
      if (k > 1)
        {
          vx = shfl_d(vx, src);
          vy = shfl_d(vy, src);
          vz = shfl_d(vz, src);
        }
      int q = p + k;
      if (q >= n) q -= n;
      const bool act = active && !((2 * k == n) && (lane >= K));
      double dx = min_image(xi - xs[base + q], box.Lx, box.ix);
      double dy = min_image(yi - xs[ns + base + q], box.Ly, box.iy);
      double dz = min_image(zi - xs[2 * ns + base + q], box.Lz, box.iz);
      double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      double g = op.pair(r2, act);
      ax = fma(g, dx, ax);
      ay = fma(g, dy, ay);
      az = fma(g, dz, az);
      vx = fma(-g, dx, vx);
      vy = fma(-g, dy, vy);
      vz = fma(-g, dz, vz);
    }
  const double s = op.grad_scale();
  if (active)
    {
      accw[base + p] = fma(s, ax, accw[base + p]);
      accw[ns + base + p] = fma(s, ay, accw[ns + base + p]);
      accw[2 * ns + base + p] = fma(s, az, accw[2 * ns + base + p]);
    }
  __syncwarp();
  if (active && K >= 1)
    {
      int q = p + K;
      if (q >= n) q -= n;
      accw[base + q] = fma(s, vx, accw[base + q]);
      accw[ns + base + q] = fma(s, vy, accw[ns + base + q]);
      accw[2 * ns + base + q] = fma(s, vz, accw[2 * ns + base + q]);
    }
  __syncwarp();
}

// Rectangular tile: nr <= 32 rows (set a) x nc <= 32 columns (set b), every pair once.
template <class Op>
QMC_DEVICE void pair_tile_rect(Op &op, const Box &box, const double *xa, int nsa, int rbase, int nr,
                               const double *xb, int nsb, int cbase, int nc, int lane, double *acca, double *accb)
{
  const bool rowok = lane < nr;
  const int p = rowok ? lane : 0;
  const double xi = xa[rbase + p], yi = xa[nsa + rbase + p], zi = xa[2 * nsa + rbase + p];
  double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0;
  const int src = (lane + 1) & 31;
#pragma unroll 2
  for (int k = 0; k < 32; ++k)
    {
      if (k > 0)
        {
          vx = shfl_d(vx, src);
          vy = shfl_d(vy, src);
          vz = shfl_d(vz, src);
        }
      const int q0 = (lane + k) & 31;
      const bool colok = q0 < nc;
      const int q = colok ? q0 : 0;
      const bool act = rowok && colok;
      double dx = min_image(xi - xb[cbase + q], box.Lx, box.ix);
      double dy = min_image(yi - xb[nsb + cbase + q], box.Ly, box.iy);
      double dz = min_image(zi - xb[2 * nsb + cbase + q], box.Lz, box.iz);
      double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      double g = op.pair(r2, act);
      ax = fma(g, dx, ax);
      ay = fma(g, dy, ay);
      az = fma(g, dz, az);
      vx = fma(-g, dx, vx);
      vy = fma(-g, dy, vy);
      vz = fma(-g, dz, vz);
    }
  const double s = op.grad_scale();
  if (rowok)
    {
      acca[rbase + p] = fma(s, ax, acca[rbase + p]);
      acca[nsa + rbase + p] = fma(s, ay, acca[nsa + rbase + p]);
      acca[2 * nsa + rbase + p] = fma(s, az, acca[2 * nsa + rbase + p]);
    }
  __syncwarp();
  {
    const int q = (lane + 31) & 31;   /* column this lane's visitor belongs to after 32 steps */
    if (q < nc)
      {
        accb[cbase + q] = fma(s, vx, accb[cbase + q]);
        accb[nsb + cbase + q] = fma(s, vy, accb[nsb + cbase + q]);
        accb[2 * nsb + cbase + q] = fma(s, vz, accb[2 * nsb + cbase + q]);
      }
  }
  __syncwarp();
}

// All tiles of one pair pass, distributed round-robin over the warps of the group.
template <int JK, int VK, bool HIST>
QMC_DEVICE void run_pair_pass(const Plan &plan, const PairPass &P, const double *coef, const Scratch &S, const Group &G,
                              const HistCtx &H, double &lp, double &lap, double &V)
{
  PairOp<JK, VK, HIST> op(P, coef, plan, H);
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  double *accw = S.acc + (size_t)G.gwarp * plan.dofs;
  const int sA = P.setA, sB = P.setB;
  const int nA = plan.N[sA], nB = plan.N[sB];
  const double *xa = S.R + plan.off[sA];
  const double *xb = S.R + plan.off[sB];
  double *acca = accw + plan.off[sA];
  double *accb = accw + plan.off[sB];
  if (sA == sB)
    {
      const int nb = (nA + 31) >> 5;
      const int noff = nb * (nb - 1) / 2;
      for (int t = G.gwarp; t < noff + nb; t += G.nwarps)
        {
          if (t < noff)
            {
              int b = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
              while (b * (b - 1) / 2 > t) --b;
              while ((b + 1) * b / 2 <= t) ++b;
              const int c = t - b * (b - 1) / 2;
              pair_tile_rect(op, box, xa, nA, b * 32, min(32, nA - b * 32), xa, nA, c * 32, 32, G.lane, acca, acca);
            }
          else
            {
              const int b = t - noff;
              pair_tile_diag(op, box, xa, nA, b * 32, min(32, nA - b * 32), G.lane, acca);
            }
        }
      if (JK == J_GAUSSIAN && G.gtid == 0) lap += -12.0 * P.J.p[0] * (0.5 * (double)nA * (double)(nA - 1));
    }
  else
    {
      const int nba = (nA + 31) >> 5, nbb = (nB + 31) >> 5;
      for (int t = G.gwarp; t < nba * nbb; t += G.nwarps)
        {
          const int b = t / nbb, c = t - b * nbb;
          pair_tile_rect(op, box, xa, nA, b * 32, min(32, nA - b * 32), xb, nB, c * 32, min(32, nB - c * 32), G.lane, acca, accb);
        }
      if (JK == J_GAUSSIAN && G.gtid == 0) lap += -12.0 * P.J.p[0] * ((double)nA * (double)nB);
    }
  op.finish(lp, lap, V);
}

// ---------------------------------------------------------------------------------------------
// Triplet pass (jastrowWavefunctionThreeBody.h:91-137, potentialThreeBody.h:23-44,
// trimerCorrelation.cpp:82-115, constraint jastrowWavefunctionThreeBody.cpp:15-36).
// The gradient of u(R), R^2 = r_ab^2+r_ac^2+r_bc^2, on particle a is
//   sum_{b != a} (r_a - r_b) * G_ab,   G_ab = sum_{c != a,b} u'(R_abc)/R_abc,
// so the O(N^3) part only produces the symmetric pair weights G (no vector scatter), and an O(N^2)
// sweep turns them into gradients.  Scalars are counted once per triplet (c > a > b).
QMC_DEVICE int pair_index(int a, int b) { return a > b ? a * (a - 1) / 2 + b : b * (b - 1) / 2 + a; }

QMC_DEVICE void run_triplet_pass(const Plan &plan, const TripletPass &T, const double *coef, const Scratch &S, const Group &G,
                                 const HistCtx &H, double &lp, double &lap, double &V, int &invalid)
{
  const int s = T.set, n = plan.N[s];
  const int P = n * (n - 1) / 2;
  const double *x = S.R + plan.off[s];
  double *r2tab = S.tab, *gtab = S.tab + P;
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  for (int k = G.gtid; k < P; k += G.gsize)
    {
      int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)k)) * 0.5f);
      while (a * (a - 1) / 2 > k) --a;
      while ((a + 1) * a / 2 <= k) ++a;
      const int b = k - a * (a - 1) / 2;
      double dx = min_image(x[a] - x[b], box.Lx, box.ix);
      double dy = min_image(x[n + a] - x[n + b], box.Ly, box.iy);
      double dz = min_image(x[2 * n + a] - x[2 * n + b], box.Lz, box.iz);
      /* the reference squares the rounded norms: R^2 = r_ji*r_ji + ... with r = sqrt(sum d^2) */
      double r = sqrt(dx * dx + dy * dy + dz * dz);
      r2tab[k] = r * r;
    }
  G.sync();
  for (int k = G.gtid; k < P; k += G.gsize)
    {
      int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)k)) * 0.5f);
      while (a * (a - 1) / 2 > k) --a;
      while ((a + 1) * a / 2 <= k) ++a;
      const int b = k - a * (a - 1) / 2;   /* a > b */
      const double rab2 = r2tab[k];
      double Gab = 0;
      for (int c = 0; c < n; ++c)
        {
          if (c == a || c == b) continue;
          const double R2 = rab2 + r2tab[pair_index(a, c)] + r2tab[pair_index(b, c)];
          const double R = sqrt(R2);
          double d0 = 0, d1 = 0, d2 = 0;
          if (T.J.kind != J_NONE) eval_radial(T.J, coef, R, d0, d1, d2);
          Gab += d1 / R;
          if (c > a)
            { /* this thread owns the triplet b < a < c */
              lp += d0;
              lap += d2 * 3 + 3 * d1 / R * 5;     /* (2D-1) = 5 */
              if (T.has_pot && R <= T.v_R0) V += T.v_val;
              if (T.has_constraint && R <= T.hs_R0) invalid = 1;
            }
        }
      gtab[k] = Gab;
    }
  G.sync();
  if (T.J.kind != J_NONE)
    {
      double *accw = S.acc + (size_t)G.gwarp * plan.dofs + plan.off[s];
      for (int a = G.gtid; a < n; a += G.gsize)
        {
          double gx = 0, gy = 0, gz = 0;
          for (int b = 0; b < n; ++b)
            {
              if (b == a) continue;
              const double w = gtab[pair_index(a, b)];
              gx = fma(w, min_image(x[a] - x[b], box.Lx, box.ix), gx);
              gy = fma(w, min_image(x[n + a] - x[n + b], box.Ly, box.iy), gy);
              gz = fma(w, min_image(x[2 * n + a] - x[2 * n + b], box.Lz, box.iz), gz);
            }
          accw[a] += gx;
          accw[n + a] += gy;
          accw[2 * n + a] += gz;
        }
    }
  G.sync();
}

// ---------------------------------------------------------------------------------------------
// One-body terms: jastrow1b (jastrowWavefunctionOneBody.h:53-85; laplacian has NO factor 2) and
// harmonic1b (potential.cpp:21-32).  The "distance from the origin" is minimum-imaged too
// (geometry.cpp:9-25).
QMC_DEVICE void run_one_body(const Plan &plan, const OneBodyTerm &T, const double *coef, const Scratch &S, const Group &G,
                             double &lp, double &lap, double &V)
{
  const int s = T.set, n = plan.N[s];
  const double *x = S.R + plan.off[s];
  double *accw = S.acc + (size_t)G.gwarp * plan.dofs + plan.off[s];
  for (int i = G.gtid; i < n; i += G.gsize)
    {
      double dx = min_image(x[i], plan.L[0], plan.Linv[0]);
      double dy = min_image(x[n + i], plan.L[1], plan.Linv[1]);
      double dz = min_image(x[2 * n + i], plan.L[2], plan.Linv[2]);
      double r = sqrt(dx * dx + dy * dy + dz * dz);
      if (T.J.kind != J_NONE)
        {
          double d0, d1, d2;
          eval_radial(T.J, coef, r, d0, d1, d2);
          lap += d2 + 2 * d1 / r;
          lp += d0;
          accw[i] += dx / r * d1;
          accw[n + i] += dy / r * d1;
          accw[2 * n + i] += dz / r * d1;
        }
      if (T.has_pot) V += T.half_omega2 * r * r;
    }
}

// ---------------------------------------------------------------------------------------------
// Full evaluation of the configuration in S.R.  On return (all threads of the group):
//   result = {log psi, laplacian, V, valid};  S.acc[0..dofs) = grad log psi.
// Equivalent to tableDistances::update + satisfyConstraints + energy::operator()
// (walkers.cpp:46-61, energy.cpp:6-35) without the distance table.
template <int JK, int VK, bool HIST>
QMC_DEVICE EvalResult evaluate_walker(const Plan &plan, const double *coef, const Scratch &S, const Group &G, const HistCtx &H)
{
  const int dofs = plan.dofs;
  {
    double *accw = S.acc + (size_t)G.gwarp * dofs;
    for (int k = G.lane; k < dofs; k += 32) accw[k] = 0.0;
    __syncwarp();
  }
  double lp = 0, lap = 0, V = 0;
  int invalid = 0;
  for (int t = 0; t < plan.n1b; ++t) run_one_body(plan, plan.t1[t], coef, S, G, lp, lap, V);
  __syncwarp();
  if (plan.n2b > 0) run_pair_pass<JK, VK, HIST>(plan, plan.p2[0], coef, S, G, H, lp, lap, V);
  for (int t = 1; t < plan.n2b; ++t) run_pair_pass<J_GENERIC, V_GENERIC, HIST>(plan, plan.p2[t], coef, S, G, H, lp, lap, V);
  if (plan.n3b > 0)
    {
      G.sync();
      for (int t = 0; t < plan.n3b; ++t) run_triplet_pass(plan, plan.p3[t], coef, S, G, H, lp, lap, V, invalid);
    }
  // scalar reduction: warp shuffle, then fixed-order sum over the group's warps
  lp = warp_sum(lp);
  lap = warp_sum(lap);
  V = warp_sum(V);
  invalid = warp_sum_i(invalid);
  if (G.lane == 0)
    {
      double *r = S.red + G.gwarp * 4;
      r[0] = lp;
      r[1] = lap;
      r[2] = V;
      r[3] = (double)invalid;
    }
  G.sync();
  EvalResult out;
  out.logpsi = 0;
  out.lap = 0;
  out.V = 0;
  double inv = 0;
  for (int w = 0; w < G.nwarps; ++w)
    {
      out.logpsi += S.red[w * 4 + 0];
      out.lap += S.red[w * 4 + 1];
      out.V += S.red[w * 4 + 2];
      inv += S.red[w * 4 + 3];
    }
  out.valid = inv == 0.0;
  // gradient reduction over the private accumulators, fixed order
  if (G.nwarps > 1)
    {
      for (int k = G.gtid; k < dofs; k += G.gsize)
        {
          double g = S.acc[k];
          for (int w = 1; w < G.nwarps; ++w) g += S.acc[(size_t)w * dofs + k];
          S.acc[k] = g;
        }
    }
  G.sync();
  return out;
}

// Histogram-only pass over the pairs/triplets of the configuration in S.R with weight H.w
// (corrects the speculative unit-weight accumulation done inside evaluate_walker; see
// propagate kernel).  Only estimators attached to pair passes / triplet passes are touched.
QMC_DEVICE void histogram_only(const Plan &plan, const Scratch &S, const Group &G, const HistCtx &H)
{
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  for (int t = 0; t < plan.n2b; ++t)
    {
      const PairPass &P = plan.p2[t];
      if (P.hist < 0) continue;
      const Estimator &est = plan.est[P.hist];
      double *hist = H.base + est.offset;
      const double hw = H.w * est.norm;
      const int nA = plan.N[P.setA], nB = plan.N[P.setB];
      const double *xa = S.R + plan.off[P.setA], *xb = S.R + plan.off[P.setB];
      const long npairs = (P.setA == P.setB) ? (long)nA * (nA - 1) / 2 : (long)nA * nB;
      for (long k = G.gtid; k < npairs; k += G.gsize)
        {
          int a, b;
          if (P.setA == P.setB)
            {
              a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)k)) * 0.5f);
              while ((long)a * (a - 1) / 2 > k) --a;
              while ((long)(a + 1) * a / 2 <= k) ++a;
              b = (int)(k - (long)a * (a - 1) / 2);
            }
          else
            {
              a = (int)(k / nB);
              b = (int)(k - (long)a * nB);
            }
          double dx = min_image(xa[a] - xb[b], box.Lx, box.ix);
          double dy = min_image(xa[nA + a] - xb[nB + b], box.Ly, box.iy);
          double dz = min_image(xa[2 * nA + a] - xb[2 * nB + b], box.Lz, box.iz);
          double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
          double y = rsqrt_fast(r2);
          double rr = r2 * y;
          if (rr < est.maxx)
            {
              int bin = (int)((rr - est.minx) * est.dxinv);
              if (bin >= 0) atomicAdd(hist + min(bin, est.bins - 1), hw * y * y);
            }
        }
    }
}

// Estimators that are evaluated on the FINAL configuration (in S.R) with its branching weight:
// trimerCorrelationUnDis (trimerCorrelation.cpp:82-143) and the structure factor
// (structureFactor.cpp:56-90 applied to coordinate 0; q_j = q0 + j*dq, structureFactor.cpp:23-40).
QMC_DEVICE void post_estimators(const Plan &plan, const Scratch &S, const Group &G, const HistCtx &H)
{
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  for (int t = 0; t < plan.n3b; ++t)
    {
      const TripletPass &T = plan.p3[t];
      if (T.hist < 0) continue;
      const Estimator &est = plan.est[T.hist];
      double *hist = H.base + est.offset;
      const double hw = H.w * est.norm;
      const int n = plan.N[T.set], P = n * (n - 1) / 2;
      const double *x = S.R + plan.off[T.set];
      double *r2tab = S.tab;
      G.sync();
      for (int k = G.gtid; k < P; k += G.gsize)
        {
          int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)k)) * 0.5f);
          while (a * (a - 1) / 2 > k) --a;
          while ((a + 1) * a / 2 <= k) ++a;
          const int b = k - a * (a - 1) / 2;
          double dx = min_image(x[a] - x[b], box.Lx, box.ix);
          double dy = min_image(x[n + a] - x[n + b], box.Ly, box.iy);
          double dz = min_image(x[2 * n + a] - x[2 * n + b], box.Lz, box.iz);
          double r = sqrt(dx * dx + dy * dy + dz * dz);
          r2tab[k] = r * r;
        }
      G.sync();
      for (int k = G.gtid; k < P; k += G.gsize)
        {
          int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)k)) * 0.5f);
          while (a * (a - 1) / 2 > k) --a;
          while ((a + 1) * a / 2 <= k) ++a;
          const int b = k - a * (a - 1) / 2;
          const double rab2 = r2tab[k];
          for (int c = a + 1; c < n; ++c)
            {
              const double R = sqrt(rab2 + r2tab[pair_index(a, c)] + r2tab[pair_index(b, c)]);
              if (R < est.maxx)
                {
                  int bin = (int)((R - est.minx) * est.dxinv);
                  if (bin >= 0) atomicAdd(hist + min(bin, est.bins - 1), hw / (R * R));
                }
            }
        }
      G.sync();
    }
  for (int e = 0; e < plan.nest; ++e)
    {
      const Estimator &est = plan.est[e];
      if (est.kind != 3 /* QMC_EST_STRUCTUREFACTOR */) continue;
      const int bins = est.bins;
      double *auxw = S.aux + (size_t)G.gwarp * 2 * bins;
      constexpr int CH = 8;
      G.sync();
      for (int j0 = 0; j0 < bins; j0 += CH)
        {
          double re[CH], im[CH];
#pragma unroll
          for (int m = 0; m < CH; ++m) re[m] = im[m] = 0.0;
          const int nsets = (est.setA == est.setB) ? 1 : 2;
          for (int ss = 0; ss < nsets; ++ss)
            {
              const int set = ss == 0 ? est.setA : est.setB;
              const double cf = ss == 0 ? 1.0 : (double)est.channel;
              const int n = plan.N[set];
              const double *x = S.R + plan.off[set];      /* coordinate 0 */
              for (int i = G.gtid; i < n; i += G.gsize)
                {
                  double s, c, ds, dc;
                  sincos((est.q0 + j0 * est.dq) * x[i], &s, &c);
                  sincos(est.dq * x[i], &ds, &dc);
                  c *= cf;
                  s *= cf;
#pragma unroll
                  for (int m = 0; m < CH; ++m)
                    {
                      re[m] += c;
                      im[m] += s;
                      const double cn = c * dc - s * ds;
                      s = fma(s, dc, c * ds);
                      c = cn;
                    }
                }
            }
#pragma unroll
          for (int m = 0; m < CH; ++m)
            {
              const double r = warp_sum(re[m]), q = warp_sum(im[m]);
              if (G.lane == 0 && j0 + m < bins)
                {
                  auxw[2 * (j0 + m)] = r;
                  auxw[2 * (j0 + m) + 1] = q;
                }
            }
        }
      G.sync();
      const double nrm = (est.setA == est.setB) ? (double)plan.N[est.setA] : (double)(plan.N[est.setA] + plan.N[est.setB]);
      for (int j = G.gtid; j < bins; j += G.gsize)
        {
          double r = 0, q = 0;
          for (int w = 0; w < G.nwarps; ++w)
            {
              r += S.aux[(size_t)w * 2 * bins + 2 * j];
              q += S.aux[(size_t)w * 2 * bins + 2 * j + 1];
            }
          atomicAdd(H.base + est.offset + j, H.w * (r * r + q * q) / nrm);
        }
      G.sync();
    }
}

} // namespace qmc
