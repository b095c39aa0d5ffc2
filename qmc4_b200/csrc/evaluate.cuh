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

#ifndef QMC_EXP_ESTRIN
#define QMC_EXP_ESTRIN 0
#endif

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
  double *base;        // this CTA's histogram replica (global memory, fp64) or nullptr
  double w;            // weight of this configuration (a small integer: branching multiplicity or +-1)
  unsigned *fx;        // this CTA's fixed-point pair histogram in shared memory: lo[stride] then hi[stride], or nullptr
};

// g(r) accumulation without global atomics on the hot path.  A contribution v = norm/r^2 is
// quantised to 64-bit fixed point (scale Estimator::fx_scale: a pair at r = maxx maps to 2^36, i.e.
// 1.5e-11 relative resolution on the smallest contribution) and added to a per-CTA histogram in
// shared memory with two NATIVE 32-bit shared atomics: the low limb add returns the old value, a
// wrap of the low limb is forwarded as +1 to the high limb by the thread that caused it, so the
// final 64-bit total is exact whatever the interleaving (two's complement: negative weights used by
// the reject/branching corrections cancel the speculative contribution bit for bit).  On B200 a
// 32-lane RED.F64 to L2 costs ~50 cycles of LSU per warp instruction, the two ATOMS ~10.
// Contributions >= 2^40 (r < maxx/16) take the exact fp64 global path (rare).
QMC_DEVICE void hist_add_fixed(unsigned *fx, int stride, int bin, long long x)
{
  const unsigned xlo = (unsigned)(unsigned long long)x, xhi = (unsigned)((unsigned long long)x >> 32);
  const unsigned old = atomicAdd(fx + bin, xlo);
  const unsigned carry = (old + xlo < old) ? 1u : 0u;
  if (xhi + carry != 0u) atomicAdd(fx + stride + bin, xhi + carry);
}
constexpr double HIST_FX_BYPASS = 1099511627776.0;   // 2^40


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
#if QMC_EXP_ESTRIN
  /* Estrin evaluation of 1 + f*(c1 + c2 f + ... + c10 f^9): 13 FP64 ops but only ~6 deep
     (Horner: 10 ops, 10 deep); c_j = p[4+j] */
  const double f2 = f * f;
  const double a0 = fma(p[6], f, p[5]);      /* c1 + c2 f  */
  const double a1 = fma(p[8], f, p[7]);      /* c3 + c4 f  */
  const double a2 = fma(p[10], f, p[9]);     /* c5 + c6 f  */
  const double a3 = fma(p[12], f, p[11]);    /* c7 + c8 f  */
  const double a4 = fma(p[14], f, p[13]);    /* c9 + c10 f */
  const double f4 = f2 * f2;
  const double b0 = fma(a1, f2, a0);
  const double b1 = fma(a3, f2, a2);
  const double f8 = f4 * f4;
  const double c0 = fma(b1, f4, b0);
  const double rr = fma(a4, f8, c0);
  double q = fma(rr, f, 1.0);
#else
  double q = p[14];
#pragma unroll
  for (int j = 13; j >= 5; --j) q = fma(q, f, p[j]);
  q = fma(q, f, 1.0);
#endif
  /* exp(-745) is the smallest double: clamp the exponent instead of flushing (result <= 2^-1000 ~ 1e-301,
     i.e. zero to every accumulator it can meet); keeps the scaling free of selects */
  n = max(n, -1000);
  int hi = __double2hiint(q) + (n << 20);
  return __hiloint2double(hi, __double2loint(q));
}

// One fused pair pass: Jastrow JK + pair potential VK.  (HIST only selects the kernel variant that
// also runs histogram_pass() on the final configuration; the hot loop is identical.)
// pair() returns the coefficient g with   grad_row += s*g*d,  grad_col -= s*g*d   (s = grad_scale()).
template <int JK, int VK, bool HIST> struct PairOp {
  const PairPass &P;
  const double *coef;
  double s0 = 0, s1 = 0, vs = 0;
  int cnt = 0;

  QMC_DEVICE PairOp(const PairPass &P_, const double *coef_, const Plan &plan, const HistCtx &h) : P(P_), coef(coef_)
  {
  }

  // MASKED=false: every lane holds a real pair (full tiles), no select on the hot path
  template <bool MASKED = true> QMC_DEVICE double pair(double r2, bool act)
  {
    if (!MASKED) act = true;
    double g = 0;
    const int jk = (JK == J_GENERIC) ? P.J.kind : JK;
    double r = 0;
    bool have_r = false;
    if (JK == J_GAUSSGAUSS)
      { /* u=u0 e, u'/r = -2 u0 a e, lap term = e*(8 a^2 u0 r2 - 12 a u0): sums of e and e*r2 */
        double e = exp_gg(r2, P.J.p);
        if (MASKED) e = act ? e : 0.0;
        s0 += e;
        s1 = fma(e, r2, s1);
        g = e;
      }
    else if (JK == J_GAUSSIAN)
      { /* u=-a r2, u'/r=-2a, u''=-2a */
        g = act ? 1.0 : 0.0;
        s0 = fma(g, r2, s0);
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

// ---------------------------------------------------------------------------------------------
// Multi-stream variants for full 32-wide tiles: the 32 (or 16) rotation steps of a tile are split
// into NS interleaved streams, each with its own visitor, so one lane works on NS independent
// pairs at a time (same row particle, columns (l+k+s*STEP) mod 32).  That gives the FP64 pipe NS
// independent dependency chains per warp (the exp Horner chain is 14 dependent DFMAs) and lets the
// NS pairs share the loop overhead and the constant loads.
#ifndef QMC_TILE_STREAMS
#define QMC_TILE_STREAMS 2
#endif
constexpr int TILE_STREAMS = QMC_TILE_STREAMS;
#ifndef QMC_SHARED_ROWACC
#define QMC_SHARED_ROWACC 1
#endif

template <int NS, class Op>
QMC_DEVICE void pair_tile_rect_ms(Op &op, const Box &box, const double *xa, int nsa, int rbase,
                                  const double *xb, int nsb, int cbase, int lane, double *acca, double *accb)
{ /* 32 rows x 32 columns, all lanes active */
  constexpr int STEPS = 32 / NS;
  const bool rowok = true;
  const int p = lane;
  const double xi = xa[rbase + p], yi = xa[nsa + rbase + p], zi = xa[2 * nsa + rbase + p];
  double ax[NS], ay[NS], az[NS], vx[NS], vy[NS], vz[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) ax[s] = ay[s] = az[s] = vx[s] = vy[s] = vz[s] = 0.0;
  const int src = (lane + 1) & 31;
  const double *xc = xb + cbase, *yc = xb + nsb + cbase, *zc = xb + 2 * nsb + cbase;
#pragma unroll 1
  for (int k = 0; k < STEPS; ++k)
    {
      if (k > 0)
        {
#pragma unroll
          for (int s = 0; s < NS; ++s)
            {
              vx[s] = shfl_d(vx[s], src);
              vy[s] = shfl_d(vy[s], src);
              vz[s] = shfl_d(vz[s], src);
            }
        }
      double dx[NS], dy[NS], dz[NS], r2[NS], g[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          const int q = (lane + k + s * STEPS) & 31;
          dx[s] = min_image(xi - xc[q], box.Lx, box.ix);
          dy[s] = min_image(yi - yc[q], box.Ly, box.iy);
          dz[s] = min_image(zi - zc[q], box.Lz, box.iz);
          r2[s] = fma(dz[s], dz[s], fma(dy[s], dy[s], dx[s] * dx[s]));
        }
#pragma unroll
      for (int s = 0; s < NS; ++s) g[s] = op.template pair<false>(r2[s], true);
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          constexpr int AS = QMC_SHARED_ROWACC ? 0 : 1;   /* one row accumulator for all streams, or one per stream */
          ax[s * AS] = fma(g[s], dx[s], ax[s * AS]);
          ay[s * AS] = fma(g[s], dy[s], ay[s * AS]);
          az[s * AS] = fma(g[s], dz[s], az[s * AS]);
          vx[s] = fma(-g[s], dx[s], vx[s]);
          vy[s] = fma(-g[s], dy[s], vy[s]);
          vz[s] = fma(-g[s], dz[s], vz[s]);
        }
    }
  const double sc = op.grad_scale();
  if (rowok)
    {
      double sx = ax[0], sy = ay[0], sz = az[0];
#pragma unroll
      for (int s = 1; s < NS; ++s)
        {
          sx += ax[s];
          sy += ay[s];
          sz += az[s];
        }
      acca[rbase + p] = fma(sc, sx, acca[rbase + p]);
      acca[nsa + rbase + p] = fma(sc, sy, acca[nsa + rbase + p]);
      acca[2 * nsa + rbase + p] = fma(sc, sz, acca[2 * nsa + rbase + p]);
    }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS; ++s)
    { /* after STEPS steps the visitor of stream s belongs to column (lane + STEPS-1 + s*STEPS) */
      const int q = (lane + STEPS - 1 + s * STEPS) & 31;
      accb[cbase + q] = fma(sc, vx[s], accb[cbase + q]);
      accb[nsb + cbase + q] = fma(sc, vy[s], accb[nsb + cbase + q]);
      accb[2 * nsb + cbase + q] = fma(sc, vz[s], accb[2 * nsb + cbase + q]);
      __syncwarp();
    }
}

template <int NS, class Op>
QMC_DEVICE void pair_tile_diag_ms(Op &op, const Box &box, const double *xs, int ns, int base, int lane, double *accw)
{ /* full block of 32: offsets k = 1..16, stream s takes offsets k + s*STEPS, k = 1..STEPS; offset 16 is half-populated */
  constexpr int STEPS = 16 / NS;
  const double *xc = xs + base, *yc = xs + ns + base, *zc = xs + 2 * ns + base;
  const double xi = xc[lane], yi = yc[lane], zi = zc[lane];
  double ax[NS], ay[NS], az[NS], vx[NS], vy[NS], vz[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) ax[s] = ay[s] = az[s] = vx[s] = vy[s] = vz[s] = 0.0;
  const int src = (lane + 1) & 31;
#pragma unroll 1
  for (int k = 1; k <= STEPS; ++k)
    {
      if (k > 1)
        {
#pragma unroll
          for (int s = 0; s < NS; ++s)
            {
              vx[s] = shfl_d(vx[s], src);
              vy[s] = shfl_d(vy[s], src);
              vz[s] = shfl_d(vz[s], src);
            }
        }
      double dx[NS], dy[NS], dz[NS], r2[NS], g[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          const int q = (lane + k + s * STEPS) & 31;
          dx[s] = min_image(xi - xc[q], box.Lx, box.ix);
          dy[s] = min_image(yi - yc[q], box.Ly, box.iy);
          dz[s] = min_image(zi - zc[q], box.Lz, box.iz);
          r2[s] = fma(dz[s], dz[s], fma(dy[s], dy[s], dx[s] * dx[s]));
        }
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          if (s == NS - 1) g[s] = op.template pair<true>(r2[s], !((k == STEPS) && (lane >= 16)));   /* offset 16: half-populated */
          else g[s] = op.template pair<false>(r2[s], true);
        }
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          constexpr int AS = QMC_SHARED_ROWACC ? 0 : 1;   /* one row accumulator for all streams, or one per stream */
          ax[s * AS] = fma(g[s], dx[s], ax[s * AS]);
          ay[s * AS] = fma(g[s], dy[s], ay[s * AS]);
          az[s * AS] = fma(g[s], dz[s], az[s * AS]);
          vx[s] = fma(-g[s], dx[s], vx[s]);
          vy[s] = fma(-g[s], dy[s], vy[s]);
          vz[s] = fma(-g[s], dz[s], vz[s]);
        }
    }
  const double sc = op.grad_scale();
  {
    double sx = ax[0], sy = ay[0], sz = az[0];
#pragma unroll
    for (int s = 1; s < NS; ++s)
      {
        sx += ax[s];
        sy += ay[s];
        sz += az[s];
      }
    accw[base + lane] = fma(sc, sx, accw[base + lane]);
    accw[ns + base + lane] = fma(sc, sy, accw[ns + base + lane]);
    accw[2 * ns + base + lane] = fma(sc, sz, accw[2 * ns + base + lane]);
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS; ++s)
    {
      const int q = (lane + STEPS + s * STEPS) & 31;
      accw[base + q] = fma(sc, vx[s], accw[base + q]);
      accw[ns + base + q] = fma(sc, vy[s], accw[ns + base + q]);
      accw[2 * ns + base + q] = fma(sc, vz[s], accw[2 * ns + base + q]);
      __syncwarp();
    }
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
              const int nrow = min(32, nA - b * 32);
              if (nrow == 32) pair_tile_rect_ms<TILE_STREAMS>(op, box, xa, nA, b * 32, xa, nA, c * 32, G.lane, acca, acca);
              else pair_tile_rect(op, box, xa, nA, b * 32, nrow, xa, nA, c * 32, 32, G.lane, acca, acca);
            }
          else
            {
              const int b = t - noff;
              const int nblk = min(32, nA - b * 32);
              if (nblk == 32) pair_tile_diag_ms<TILE_STREAMS>(op, box, xa, nA, b * 32, G.lane, acca);
              else pair_tile_diag(op, box, xa, nA, b * 32, nblk, G.lane, acca);
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
          const int ncol = min(32, nB - c * 32), nrow = min(32, nA - b * 32);
          if (ncol == 32 && nrow == 32) pair_tile_rect_ms<TILE_STREAMS>(op, box, xa, nA, b * 32, xb, nB, c * 32, G.lane, acca, accb);
          else pair_tile_rect(op, box, xa, nA, b * 32, nrow, xb, nB, c * 32, ncol, G.lane, acca, accb);
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
// Triplet pass, every unordered triplet ONCE (jastrowWavefunctionThreeBody.h:91-137 visits i<j<k once; so does this).
// For a set of M <= 32 particles lane l owns particle l and meets the triplets (l, l+p, l+q) mod M for the canonical offset
// pairs (p, q) of the host-built table: the three gaps (p, q-p, M-q) of a triplet are defined up to rotation, the table keeps
// the lexicographically smallest rotation, so the M lanes x table entries enumerate each triplet exactly once (the fully
// symmetric gap triple M/3,M/3,M/3 has only M/3 distinct triplets: lanes < M/3).  All M lanes work at every step.
// The O(M^3) part only needs the pair table r2[gap][lane] (r2 between l and l+gap) and produces the symmetric pair weights
// G[gap][lane] += u'(R)/R with three conflict-free shared-memory read-modify-writes (distinct lanes hit distinct slots);
// an O(M^2) sweep turns the weights into gradients.  gaussiangaussian: u'/R = -2 alpha u, so R itself is never needed.
QMC_DEVICE int wrap_mod(int i, int M) { return i >= M ? i - M : i; }

QMC_DEVICE void run_triplet_pass_cyclic(const Plan &plan, const TripletPass &T, const double *coef, const Scratch &S, const Group &G,
                                        double &lp, double &lap, double &V, int &invalid)
{
  const int s = T.set, M = plan.N[s];
  const double *x = S.R + plan.off[s];
  double *r2t = S.tab;                                   /* [32][32]: r2t[d*32 + l] = |r_l - r_(l+d)|^2, d = 1..M-1 */
  double *Gw = S.tab + 1024 + (size_t)G.gwarp * 1024;    /* this warp's private pair weights, same indexing */
  double *Gt = S.tab + 1024;                             /* after the reduction: warp 0's table holds the totals */
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  const int lane = G.lane;
  for (int d = 1 + G.gwarp; d < M; d += G.nwarps)
    if (lane < M)
      {
        const int b = wrap_mod(lane + d, M);
        const double dx = min_image(x[lane] - x[b], box.Lx, box.ix);
        const double dy = min_image(x[M + lane] - x[M + b], box.Ly, box.iy);
        const double dz = min_image(x[2 * M + lane] - x[2 * M + b], box.Lz, box.iz);
        const double r = sqrt(dx * dx + dy * dy + dz * dz);      /* the reference squares the rounded norms */
        r2t[d * 32 + lane] = r * r;
      }
  for (int k = lane; k < 1024; k += 32) Gw[k] = 0.0;
  G.sync();
  const bool gg = T.J.kind == J_GAUSSGAUSS;
  double se = 0, ser = 0;
  int npot = 0, bad = 0;
  /* everything the loop needs in registers: the pass constants live in the kernel-parameter bank behind a run-time index */
  const int jkind = T.J.kind, has_pot = T.has_pot, has_con = T.has_constraint;
  const double v_thr2 = T.v_thr2, hs_thr2 = T.hs_thr2;
  double jp[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) jp[k] = T.J.p[k];
  const int *tabp = plan.trip_tab;
  const int nent = plan.trip_count, third = M / 3;
  int ent = G.gwarp < nent ? tabp[G.gwarp] : 0;
  for (int t = G.gwarp; t < nent; t += G.nwarps)
    {
      const int p = ent & 255, q = (ent >> 8) & 255, sym = ent >> 16;
      if (t + G.nwarps < nent) ent = tabp[t + G.nwarps];            /* next entry in flight while this one is worked on */
      const bool act = lane < (sym ? third : M);
      const int lb = wrap_mod(lane + p, M), lc = wrap_mod(lane + q, M);
      const int ia = p * 32 + lane, ib = (q - p) * 32 + lb, ic = (M - q) * 32 + lc;
      double R2 = r2t[ia] + r2t[ic] + r2t[ib];
      R2 = act ? R2 : 0.0;                    /* idle lanes read slots nobody wrote: keep NaN/Inf out of 0 * R2 */
      double w = 0.0;
      if (gg)
        {
          double e = exp_gg(R2, jp);
          e = act ? e : 0.0;
          se += e;
          ser = fma(e, R2, ser);
          w = e;
        }
      else if (jkind != J_NONE && act)
        {
          const double R = sqrt(R2);
          double d0, d1, d2;
          eval_radial(T.J, coef, R, d0, d1, d2);
          w = d1 / R;
          lp += d0;
          lap += d2 * 3 + 3 * w * 5;          /* (2D-1) = 5 */
        }
      npot += (has_pot && act && R2 <= v_thr2) ? 1 : 0;
      bad |= (has_con && act && R2 <= hs_thr2) ? 1 : 0;
      if (jkind != J_NONE)
        { /* three read-modify-writes; within each, the active lanes hit distinct slots */
          if (act) Gw[ia] += w;
          __syncwarp();
          if (act) Gw[ib] += w;
          __syncwarp();
          if (act) Gw[ic] += w;
          __syncwarp();
        }
    }
  invalid |= bad;
  if (gg)
    { /* u = u0 e ; u'/R = -2 a u ; 3u'' + 15 u'/R = u0 e (12 a^2 R^2 - 36 a) */
      const double a = T.J.p[0], u0 = T.J.p[1];
      lp += u0 * se;
      lap += u0 * (12.0 * a * a * ser - 36.0 * a * se);
    }
  V += T.v_val * (double)npot;
  G.sync();
  if (T.J.kind != J_NONE)
    {
      for (int k = G.gtid; k < 1024; k += G.gsize)
        {
          double g = Gt[k];
          for (int w = 1; w < G.nwarps; ++w) g += Gt[(size_t)w * 1024 + k];
          Gt[k] = g;
        }
      G.sync();
      const double scale = gg ? -2.0 * T.J.p[0] * T.J.p[1] : 1.0;
      double *accw = S.acc + (size_t)G.gwarp * plan.dofs + plan.off[s];
      for (int a = G.gtid; a < M; a += G.gsize)
        {
          double gx = 0, gy = 0, gz = 0;
          for (int d = 1; d < M; ++d)
            {
              const int b = wrap_mod(a + d, M);
              const double w = Gt[d * 32 + a] + Gt[(M - d) * 32 + b];      /* the pair (a, b) seen from a and from b */
              gx = fma(w, min_image(x[a] - x[b], box.Lx, box.ix), gx);
              gy = fma(w, min_image(x[M + a] - x[M + b], box.Ly, box.iy), gy);
              gz = fma(w, min_image(x[2 * M + a] - x[2 * M + b], box.Lz, box.iz), gz);
            }
          accw[a] = fma(scale, gx, accw[a]);
          accw[M + a] = fma(scale, gy, accw[M + a]);
          accw[2 * M + a] = fma(scale, gz, accw[2 * M + a]);
        }
    }
  G.sync();
}

// ---------------------------------------------------------------------------------------------
// Distinguishable triplet pass on three different sets (A,B,C): jastrow3bDis
// (jastrowWavefunctionThreeBody.h:170-300), potentialThreeBodyDis (potentialThreeBody.h:62-120),
// hardSphereConstraintThreeBodyDis (jastrowWavefunctionThreeBody.cpp:54-80); triplets (i in A, j in B,
// k in C) in LOOP3B_DIS order (tools3B.h:25-43), R^2 = r_ij^2 + r_jk^2 + r_ik^2 from the three
// minimum-image norms.  Same pair-weight formulation as the undistinguishable pass: the O(NA NB NC)
// part only produces G^AB_ij = sum_k u'(R)/R, G^BC_jk = sum_i, G^AC_ik = sum_j (three rectangular
// tables in shared memory next to the r^2 tables), an O(N^2) sweep turns them into gradients:
//   grad_i += sum_j G^AB_ij d^AB_ij + sum_k G^AC_ik d^AC_ik,   grad_j += -sum_i G^AB_ij d^AB_ij + sum_k G^BC_jk d^BC_jk,
//   grad_k -= sum_i G^AC_ik d^AC_ik + sum_j G^BC_jk d^BC_jk          (d^XY = min-image(x_X - x_Y), :262-266).
// Not inlined: it is a cold path that must not weigh on the register allocation of the pair loop.
struct TripletSums {
  double lp, lap, V;
  int invalid;
};

QMC_DEVICE double dis_r2(const Box &box, const double *xa, int na, int i, const double *xb, int nb, int j)
{
  const double dx = min_image(xa[i] - xb[j], box.Lx, box.ix);
  const double dy = min_image(xa[na + i] - xb[nb + j], box.Ly, box.iy);
  const double dz = min_image(xa[2 * na + i] - xb[2 * nb + j], box.Lz, box.iz);
  const double r = sqrt(dx * dx + dy * dy + dz * dz);      /* the reference squares the rounded norms */
  return r * r;
}

// r^2 tables of the three rectangular pair tables (A,B), (B,C), (A,C) into tab[0 .. nAB+nBC+nAC)
QMC_DEVICE void dis_fill_tables(const Plan &plan, const TripletPass &T, const Scratch &S, const Group &G, double *tab)
{
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  const int nA = plan.N[T.set], nB = plan.N[T.setB], nC = plan.N[T.setC];
  const double *xA = S.R + plan.off[T.set], *xB = S.R + plan.off[T.setB], *xC = S.R + plan.off[T.setC];
  const int nAB = nA * nB, nBC = nB * nC, nAC = nA * nC;
  for (int k = G.gtid; k < nAB + nBC + nAC; k += G.gsize)
    {
      if (k < nAB) tab[k] = dis_r2(box, xA, nA, k / nB, xB, nB, k % nB);
      else if (k < nAB + nBC) tab[k] = dis_r2(box, xB, nB, (k - nAB) / nC, xC, nC, (k - nAB) % nC);
      else tab[k] = dis_r2(box, xA, nA, (k - nAB - nBC) / nC, xC, nC, (k - nAB - nBC) % nC);
    }
}

// Scratch and Group cross the call BY VALUE, member by member: taking their address in the caller would push
// them to local memory and the pair loop would lose the shared-memory address space of S.R (LD instead of LDS).
struct ColdArgs {
  double *R, *acc, *red, *tab;
  int gtid, gsize, gwarp, nwarps, lane, bar_id;
};
QMC_DEVICE ColdArgs cold_args(const Scratch &S, const Group &G)
{
  ColdArgs a = {S.R, S.acc, S.red, S.tab, G.gtid, G.gsize, G.gwarp, G.nwarps, G.lane, G.bar_id};
  return a;
}
QMC_DEVICE void cold_unpack(const ColdArgs &a, Scratch &S, Group &G)
{
  S.R = a.R;
  S.acc = a.acc;
  S.red = a.red;
  S.tab = a.tab;
  S.aux = nullptr;
  G.gtid = a.gtid;
  G.gsize = a.gsize;
  G.gwarp = a.gwarp;
  G.nwarps = a.nwarps;
  G.lane = a.lane;
  G.bar_id = a.bar_id;
}

static __device__ QMC_NOINLINE TripletSums run_triplet_pass_dis(const Plan &plan, int t, const double *coef, ColdArgs ca)
{
  const TripletPass &T = plan.p3[t];
  Scratch S;
  Group G;
  cold_unpack(ca, S, G);
  TripletSums out = {0.0, 0.0, 0.0, 0};
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  const int nA = plan.N[T.set], nB = plan.N[T.setB], nC = plan.N[T.setC];
  const double *xA = S.R + plan.off[T.set], *xB = S.R + plan.off[T.setB], *xC = S.R + plan.off[T.setC];
  const int nAB = nA * nB, nBC = nB * nC, nAC = nA * nC, nT = nAB + nBC + nAC;
  double *r2AB = S.tab, *r2BC = r2AB + nAB, *r2AC = r2BC + nBC;
  double *gAB = S.tab + nT, *gBC = gAB + nAB, *gAC = gBC + nBC;
  const bool hasJ = T.J.kind != J_NONE;
  dis_fill_tables(plan, T, S, G, S.tab);
  G.sync();
  // (i,j) owners: scalars once per triplet + G^AB
  for (int p = G.gtid; p < nAB; p += G.gsize)
    {
      const int i = p / nB, j = p - i * nB;
      const double rij2 = r2AB[p];
      double g = 0;
      for (int k = 0; k < nC; ++k)
        {
          const double R = sqrt(rij2 + r2BC[j * nC + k] + r2AC[i * nC + k]);
          double d0 = 0, d1 = 0, d2 = 0;
          if (hasJ) eval_radial(T.J, coef, R, d0, d1, d2);
          g += d1 / R;
          out.lp += d0;
          out.lap += d2 * 3 + 3 * d1 / R * 5;      /* (2D-1) = 5 */
          if (T.has_pot && R <= T.v_R0) out.V += T.v_val;
          if (T.has_constraint && R <= T.hs_R0) out.invalid = 1;
        }
      gAB[p] = g;
    }
  if (hasJ)
    {
      for (int p = G.gtid; p < nBC + nAC; p += G.gsize)
        {
          double g = 0;
          if (p < nBC)
            {
              const int j = p / nC, k = p - j * nC;
              for (int i = 0; i < nA; ++i)
                {
                  const double R = sqrt(r2AB[i * nB + j] + r2BC[p] + r2AC[i * nC + k]);
                  double d0, d1, d2;
                  eval_radial(T.J, coef, R, d0, d1, d2);
                  g += d1 / R;
                }
              gBC[p] = g;
            }
          else
            {
              const int q = p - nBC, i = q / nC, k = q - i * nC;
              for (int j = 0; j < nB; ++j)
                {
                  const double R = sqrt(r2AB[i * nB + j] + r2BC[j * nC + k] + r2AC[q]);
                  double d0, d1, d2;
                  eval_radial(T.J, coef, R, d0, d1, d2);
                  g += d1 / R;
                }
              gAC[q] = g;
            }
        }
      G.sync();
      double *accw = S.acc + (size_t)G.gwarp * plan.dofs;
      double *aA = accw + plan.off[T.set], *aB = accw + plan.off[T.setB], *aC = accw + plan.off[T.setC];
      for (int p = G.gtid; p < nA + nB + nC; p += G.gsize)
        {
          double gx = 0, gy = 0, gz = 0;
          if (p < nA)
            {
              const int i = p;
              for (int j = 0; j < nB; ++j)
                {
                  const double w = gAB[i * nB + j];
                  gx = fma(w, min_image(xA[i] - xB[j], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xA[nA + i] - xB[nB + j], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xA[2 * nA + i] - xB[2 * nB + j], box.Lz, box.iz), gz);
                }
              for (int k = 0; k < nC; ++k)
                {
                  const double w = gAC[i * nC + k];
                  gx = fma(w, min_image(xA[i] - xC[k], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xA[nA + i] - xC[nC + k], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xA[2 * nA + i] - xC[2 * nC + k], box.Lz, box.iz), gz);
                }
              aA[i] += gx;
              aA[nA + i] += gy;
              aA[2 * nA + i] += gz;
            }
          else if (p < nA + nB)
            {
              const int j = p - nA;
              for (int i = 0; i < nA; ++i)
                {
                  const double w = -gAB[i * nB + j];
                  gx = fma(w, min_image(xA[i] - xB[j], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xA[nA + i] - xB[nB + j], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xA[2 * nA + i] - xB[2 * nB + j], box.Lz, box.iz), gz);
                }
              for (int k = 0; k < nC; ++k)
                {
                  const double w = gBC[j * nC + k];
                  gx = fma(w, min_image(xB[j] - xC[k], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xB[nB + j] - xC[nC + k], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xB[2 * nB + j] - xC[2 * nC + k], box.Lz, box.iz), gz);
                }
              aB[j] += gx;
              aB[nB + j] += gy;
              aB[2 * nB + j] += gz;
            }
          else
            {
              const int k = p - nA - nB;
              for (int i = 0; i < nA; ++i)
                {
                  const double w = -gAC[i * nC + k];
                  gx = fma(w, min_image(xA[i] - xC[k], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xA[nA + i] - xC[nC + k], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xA[2 * nA + i] - xC[2 * nC + k], box.Lz, box.iz), gz);
                }
              for (int j = 0; j < nB; ++j)
                {
                  const double w = -gBC[j * nC + k];
                  gx = fma(w, min_image(xB[j] - xC[k], box.Lx, box.ix), gx);
                  gy = fma(w, min_image(xB[nB + j] - xC[nC + k], box.Ly, box.iy), gy);
                  gz = fma(w, min_image(xB[2 * nB + j] - xC[2 * nC + k], box.Lz, box.iz), gz);
                }
              aC[k] += gx;
              aC[nC + k] += gy;
              aC[2 * nC + k] += gz;
            }
        }
    }
  G.sync();
  return out;
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
      for (int t = 0; t < plan.n3b; ++t)
        {
          if (plan.p3[t].setB < 0 && plan.trip_tab && plan.N[plan.p3[t].set] == plan.trip_mod)
            run_triplet_pass_cyclic(plan, plan.p3[t], coef, S, G, lp, lap, V, invalid);
          else if (plan.p3[t].setB < 0) run_triplet_pass(plan, plan.p3[t], coef, S, G, H, lp, lap, V, invalid);
          else
            {
              const TripletSums ts = run_triplet_pass_dis(plan, t, coef, cold_args(S, G));
              lp += ts.lp;
              lap += ts.lap;
              V += ts.V;
              invalid |= ts.invalid;
            }
        }
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

// g(r) pass over the pairs of the FINAL configuration (in S.R) with its branching weight H.w
// (pairCorrelation::accumulate, pairCorrelation.cpp:17-39, over the post-branching population,
// dmcDriver.cpp:257-264).  Same tiles as the physics pass but no gradients, so no visitors and no
// shuffles: lane l holds row particle l and meets column (l+k) mod 32; UNR pairs are in flight per
// lane.  Bins are filled through hist_add_fixed (shared memory, native 32-bit atomics).
template <int UNR>
QMC_DEVICE void hist_tile(const Box &box, const double *xa, int nsa, int rbase, int nr, const double *xb, int nsb, int cbase, int nc,
                          bool diag, int lane, const Estimator &est, unsigned *fx, double *hist, int stride, int iw, double hw)
{
  const bool rowok = lane < nr;
  const int p = rowok ? lane : 0;
  const double xi = xa[rbase + p], yi = xa[nsa + rbase + p], zi = xa[2 * nsa + rbase + p];
  const double *xc = xb + cbase, *yc = xb + nsb + cbase, *zc = xb + 2 * nsb + cbase;
  const double maxx = est.maxx, minx = est.minx, dxinv = est.dxinv, fxs = est.fx_scale;
  const int last = est.bins - 1;
  // diag: offsets 1..nc/2 (cyclic schedule, half-populated last offset when nc is even); else 0..31
  const int k0 = diag ? 1 : 0, k1 = diag ? (nc >> 1) : 31;
  for (int kb = k0; kb <= k1; kb += UNR)
    {
      double r2[UNR];
      bool ok[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        {
          const int k = kb + u;
          int q = p + k;
          if (diag) { if (q >= nc) q -= nc; }
          else q &= 31;
          ok[u] = rowok && (k <= k1) && (diag ? !((2 * k == nc) && (lane >= (nc >> 1))) : (q < nc));
          const int qq = ok[u] ? q : 0;
          const double dx = min_image_abs(xi - xc[qq], box.Lx, box.ix);
          const double dy = min_image_abs(yi - yc[qq], box.Ly, box.iy);
          const double dz = min_image_abs(zi - zc[qq], box.Lz, box.iz);
          r2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
        }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        {
          const double y = rsqrt_fast(r2[u]);
          const double rr = r2[u] * y;
          const int b = (int)((rr - minx) * dxinv);
          const double t = y * y * fxs;
          if (ok[u] && rr < maxx && b >= 0)
            {
              if (fx && t < HIST_FX_BYPASS)
                {
                  long long x = (long long)t;
                  if (iw != 1) x *= iw;
                  hist_add_fixed(fx, stride, min(b, last), x);
                }
              else atomicAdd(hist + min(b, last), hw * y * y);
            }
        }
    }
}

// one binned pair (shared by the tile variants below)
QMC_DEVICE void hist_bin_pair(double r2, bool ok, double maxx, double minx, double dxinv, double fxs, int last, unsigned *fx, double *hist,
                              int stride, int iw, double hw)
{
  const double y = rsqrt_fast(r2);
  const double rr = r2 * y;
  const int b = (int)((rr - minx) * dxinv);
  const double t = y * y * fxs;
  if (ok && rr < maxx && b >= 0)
    {
      if (fx && t < HIST_FX_BYPASS)
        {
          long long x = (long long)t;
          if (iw != 1) x *= iw;
          hist_add_fixed(fx, stride, min(b, last), x);
        }
      else atomicAdd(hist + min(b, last), hw * y * y);
    }
}

// Full tiles (32 rows x 32 columns, or a full diagonal block): no predication, incremental column index.
template <int UNR, bool DIAG>
QMC_DEVICE void hist_tile_full(const Box &box, const double *xa, int nsa, int rbase, const double *xb, int nsb, int cbase, int lane,
                               const Estimator &est, unsigned *fx, double *hist, int stride, int iw, double hw)
{
  const double xi = xa[rbase + lane], yi = xa[nsa + rbase + lane], zi = xa[2 * nsa + rbase + lane];
  const double *xc = xb + cbase, *yc = xb + nsb + cbase, *zc = xb + 2 * nsb + cbase;
  const double maxx = est.maxx, minx = est.minx, dxinv = est.dxinv, fxs = est.fx_scale;
  const int last = est.bins - 1;
  // DIAG: offsets 1..15 full + offset 16 half-populated ; else offsets 0..31
  constexpr int K0 = DIAG ? 1 : 0, NFULL = DIAG ? 15 : 32;
  static_assert(NFULL % UNR == 0 || DIAG, "unroll must divide 32");
  int kb = K0;
  for (; kb + UNR <= K0 + NFULL; kb += UNR)
    {
      double r2[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        {
          const int q = (lane + kb + u) & 31;
          const double dx = min_image_abs(xi - xc[q], box.Lx, box.ix);
          const double dy = min_image_abs(yi - yc[q], box.Ly, box.iy);
          const double dz = min_image_abs(zi - zc[q], box.Lz, box.iz);
          r2[u] = fma(dz, dz, fma(dy, dy, dx * dx));
        }
#pragma unroll
      for (int u = 0; u < UNR; ++u) hist_bin_pair(r2[u], true, maxx, minx, dxinv, fxs, last, fx, hist, stride, iw, hw);
    }
  if (DIAG)
    { /* offsets left over (15 is not a multiple of UNR) and the half-populated offset 16 */
      for (; kb <= 16; ++kb)
        {
          const int q = (lane + kb) & 31;
          const double dx = min_image_abs(xi - xc[q], box.Lx, box.ix);
          const double dy = min_image_abs(yi - yc[q], box.Ly, box.iy);
          const double dz = min_image_abs(zi - zc[q], box.Lz, box.iz);
          const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
          hist_bin_pair(r2, !(kb == 16 && lane >= 16), maxx, minx, dxinv, fxs, last, fx, hist, stride, iw, hw);
        }
    }
}

QMC_DEVICE void histogram_pass(const Plan &plan, const Scratch &S, const Group &G, const HistCtx &H)
{
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  for (int t = 0; t < plan.n2b; ++t)
    {
      const PairPass &P = plan.p2[t];
      if (P.hist < 0) continue;
      const Estimator &est = plan.est[P.hist];
      double *hist = H.base + est.offset;
      unsigned *fx = H.fx ? H.fx + est.offset : nullptr;
      const int iw = (int)H.w;
      const double hw = H.w * est.norm;
      const int nA = plan.N[P.setA], nB = plan.N[P.setB];
      const double *xa = S.R + plan.off[P.setA], *xb = S.R + plan.off[P.setB];
      if (P.setA == P.setB)
        {
          const int nb = (nA + 31) >> 5, noff = nb * (nb - 1) / 2;
          for (int w = G.gwarp; w < noff + nb; w += G.nwarps)
            {
              if (w < noff)
                {
                  int b = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)w)) * 0.5f);
                  while (b * (b - 1) / 2 > w) --b;
                  while ((b + 1) * b / 2 <= w) ++b;
                  const int c = w - b * (b - 1) / 2;
                  if (nA - b * 32 >= 32) hist_tile_full<4, false>(box, xa, nA, b * 32, xa, nA, c * 32, G.lane, est, fx, hist, plan.hist_stride, iw, hw);
                  else hist_tile<4>(box, xa, nA, b * 32, nA - b * 32, xa, nA, c * 32, 32, false, G.lane, est, fx, hist, plan.hist_stride, iw, hw);
                }
              else
                {
                  const int b = w - noff, n = min(32, nA - b * 32);
                  if (n == 32) hist_tile_full<5, true>(box, xa, nA, b * 32, xa, nA, b * 32, G.lane, est, fx, hist, plan.hist_stride, iw, hw);
                  else hist_tile<4>(box, xa, nA, b * 32, n, xa, nA, b * 32, n, true, G.lane, est, fx, hist, plan.hist_stride, iw, hw);
                }
            }
        }
      else
        {
          const int nba = (nA + 31) >> 5, nbb = (nB + 31) >> 5;
          for (int w = G.gwarp; w < nba * nbb; w += G.nwarps)
            {
              const int b = w / nbb, c = w - b * nbb;
              hist_tile<4>(box, xa, nA, b * 32, min(32, nA - b * 32), xb, nB, c * 32, min(32, nB - c * 32), false, G.lane, est, fx, hist,
                           plan.hist_stride, iw, hw);
            }
        }
    }
}

// zero / flush the CTA's fixed-point histogram (all threads of the CTA; callers synchronise)
QMC_DEVICE void hist_fx_zero(unsigned *fx, int stride)
{
  for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) fx[i] = 0u;
}
QMC_DEVICE void hist_fx_flush(const Plan &plan, unsigned *fx, double *replica)
{
  const int stride = plan.hist_stride;
  for (int e = 0; e < plan.nest; ++e)
    {
      const Estimator &est = plan.est[e];
      if (est.fx_scale == 0.0) continue;
      const double inv = est.norm / est.fx_scale;
      for (int b = threadIdx.x; b < est.bins; b += blockDim.x)
        {
          const unsigned lo = atomicExch(fx + est.offset + b, 0u), hi = atomicExch(fx + stride + est.offset + b, 0u);
          const long long t = (long long)(((unsigned long long)hi << 32) | (unsigned long long)lo);
          if (t != 0) atomicAdd(replica + est.offset + b, (double)t * inv);
        }
    }
}

// Three-set observables on the FINAL configuration (cold path, not inlined): trimerCorrelation
// (trimerCorrelation.cpp:9-41, set C holds one particle), trimerCorrelationDis (:163-199) and the scalar
// trimerhyperRadius (hyperRadius.cpp:17-43: mean of R over the (i,j) pairs).  The order of the three
// squared norms under the square root follows each reference routine, so bins match bit for bit.
static __device__ QMC_NOINLINE void post_three_set_observable(const Plan &plan, int e, ColdArgs ca, double *replica, double w)
{
  const Estimator &est = plan.est[e];
  Scratch S;
  Group G;
  cold_unpack(ca, S, G);
  TripletPass T;
  T.set = est.setA;
  T.setB = est.setB;
  T.setC = est.setC;
  const int nA = plan.N[T.set], nB = plan.N[T.setB], nC = plan.N[T.setC];
  const int nAB = nA * nB, nBC = nB * nC;
  const double *r2AB = S.tab, *r2BC = r2AB + nAB, *r2AC = r2BC + nBC;
  G.sync();
  dis_fill_tables(plan, T, S, G, S.tab);
  G.sync();
  if (est.kind == 6 /* QMC_EST_TRIMERHYPERRADIUS */)
    {
      double sum = 0;
      for (int p = G.gtid; p < nAB; p += G.gsize)
        {
          const int i = p / nB, j = p - i * nB;
          sum += sqrt(r2AC[i] + r2BC[j] + r2AB[p]);
        }
      sum = warp_sum(sum);
      if (G.lane == 0) S.red[G.gwarp * 4] = sum;
      G.sync();
      if (G.gtid == 0)
        {
          double tot = 0;
          for (int q = 0; q < G.nwarps; ++q) tot += S.red[q * 4];
          atomicAdd(replica + est.offset, w * (tot / (double)(nA * nB)));
        }
    }
  else
    {
      double *hist = replica + est.offset;
      const double hw = w * est.norm;
      const bool plain = est.kind == 4 /* QMC_EST_TRIMERCORRELATION */;
      for (int p = G.gtid; p < nAB; p += G.gsize)
        {
          const int i = p / nB, j = p - i * nB;
          for (int k = 0; k < nC; ++k)
            {
              const double ab = r2AB[p], bc = r2BC[j * nC + k], ac = r2AC[i * nC + k];
              const double R = plain ? sqrt(ac + bc + ab) : sqrt(ab + ac + bc);
              if (R < est.maxx)
                {
                  const int bin = (int)((R - est.minx) * est.dxinv);
                  if (bin >= 0) atomicAdd(hist + min(bin, est.bins - 1), hw / (R * R));
                }
            }
        }
    }
  G.sync();
}

// centerOfMassSquared (centerOfMassSquared.cpp:8-18): (sum_i |r_i|)^2 / N with the minimum-imaged distance
// from the origin (geometry.cpp:9-25) — the reference's definition, not the centre of mass.
static __device__ QMC_NOINLINE void post_center_of_mass_squared(const Plan &plan, int e, ColdArgs ca, double *replica, double w)
{
  const Estimator &est = plan.est[e];
  Scratch S;
  Group G;
  cold_unpack(ca, S, G);
  const int n = plan.N[est.setA];
  const double *x = S.R + plan.off[est.setA];
  double sum = 0;
  for (int i = G.gtid; i < n; i += G.gsize)
    {
      const double dx = min_image(x[i], plan.L[0], plan.Linv[0]);
      const double dy = min_image(x[n + i], plan.L[1], plan.Linv[1]);
      const double dz = min_image(x[2 * n + i], plan.L[2], plan.Linv[2]);
      sum += sqrt(dx * dx + dy * dy + dz * dz);
    }
  sum = warp_sum(sum);
  G.sync();
  if (G.lane == 0) S.red[G.gwarp * 4] = sum;
  G.sync();
  if (G.gtid == 0)
    {
      double tot = 0;
      for (int q = 0; q < G.nwarps; ++q) tot += S.red[q * 4];
      atomicAdd(replica + est.offset, w * (tot * tot / (double)n));
    }
  G.sync();
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 + 8 + 4 + 2 + 2 = 32 shuffles of 64 bits instead of 16 x 5: at every
// stage a lane keeps half of its values and sends the other half to its partner.  On return lanes 2v and 2v+1 hold the total of
// value v in v16[0].
QMC_DEVICE void warp_reduce16(double (&v)[16], int lane)
{
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1)
    {
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i)
        {
          const double keep = up ? v[i + half] : v[i];
          const double send = up ? v[i] : v[i + half];
          v[i] = keep + shfl_xor_d(send, bit);
        }
    }
  v[0] += shfl_xor_d(v[0], 1);
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
      if (T.hist < 0 || T.setB >= 0) continue;
      const Estimator &est = plan.est[T.hist];
      double *hist = H.base + est.offset;
      const double hw = H.w * est.norm;
      const int n = plan.N[T.set], P = n * (n - 1) / 2;
      const double *x = S.R + plan.off[T.set];
      double *r2tab = S.tab;
      G.sync();
      if (plan.trip_tab && n == plan.trip_mod)
        { /* cyclic schedule (run_triplet_pass_cyclic): r2[gap][lane] table, every triplet once, sqrt / divide / atomic only for the
             triplets inside the histogram range */
          const int M = n, lane = G.lane;
          for (int d = 1 + G.gwarp; d < M; d += G.nwarps)
            if (lane < M)
              {
                const int b = wrap_mod(lane + d, M);
                const double dx = min_image(x[lane] - x[b], box.Lx, box.ix);
                const double dy = min_image(x[M + lane] - x[M + b], box.Ly, box.iy);
                const double dz = min_image(x[2 * M + lane] - x[2 * M + b], box.Lz, box.iz);
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                r2tab[d * 32 + lane] = r * r;
              }
          G.sync();
          const double pre = est.maxx * est.maxx * (1.0 + 1e-12);
          for (int t = G.gwarp; t < plan.trip_count; t += G.nwarps)
            {
              const int ent = plan.trip_tab[t];
              const int p = ent & 255, q = (ent >> 8) & 255, sym = ent >> 16;
              if (lane >= (sym ? M / 3 : M)) continue;
              const int lb = wrap_mod(lane + p, M), lc = wrap_mod(lane + q, M);
              const double R2 = r2tab[p * 32 + lane] + r2tab[(M - q) * 32 + lc] + r2tab[(q - p) * 32 + lb];
              if (R2 < pre)
                {
                  const double R = sqrt(R2);
                  if (R < est.maxx)
                    {
                      const int bin = (int)((R - est.minx) * est.dxinv);
                      if (bin >= 0) atomicAdd(hist + min(bin, est.bins - 1), hw / (R * R));
                    }
                }
            }
          G.sync();
          continue;
        }
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
      if (est.kind >= 4 && est.kind <= 6) post_three_set_observable(plan, e, cold_args(S, G), H.base, H.w);
      else if (est.kind == 7 /* QMC_EST_CENTEROFMASSSQUARED */) post_center_of_mass_squared(plan, e, cold_args(S, G), H.base, H.w);
    }
  for (int e = 0; e < plan.nest; ++e)
    {
      const Estimator &est = plan.est[e];
      if (est.kind != 3 /* QMC_EST_STRUCTUREFACTOR */) continue;
      const int bins = est.bins;
      double *auxw = S.aux + (size_t)G.gwarp * 2 * bins;
      constexpr int CH = 8;
      G.sync();
      const int nsets = (est.setA == est.setB) ? 1 : 2;
      const int nA = plan.N[est.setA], ntot = nA + (nsets == 2 ? plan.N[est.setB] : 0);
      constexpr int PMAX = 8;          /* particles per thread whose phasors are carried in registers */
      if (ntot <= PMAX * G.gsize)
        { /* e^{i q_j x} = e^{i q_0 x} (e^{i dq x})^j: two sincos per particle in all, the phasor of every particle advances through
             the whole q grid by the complex recurrence (4 FP64 per point); rounding grows like bins x 1e-16 */
          double pc[PMAX], ps[PMAX], pdc[PMAX], pds[PMAX];
#pragma unroll
          for (int p = 0; p < PMAX; ++p)
            {
              const int i = G.gtid + p * G.gsize;
              pc[p] = ps[p] = pdc[p] = pds[p] = 0.0;
              if (i < ntot)
                {
                  const bool inA = i < nA;
                  const double x = inA ? S.R[plan.off[est.setA] + i] : S.R[plan.off[est.setB] + (i - nA)];      /* coordinate 0 */
                  const double cf = inA ? 1.0 : (double)est.channel;
                  double s0, c0;
                  sincos(est.q0 * x, &s0, &c0);
                  sincos(est.dq * x, &pds[p], &pdc[p]);
                  pc[p] = c0 * cf;
                  ps[p] = s0 * cf;
                }
            }
          for (int j0 = 0; j0 < bins; j0 += CH)
            {
              double re[CH], im[CH];
#pragma unroll
              for (int m = 0; m < CH; ++m) re[m] = im[m] = 0.0;
#pragma unroll
              for (int p = 0; p < PMAX; ++p)
                {
                  if (p * G.gsize >= ntot) break;
                  double c = pc[p], sn = ps[p];
                  const double dc = pdc[p], ds = pds[p];
#pragma unroll
                  for (int m = 0; m < CH; ++m)
                    {
                      re[m] += c;
                      im[m] += sn;
                      const double cn = c * dc - sn * ds;
                      sn = fma(sn, dc, c * ds);
                      c = cn;
                    }
                  pc[p] = c;
                  ps[p] = sn;
                }
              { /* 16 sums over the warp in one transposed reduction: lane 2v ends with value v (v < 8: re[v], else im[v-8]) */
                double v16[16];
#pragma unroll
                for (int m = 0; m < CH; ++m)
                  {
                    v16[m] = re[m];
                    v16[CH + m] = im[m];
                  }
                warp_reduce16(v16, G.lane);
                const int vi = G.lane >> 1, m = vi & 7;
                if ((G.lane & 1) == 0 && j0 + m < bins) auxw[2 * (j0 + m) + (vi >> 3)] = v16[0];
              }
            }
        }
      else
      for (int j0 = 0; j0 < bins; j0 += CH)
        {
          double re[CH], im[CH];
#pragma unroll
          for (int m = 0; m < CH; ++m) re[m] = im[m] = 0.0;
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
