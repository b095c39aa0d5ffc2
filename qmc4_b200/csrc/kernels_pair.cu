// kernels_pair.cu — the hot kernel of the BASELINE periodic-box shapes (configs 2, 3, 5): the move step of pipeline 2
// specialised for ONE species and ONE pair pass with the gaussiangaussian Jastrow (jastrowGaussGauss.h:19-21) and an
// optional step potential (potentialFunctors.h:5-36).  Same arithmetic per pair as the general path (evaluate.cuh:
// min_image, exp_gg, PairOp) and the same Philox draws as move_kernel; what changes is the organisation around the pair:
//
//  * tournament schedule.  The particles are cut in blocks of 32; the off-diagonal tiles (block a, block b) are visited
//    in nb-1 rounds of a round-robin tournament (circle method): in one round every block appears in at most one tile, so
//    the warps of a walker group add the row AND the column gradients of their tile straight into the ONE gradient
//    accumulator of the walker, without atomics and without per-warp private copies (general path: (1+warps) x dofs
//    doubles of shared memory per walker, 110 KB at N=512; here 2 x dofs + 1.5 KB per warp = 37 KB).  A named barrier
//    separates the rounds; the order of every floating-point sum is fixed, so results are bitwise reproducible.
//  * column staging.  A warp copies the 32 column particles of its tile into a private 3 x 64 strip in which every
//    coordinate is stored twice, so lane l reads column (l+k) mod 32 at the fixed offset l+k: one running pointer and
//    immediate offsets instead of a masked index computation per load.
//  * the visitor accumulators rotate at the end of every step of a tile (no conditional in front of the shuffles).
//  * minimum image in fixed point.  geometryPBC::difference (geometry.h:50) is t - round(t/L) L; here every coordinate of
//    the configuration is converted once per move to a 64-bit fraction of the box, X = rint(frac(x/L) 2^64), and the
//    minimum-image difference of a pair is the WRAPPING integer subtraction Xi - Xj: the two's-complement wrap-around is
//    the periodic wrap.  Resolution L 2^-64 (5e-20 L) — finer than the rounding of x itself; the pair then costs one
//    integer subtraction + one int64->fp64 conversion per component instead of 3 FP64-pipe instructions + FRND + LOP3.
//    In a cubic box the pair sums stay in box units (all scale factors are folded into the constants on the host).
//    Exact ties (t = +-L/2 to the last bit) map to -L/2 where the reference's round-half-away gives -sign(t) L/2: same
//    distance, measure zero after one diffusion step (the evaluation entry qmc_eval keeps the reference's rounding).
//
// Reference semantics: dmcDriver::update (dmcDriver.cpp:56-84) = driftDiffusionFirstOrder::move (dmcMoves.cpp:12-30) +
// tableDistances::update (tableDistances.cpp:81-108) + jastrowTwoBodyWavefunctionUndistinguishable::accumulateDerivatives
// (jastrowWavefunctionTwoBody.h:124-164) + potentialTwoBody (potentialTwoBody.h:23-33) + energy (energy.cpp:6-35) +
// transitionProbabilityRatio (dmcMoves.cpp:32-57) + metropolis::acceptLog (metropolis.cpp:5-29).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

// exp(-alpha r2) by table: x = -alpha r2 = (k + j/256) ln2 + f, |f| <= ln2/512, so exp(x) = 2^k * 2^(j/256) * exp(f) with the 256
// values 2^(j/256) in shared memory and exp(f) = 1 + f + f^2/2 + f^3/6 + f^4/24 (remainder f^5/120 < 4e-17).  8 FP64-pipe
// instructions (the degree-10 polynomial of the general path: 14).  All constants are folded into the r2 scale on the host.
struct LeanParams {
  double k;            // -alpha * 256 / ln2            : n = rint(r2 * k)
  double c;            // ln2 / (256 alpha)             : f' = r2 + n c   (f = -alpha f')
  double c1, c2, c3, c4;   // (-alpha)^j / j!
  double sc;           // -2 u0 alpha  (u'/r = sc * e)
  double u0, lap0, lap1;   // log psi = u0 sum e ; laplacian = lap0 sum e + lap1 sum e r2  (jastrowWavefunctionTwoBody.h:150, D = 3)
  double v_val;
  long long thr_bits;  // bit pattern of the largest r2 inside the step potential: r2 >= 0, so the integer order is the fp64 order
  const double *tab;   // [256] 2^(j/256), global memory
  // fixed-point minimum image: X = rint(frac(x Linv) 2^64); d = (double)(Xi - Xj) in units s_d = L_d 2^-64
  double s[3];         // L_d 2^-64
  double gs[3];        // sc * s_d : gradient = gs_d * sum e d'_d
  int cubic;           // 1: k, c, c1..c4, lap1, thr_bits are expressed in r2' = sum d'^2 (box units); 0: d is scaled to real units first
};

QMC_DEVICE long long to_box_fraction(const double x, const double Linv)
{
  double u = x * Linv;
  u -= rint(u);                                   /* [-0.5, 0.5] */
  return __double2ll_rn(u * 18446744073709551616.0);      /* saturates at +-2^63: one unit of 5e-20 L */
}
QMC_DEVICE double fx_diff(const long long a, const long long b)
{
  return __ll2double_rn((long long)((unsigned long long)a - (unsigned long long)b));   /* wrapping subtraction = minimum image */
}

// TS: table stride in doubles.  TS = 1: one copy, every lane gathers from the same 2 KB (random bank conflicts, 3.6 excess wavefronts per
// lookup measured); TS = 16: 16 interleaved copies, `tab` already points at the lane's own copy (lane & 15), so the 16 lanes of a half-warp
// hit 16 different 8-byte bank pairs whatever their indices: conflict-free by construction.
template <int TS>
QMC_DEVICE double exp_tab(const double r2, const LeanParams &E, const double *tab)
{
  const double MAGIC = 6755399441055744.0;      // 2^52 + 2^51
  const double t = fma(r2, E.k, MAGIC);
  int n = __double2loint(t);
  const double nd = t - MAGIC;
  const double f = fma(nd, E.c, r2);
  double q = fma(f, E.c4, E.c3);
  q = fma(q, f, E.c2);
  q = fma(q, f, E.c1);
  q = fma(q, f, 1.0);
  n = max(n, -256000);                          /* exp(-693) ~ 1e-301: zero to every accumulator it can meet */
  const double e = tab[(n & 255) * TS] * q;
  const int hi = __double2hiint(e) + ((n >> 8) << 20);
  return __hiloint2double(hi, __double2loint(e));
}

struct LeanSums {
  double s0, s1;      // sum e, sum e*r2 over this thread's pairs (e = exp(-alpha r2))
  int cnt;            // pairs inside the step potential
};

constexpr int STRIP = 192;     // doubles per warp: x[64] y[64] z[64], every column twice
#ifndef QMC_PAIR_UNROLL
#define QMC_PAIR_UNROLL 1
#endif
constexpr int PAIR_UNROLL = QMC_PAIR_UNROLL;

// One tile: rows = block rb (lane l owns row particle rb*32+l), columns = block cb.  DIAG: rb == cb, cyclic offsets 1..16
// (offset 16 half-populated).  MASKED: the row and/or column block is the ragged last block of the species.
// Two independent pairs per lane and step: stream 0 meets column (l + OFF0 + k), stream 1 column (l + OFF1 + k).
template <int VK, bool CUBIC, bool DIAG, bool MASKED, int TS>
QMC_DEVICE void lean_tile(const LeanParams &LP, const double *tab, const long long *xs, const int n, const int rb,
                          const int cb, long long *T, const int lane, double *acc, LeanSums &S)
{
  constexpr int IT = DIAG ? 8 : 16, OFF0 = DIAG ? 1 : 0, OFF1 = DIAG ? 9 : 16;
  const int ci = cb * 32 + lane;
  const bool cok = !MASKED || ci < n;
  const long long cx = cok ? xs[ci] : 0, cy = cok ? xs[n + ci] : 0, cz = cok ? xs[2 * n + ci] : 0;
  __syncwarp();                      /* the strip may still be read by a slower lane of the previous tile */
  T[lane] = cx;
  T[32 + lane] = cx;
  T[64 + lane] = cy;
  T[96 + lane] = cy;
  T[128 + lane] = cz;
  T[160 + lane] = cz;
  const int ri = rb * 32 + lane;
  const bool rok = !MASKED || ri < n;
  long long xi = cx, yi = cy, zi = cz;
  if (!DIAG)
    {
      xi = rok ? xs[ri] : 0;
      yi = rok ? xs[n + ri] : 0;
      zi = rok ? xs[2 * n + ri] : 0;
    }
  __syncwarp();
  double ax = 0, ay = 0, az = 0;
  double v0x = 0, v0y = 0, v0z = 0, v1x = 0, v1y = 0, v1z = 0;
  const int src = (lane + 1) & 31;
  const long long *P = T + lane;
  int ncol = MASKED ? n - cb * 32 : 32;       /* valid columns of the block (>= 32: all) */
#pragma unroll 1
  for (int k = 0; k < IT; ++k)
    {
      const long long x0 = P[OFF0], y0 = P[64 + OFF0], z0 = P[128 + OFF0];
      const long long x1 = P[OFF1], y1 = P[64 + OFF1], z1 = P[128 + OFF1];
      ++P;
      double dx0 = fx_diff(xi, x0), dy0 = fx_diff(yi, y0), dz0 = fx_diff(zi, z0);
      double dx1 = fx_diff(xi, x1), dy1 = fx_diff(yi, y1), dz1 = fx_diff(zi, z1);
      if (!CUBIC)
        {
          dx0 *= LP.s[0];
          dy0 *= LP.s[1];
          dz0 *= LP.s[2];
          dx1 *= LP.s[0];
          dy1 *= LP.s[1];
          dz1 *= LP.s[2];
        }
      const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
      const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
      double e0 = exp_tab<TS>(r20, LP, tab);
      double e1 = exp_tab<TS>(r21, LP, tab);
      bool a0 = true, a1 = true;
      if (DIAG) a1 = !((k == IT - 1) && (lane >= 16));          /* offset 16: every pair appears in two lanes, keep one */
      if (MASKED)
        {
          a0 = a0 && rok && (((lane + OFF0 + k) & 31) < ncol);
          a1 = a1 && rok && (((lane + OFF1 + k) & 31) < ncol);
          e0 = a0 ? e0 : 0.0;
        }
      if (DIAG || MASKED) e1 = a1 ? e1 : 0.0;
      S.s0 += e0;
      S.s1 = fma(e0, r20, S.s1);
      S.s0 += e1;
      S.s1 = fma(e1, r21, S.s1);
      if (VK == V_STEP)
        {
          S.cnt += (a0 && __double_as_longlong(r20) <= LP.thr_bits) ? 1 : 0;
          S.cnt += (a1 && __double_as_longlong(r21) <= LP.thr_bits) ? 1 : 0;
        }
      ax = fma(e0, dx0, ax);
      ay = fma(e0, dy0, ay);
      az = fma(e0, dz0, az);
      v0x = fma(-e0, dx0, v0x);
      v0y = fma(-e0, dy0, v0y);
      v0z = fma(-e0, dz0, v0z);
      ax = fma(e1, dx1, ax);
      ay = fma(e1, dy1, ay);
      az = fma(e1, dz1, az);
      v1x = fma(-e1, dx1, v1x);
      v1y = fma(-e1, dy1, v1y);
      v1z = fma(-e1, dz1, v1z);
      /* pass the column accumulators on: lane l+1 held the column this lane meets next */
      v0x = shfl_d(v0x, src);
      v0y = shfl_d(v0y, src);
      v0z = shfl_d(v0z, src);
      v1x = shfl_d(v1x, src);
      v1y = shfl_d(v1y, src);
      v1z = shfl_d(v1z, src);
    }
  /* the tile owns its row block and its column block for the whole round: plain read-modify-write, no atomics */
  const double gx = CUBIC ? LP.gs[0] : LP.sc, gy = CUBIC ? LP.gs[1] : LP.sc, gz = CUBIC ? LP.gs[2] : LP.sc;
  if (rok)
    {
      acc[ri] = fma(gx, ax, acc[ri]);
      acc[n + ri] = fma(gy, ay, acc[n + ri]);
      acc[2 * n + ri] = fma(gz, az, acc[2 * n + ri]);
    }
  __syncwarp();
  {
    /* after IT rotations the stream-0 visitor of lane l belongs to column l + OFF0 + IT, the stream-1 visitor to l + OFF1 + IT */
    const int q0 = (lane + OFF0 + IT) & 31, q1 = (lane + OFF1 + IT) & 31;
    const int c0 = cb * 32 + q0, c1 = cb * 32 + q1;
    if (!MASKED || c0 < n)
      {
        acc[c0] = fma(gx, v0x, acc[c0]);
        acc[n + c0] = fma(gy, v0y, acc[n + c0]);
        acc[2 * n + c0] = fma(gz, v0z, acc[2 * n + c0]);
      }
    __syncwarp();
    if (!MASKED || c1 < n)
      {
        acc[c1] = fma(gx, v1x, acc[c1]);
        acc[n + c1] = fma(gy, v1y, acc[n + c1]);
        acc[2 * n + c1] = fma(gz, v1z, acc[2 * n + c1]);
      }
    __syncwarp();
  }
}

// Two row blocks against one column block ("2 x 1 register tile"): lane l owns the row particles r0 = 64 m + l and r1 = r0 + 32
// and meets column (l + k) mod 32 with BOTH, so one column load and one visitor rotation serve two pairs (3 LDS + 6 SHFL per
// two pairs; lean_tile: 6 + 12).  The row gradients stay in registers over both column blocks of the 64-block m'.
// Full blocks only (the caller falls back to lean_tile on ragged edges).
template <int VK, bool CUBIC, int TS>
QMC_DEVICE void lean_pairblock(const LeanParams &LP, const double *tab, const long long *xs, const int n, const int m, const int mp,
                               long long *T, const int lane, double *acc, LeanSums &S)
{
  const int r0 = m * 64 + lane, r1 = r0 + 32;
  const long long x0 = xs[r0], y0 = xs[n + r0], z0 = xs[2 * n + r0];
  const long long x1 = xs[r1], y1 = xs[n + r1], z1 = xs[2 * n + r1];
  double a0x = 0, a0y = 0, a0z = 0, a1x = 0, a1y = 0, a1z = 0;
  const int src = (lane + 1) & 31;
  const double gx = CUBIC ? LP.gs[0] : LP.sc, gy = CUBIC ? LP.gs[1] : LP.sc, gz = CUBIC ? LP.gs[2] : LP.sc;
#pragma unroll 1
  for (int h = 0; h < 2; ++h)
    {
      const int cb = mp * 2 + h;
      const int ci = cb * 32 + lane;
      const long long cx = xs[ci], cy = xs[n + ci], cz = xs[2 * n + ci];
      __syncwarp();
      T[lane] = cx;
      T[32 + lane] = cx;
      T[64 + lane] = cy;
      T[96 + lane] = cy;
      T[128 + lane] = cz;
      T[160 + lane] = cz;
      __syncwarp();
      double vx = 0, vy = 0, vz = 0;
      const long long *P = T + lane;
#pragma unroll PAIR_UNROLL
      for (int k = 0; k < 32; ++k)
        {
          const long long xc = P[0], yc = P[64], zc = P[128];
          ++P;
          double dx0 = fx_diff(x0, xc), dy0 = fx_diff(y0, yc), dz0 = fx_diff(z0, zc);
          double dx1 = fx_diff(x1, xc), dy1 = fx_diff(y1, yc), dz1 = fx_diff(z1, zc);
          if (!CUBIC)
            {
              dx0 *= LP.s[0];
              dy0 *= LP.s[1];
              dz0 *= LP.s[2];
              dx1 *= LP.s[0];
              dy1 *= LP.s[1];
              dz1 *= LP.s[2];
            }
          const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
          const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
          const double e0 = exp_tab<TS>(r20, LP, tab);
          const double e1 = exp_tab<TS>(r21, LP, tab);
          S.s0 += e0;
          S.s1 = fma(e0, r20, S.s1);
          S.s0 += e1;
          S.s1 = fma(e1, r21, S.s1);
          if (VK == V_STEP)
            {
              S.cnt += (__double_as_longlong(r20) <= LP.thr_bits) ? 1 : 0;
              S.cnt += (__double_as_longlong(r21) <= LP.thr_bits) ? 1 : 0;
            }
          a0x = fma(e0, dx0, a0x);
          a0y = fma(e0, dy0, a0y);
          a0z = fma(e0, dz0, a0z);
          a1x = fma(e1, dx1, a1x);
          a1y = fma(e1, dy1, a1y);
          a1z = fma(e1, dz1, a1z);
          vx = fma(-e0, dx0, vx);
          vy = fma(-e0, dy0, vy);
          vz = fma(-e0, dz0, vz);
          vx = fma(-e1, dx1, vx);
          vy = fma(-e1, dy1, vy);
          vz = fma(-e1, dz1, vz);
          vx = shfl_d(vx, src);
          vy = shfl_d(vy, src);
          vz = shfl_d(vz, src);
        }
      /* after 32 rotations the visitor is back at the lane whose index is its column */
      acc[ci] = fma(gx, vx, acc[ci]);
      acc[n + ci] = fma(gy, vy, acc[n + ci]);
      acc[2 * n + ci] = fma(gz, vz, acc[2 * n + ci]);
    }
  acc[r0] = fma(gx, a0x, acc[r0]);
  acc[n + r0] = fma(gy, a0y, acc[n + r0]);
  acc[2 * n + r0] = fma(gz, a0z, acc[2 * n + r0]);
  acc[r1] = fma(gx, a1x, acc[r1]);
  acc[n + r1] = fma(gy, a1y, acc[n + r1]);
  acc[2 * n + r1] = fma(gz, a1z, acc[2 * n + r1]);
  __syncwarp();
}

struct LeanScratch {
  double *R;        // [dofs] configuration being evaluated
  double *acc;      // [dofs] gradient accumulator (single copy)
  long long *X;     // [dofs] the configuration as 64-bit fractions of the box
  double *red;      // [nwarps*4 + 8]
  long long *strip; // this warp's column strip
};

QMC_DEVICE LeanScratch make_lean_scratch(const Plan &plan, unsigned char *smem, int group_index, int nwarps, int gwarp, size_t group_bytes)
{
  LeanScratch S;
  double *base = reinterpret_cast<double *>(smem + (size_t)group_index * group_bytes);
  S.R = base;
  S.acc = base + plan.dofs;
  S.X = reinterpret_cast<long long *>(S.acc + plan.dofs);
  S.red = S.acc + 2 * (size_t)plan.dofs;
  S.strip = reinterpret_cast<long long *>(S.red + (nwarps * 4 + 8)) + (size_t)gwarp * STRIP;
  return S;
}

// log psi, laplacian, V of the configuration in S.R; S.acc = grad log psi on return (all threads of the group).
// Schedule: the particles are cut in 64-blocks; the pairs of DIFFERENT 64-blocks are visited in the rounds of a round-robin
// tournament (in one round every 64-block belongs to at most one warp, which therefore owns its gradient rows), then every
// 64-block is paired with itself (two diagonal 32-tiles + one 32 x 32 tile).
template <int VK, bool CUBIC, int TS>
QMC_DEVICE EvalResult lean_evaluate(const Plan &plan, const LeanParams &LP, const double *tab, const LeanScratch &S, const Group &G)
{
  const int n = plan.N[0], dofs = plan.dofs;
  for (int k = G.gtid; k < dofs; k += G.gsize)
    {
      S.acc[k] = 0.0;
      const int d = k / n;                          /* walker layout: x[n] y[n] z[n] */
      S.X[k] = to_box_fraction(S.R[k], plan.Linv[d]);
    }
  G.sync();
  LeanSums sums = {0.0, 0.0, 0};
  const int nb = (n + 31) >> 5;                   /* 32-blocks */
  const int nB = (n + 63) >> 6;                   /* 64-blocks */
  const int nt = nB + (nB & 1);                   /* teams of the tournament (an odd count gets a phantom block) */
  const int m = nt - 1;
  for (int r = 0; r < m; ++r)
    {
      for (int i = G.gwarp; i < nt / 2; i += G.nwarps)
        {
          int a, b;
          if (i == 0)
            {
              a = m;
              b = r;
            }
          else
            {
              a = r + i;
              if (a >= m) a -= m;
              b = r - i;
              if (b < 0) b += m;
            }
          if (a >= nB || b >= nB) continue;        /* bye */
          const int M = a > b ? a : b, Mp = a > b ? b : a;
          if (M * 64 + 64 <= n) lean_pairblock<VK, CUBIC, TS>(LP, tab, S.X, n, M, Mp, S.strip, G.lane, S.acc, sums);
          else
            { /* ragged last 64-block: its one or two 32-blocks against the two (full) 32-blocks of Mp */
              for (int rb = 2 * M; rb < nb; ++rb)
                for (int cb = 2 * Mp; cb < 2 * Mp + 2; ++cb)
                  {
                    if (rb * 32 + 32 <= n) lean_tile<VK, CUBIC, false, false, TS>(LP, tab, S.X, n, rb, cb, S.strip, G.lane, S.acc, sums);
                    else lean_tile<VK, CUBIC, false, true, TS>(LP, tab, S.X, n, rb, cb, S.strip, G.lane, S.acc, sums);
                  }
            }
        }
      G.sync();
    }
  for (int d = G.gwarp; d < nB; d += G.nwarps)
    {
      const int b0 = 2 * d, b1 = 2 * d + 1;
      if (b0 * 32 + 32 <= n) lean_tile<VK, CUBIC, true, false, TS>(LP, tab, S.X, n, b0, b0, S.strip, G.lane, S.acc, sums);
      else lean_tile<VK, CUBIC, true, true, TS>(LP, tab, S.X, n, b0, b0, S.strip, G.lane, S.acc, sums);
      if (b1 < nb)
        {
          if (b1 * 32 + 32 <= n)
            {
              lean_tile<VK, CUBIC, true, false, TS>(LP, tab, S.X, n, b1, b1, S.strip, G.lane, S.acc, sums);
              lean_tile<VK, CUBIC, false, false, TS>(LP, tab, S.X, n, b1, b0, S.strip, G.lane, S.acc, sums);
            }
          else
            {
              lean_tile<VK, CUBIC, true, true, TS>(LP, tab, S.X, n, b1, b1, S.strip, G.lane, S.acc, sums);
              lean_tile<VK, CUBIC, false, true, TS>(LP, tab, S.X, n, b1, b0, S.strip, G.lane, S.acc, sums);
            }
        }
    }
  double lp = LP.u0 * sums.s0;
  double lap = fma(LP.lap1, sums.s1, LP.lap0 * sums.s0);
  double V = (VK == V_STEP) ? LP.v_val * (double)sums.cnt : 0.0;
  lp = warp_sum(lp);
  lap = warp_sum(lap);
  V = warp_sum(V);
  if (G.lane == 0)
    {
      double *rr = S.red + G.gwarp * 4;
      rr[0] = lp;
      rr[1] = lap;
      rr[2] = V;
    }
  G.sync();
  EvalResult out;
  out.logpsi = 0;
  out.lap = 0;
  out.V = 0;
  for (int w = 0; w < G.nwarps; ++w)
    {
      out.logpsi += S.red[w * 4 + 0];
      out.lap += S.red[w * 4 + 1];
      out.V += S.red[w * 4 + 2];
    }
  out.valid = 1;
  G.sync();
  return out;
}

template <int VK, int THREADS, int MINB, int TS>
__global__ void __launch_bounds__(THREADS, MINB)
pair_move_kernel(const __grid_constant__ Plan plan, const __grid_constant__ DmcParams dmc, const __grid_constant__ LeanParams LP, const StepArgs a,
                 double *e_old, int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  double *tab0 = reinterpret_cast<double *>(smem);       /* 2^(j/256), TS interleaved copies, at the start of the CTA's shared memory */
  const LeanScratch S = make_lean_scratch(plan, smem + (size_t)256 * TS * sizeof(double), g, nwarps_per_group, G.gwarp, group_bytes);
  for (int i = threadIdx.x; i < 256 * TS; i += THREADS) tab0[i] = LP.tab[i / TS];
  __syncthreads();
  const double *tab = tab0 + (threadIdx.x & (TS - 1));   /* this lane's copy */
  const int dofs = plan.dofs;
  const Ctl *ctl = a.ctl;
  const int j0 = a.count < 0 ? 0 : (a.first < 0 ? ctl->n_cur + a.first + 1 : a.first);
  const int j1 = a.count < 0 ? ctl->n_cur : j0 + a.count;
  const unsigned long long step = ctl->step;
  const uint32_t step_lo = (uint32_t)step, step_hi = (uint32_t)(step >> 32) | ((uint32_t)dmc.rank << 24);

  for (int j = j0 + blockIdx.x * groups + g; j < j1; j += gridDim.x * groups)
    {
      const int src = a.parent[j];
      const double *Rold = a.in.R + (size_t)src * dofs;
      const double *Fold = a.in.F + (size_t)src * dofs;
      // ---- drift + diffusion (dmcMoves.cpp:23-26, vmcMoves.cpp:25-33): 4 coordinates per Philox call
      for (int c = G.gtid; c * 4 < dofs; c += G.gsize)
        {
          u32x4 rnd = philox4x32_10({(uint32_t)j, (uint32_t)c, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
          double z[4];
          normal_pair(rnd.x, rnd.y, z[0], z[1]);
          normal_pair(rnd.z, rnd.w, z[2], z[3]);
#pragma unroll
          for (int m = 0; m < 4; ++m)
            {
              const int k = c * 4 + m;
              if (k < dofs)
                {
                  S.R[k] = fma(dmc.sqrt_tau, z[m], fma(dmc.tau, Fold[k], Rold[k]));
                  if (a.trace.xi) a.trace.xi[(size_t)j * dofs + k] = z[m];
                }
            }
        }
      G.sync();
      const EvalResult r = LP.cubic ? lean_evaluate<VK, true, TS>(plan, LP, tab, S, G) : lean_evaluate<VK, false, TS>(plan, LP, tab, S, G);
      // ---- log Green ratio and |grad|^2
      double q = 0, f2 = 0;
      for (int k = G.gtid; k < dofs; k += G.gsize)
        {
          const double fo = Fold[k], ro = Rold[k], fn = S.acc[k], rn = S.R[k];
          const double aa = 0.5 * dmc.tau * (fo - fn) - rn + ro;      /* dmcMoves.cpp:46-51 */
          q = fma(aa, fo + fn, q);
          f2 = fma(fn, fn, f2);
        }
      q = warp_sum(q);
      f2 = warp_sum(f2);
      if (G.lane == 0)
        {
          S.red[G.gwarp * 4 + 0] = q;
          S.red[G.gwarp * 4 + 1] = f2;
        }
      G.sync();
      double qs = 0, f2s = 0;
      for (int w = 0; w < G.nwarps; ++w)
        {
          qs += S.red[w * 4 + 0];
          f2s += S.red[w * 4 + 1];
        }
      const double Enew = r.V - 0.5 * (r.lap + f2s);                  /* energy.cpp:6-35 */
      const double lp_old = a.in.logpsi[src];
      const double Eold = a.in.E[src];
      // ---- Metropolis (dmcDriver.cpp:27-38, metropolis.cpp:5-29); every thread computes the same draw
      const u32x4 rd = philox4x32_10({(uint32_t)j, 0xffffffffu, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
      const double u_m = u01_53(rd.x, rd.y);
      const double A = 2.0 * (r.logpsi - lp_old) + (dmc.vmc ? 0.0 : qs);
      const bool accepted = A > 0.0 || A > log(u_m);
      const double Efin = accepted ? Enew : Eold;
      // ---- commit (accept) or carry the parent through (reject: wNew = wOld, dmcDriver.cpp:73-76)
      double *Rout = a.out.R + (size_t)j * dofs;
      double *Fout = a.out.F + (size_t)j * dofs;
      if (accepted)
        for (int k = G.gtid; k < dofs; k += G.gsize)
          {
            Rout[k] = S.R[k];
            Fout[k] = S.acc[k];
          }
      else
        for (int k = G.gtid; k < dofs; k += G.gsize)
          {
            Rout[k] = Rold[k];
            Fout[k] = Fold[k];
          }
      if (G.gtid == 0)
        {
          a.out.logpsi[j] = accepted ? r.logpsi : lp_old;
          a.out.E[j] = Efin;
          e_old[j] = Eold;            /* branching needs (E' + E)/2 (branching.cpp:13); decided in scan_v2 */
          a.flags[j] = accepted ? 1 : 0;
        }
      if (a.trace.Rprop)
        {
          for (int k = G.gtid; k < dofs; k += G.gsize)
            {
              a.trace.Rprop[(size_t)j * dofs + k] = S.R[k];
              a.trace.Fprop[(size_t)j * dofs + k] = S.acc[k];
            }
          if (G.gtid == 0)
            {
              a.trace.logpsi[j] = r.logpsi;
              a.trace.lap[j] = r.lap;
              a.trace.V[j] = r.V;
              a.trace.Eprop[j] = Enew;
              a.trace.q[j] = qs;
              a.trace.u_m[j] = u_m;
              a.trace.accepted[j] = accepted;
              a.trace.valid[j] = 1;
              a.trace.parent[j] = src;
            }
        }
      G.sync();
    }
}

#ifdef QMC_EMULATE   /* tests/emu: cudaLaunchKernel gets a trampoline that unpacks the argument array */
template <int VK, int THREADS, int MINB, int TS> void pair_tramp(void **p)
{
  pair_move_kernel<VK, THREADS, MINB, TS>(*(Plan *)p[0], *(DmcParams *)p[1], *(LeanParams *)p[2], *(StepArgs *)p[3], *(double **)p[4], *(int *)p[5],
                                      *(size_t *)p[6]);
}
template <int VK, int THREADS, int MINB, int TS> const void *pair_fn() { return (const void *)&pair_tramp<VK, THREADS, MINB, TS>; }
#else
template <int VK, int THREADS, int MINB, int TS> const void *pair_fn() { return (const void *)pair_move_kernel<VK, THREADS, MINB, TS>; }
#endif

// launch geometries compiled: variant 0 = 256 threads x 2 CTAs/SM (<= 128 registers), 1 = 128 x 4, 2 = 128 x 5 (<= 96 registers),
// 3 = 512 x 1 with 16 interleaved copies of the exp table (32 KB, shared by the 4 walker groups of the SM's single CTA: two CTAs could not
// afford it) -- the same 16 warps per SM as variant 0, conflict-free table gathers
template <int VK> const void *pair_variant(int v)
{
  return v == 3 ? pair_fn<VK, 512, 1, 16>() : v == 2 ? pair_fn<VK, 128, 5, 1>() : v == 1 ? pair_fn<VK, 128, 4, 1>() : pair_fn<VK, 256, 2, 1>();
}
const void *pair_ptr(int vk, int v) { return vk == V_STEP ? pair_variant<V_STEP>(v) : pair_variant<V_NONE>(v); }


// ---------------------------------------------------------------------------------------------------------------------------
// g(r) of the one-species shapes over the post-branching population (pairCorrelation::accumulate, pairCorrelation.cpp:17-39,
// over the population dmcDriver::accumulate iterates, dmcDriver.cpp:257-264: unit weight per walker).  Distances come from the
// same fixed-point coordinates as the move kernel (wrapping integer subtraction, 3 FP64 per r^2 instead of 12 + 3 FRND), one warp
// per walker, no gradients so no visitors and no shuffles, four pairs in flight per lane.  c/r^2 is quantised to 64-bit fixed point
// and added to the CTA's shared-memory histogram with two native 32-bit atomics (hist_add_fixed), one flush per CTA per launch.
struct HistParams {
  double minx, maxx, dxinv;      // in box units (x / s) when the box is cubic, else in real units
  double fxs;                    // fixed-point scale of 1/r2: a pair at r = maxx maps to 2^36
  double flush;                  // (norm / s^2) / fxs : fixed-point total -> histogram sum
  double direct;                 // norm / s^2         : weight of the exact fp64 path (contributions >= 2^40)
  double s[3];                   // L_d 2^-64
  int bins, offset, stride, cubic;
};

QMC_DEVICE void lean_hist_pair(const double r2, const bool ok, const HistParams &H, unsigned *fx, double *replica)
{
  const double y = rsqrt_fast(r2);
  const double rr = r2 * y;
  const int b = (int)((rr - H.minx) * H.dxinv);
  const double t = y * y * H.fxs;
  if (ok && rr < H.maxx && b >= 0)
    {
      const int bb = min(b, H.bins - 1);
      if (t < HIST_FX_BYPASS) hist_add_fixed(fx + H.offset, H.stride, bb, (long long)t);
      else atomicAdd(replica + H.offset + bb, H.direct * y * y);
    }
}

template <bool CUBIC>
__global__ void __launch_bounds__(CTA_THREADS, 4)
pair_hist_kernel(const __grid_constant__ Plan plan, const __grid_constant__ HistParams H, const WalkerBuf buf, const int *parent, const Ctl *ctl,
                 double *hist)
{
  QMC_DYNAMIC_SMEM(smem);
  const int n = plan.N[0], dofs = plan.dofs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned *fx = reinterpret_cast<unsigned *>(smem);                                  /* lo[stride], hi[stride] */
  long long *X = reinterpret_cast<long long *>(smem + (size_t)2 * H.stride * sizeof(unsigned)) + (size_t)warp * dofs;
  double *replica = hist + (size_t)blockIdx.x * H.stride;
  for (int i = threadIdx.x; i < 2 * H.stride; i += CTA_THREADS) fx[i] = 0u;
  __syncthreads();
  const int nW = ctl->n_cur, nb = (n + 31) >> 5;
  for (int j = blockIdx.x * (CTA_THREADS / 32) + warp; j < nW; j += gridDim.x * (CTA_THREADS / 32))
    {
      const double *R = buf.R + (size_t)parent[j] * dofs;
      __syncwarp();
      for (int k = lane; k < dofs; k += 32) X[k] = to_box_fraction(R[k], plan.Linv[k / n]);
      __syncwarp();
      for (int b = 0; b < nb; ++b)
        {
          const int ri = b * 32 + lane;
          const bool rok = ri < n;
          const long long xi = rok ? X[ri] : 0, yi = rok ? X[n + ri] : 0, zi = rok ? X[2 * n + ri] : 0;
          /* full column blocks c < b: 32 columns each, lane l meets column (l + k) & 31 */
          for (int c = 0; c < b; ++c)
            {
              const long long *xc = X + c * 32, *yc = xc + n, *zc = yc + n;
#pragma unroll 4
              for (int k = 0; k < 32; ++k)
                {
                  const int q = (lane + k) & 31;
                  double dx = fx_diff(xi, xc[q]), dy = fx_diff(yi, yc[q]), dz = fx_diff(zi, zc[q]);
                  if (!CUBIC)
                    {
                      dx *= H.s[0];
                      dy *= H.s[1];
                      dz *= H.s[2];
                    }
                  lean_hist_pair(fma(dz, dz, fma(dy, dy, dx * dx)), rok, H, fx, replica);
                }
            }
          /* the block with itself: cyclic offsets 1..nc/2 (offset nc/2 half-populated when nc is even) */
          {
            const int nc = min(32, n - b * 32), K = nc >> 1;
            const long long *xc = X + b * 32, *yc = xc + n, *zc = yc + n;
#pragma unroll 4
            for (int k = 1; k <= K; ++k)
              {
                int q = lane + k;
                if (q >= nc) q -= nc;
                const bool ok = rok && !((2 * k == nc) && (lane >= K));
                const int qq = ok ? q : 0;
                double dx = fx_diff(xi, xc[qq]), dy = fx_diff(yi, yc[qq]), dz = fx_diff(zi, zc[qq]);
                if (!CUBIC)
                  {
                    dx *= H.s[0];
                    dy *= H.s[1];
                    dz *= H.s[2];
                  }
                double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                r2 = ok ? r2 : 1.0;              /* idle lanes: keep rsqrt away from 0 */
                lean_hist_pair(r2, ok, H, fx, replica);
              }
          }
        }
    }
  __syncthreads();
  for (int b = threadIdx.x; b < H.bins; b += CTA_THREADS)
    {
      const unsigned lo = fx[H.offset + b], hi = fx[H.stride + H.offset + b];
      const long long t = (long long)(((unsigned long long)hi << 32) | (unsigned long long)lo);
      if (t != 0) atomicAdd(replica + H.offset + b, (double)t * H.flush);
    }
}

#ifdef QMC_EMULATE
template <bool CUBIC> void hist_tramp(void **p)
{
  pair_hist_kernel<CUBIC>(*(Plan *)p[0], *(HistParams *)p[1], *(WalkerBuf *)p[2], *(const int **)p[3], *(const Ctl **)p[4], *(double **)p[5]);
}
template <bool CUBIC> const void *hist_fn() { return (const void *)&hist_tramp<CUBIC>; }
#else
template <bool CUBIC> const void *hist_fn() { return (const void *)pair_hist_kernel<CUBIC>; }
#endif

} // namespace

// warps per walker group of the tournament schedule: one warp per pair of 64-blocks of a round, a power of two <= 8
int pair_warps_per_walker(int n)
{
  const int nB = (n + 63) / 64, half = nB / 2;
  int w = 1;
  while (w * 2 <= half && w < 8) w *= 2;
  return w;
}

size_t pair_group_bytes(const Plan &plan, int nwarps)
{
  size_t d = (size_t)3 * plan.dofs + (nwarps * 4 + 8) + (size_t)nwarps * STRIP;
  d = (d + 1) & ~(size_t)1;
  return d * 8;
}
size_t pair_cta_extra_bytes(int variant) { return (size_t)256 * sizeof(double) * (variant == 3 ? 16 : 1); }     /* the exp table(s) */

// the 256 table values 2^(j/256) (host, correctly rounded from long double)
void pair_exp_table(double *tab256)
{
  for (int j = 0; j < 256; ++j) tab256[j] = (double)exp2l((long double)j / 256.0L);
}


size_t pair_hist_smem(const Plan &plan) { return (size_t)2 * plan.hist_stride * sizeof(unsigned) + (size_t)(CTA_THREADS / 32) * plan.dofs * sizeof(long long); }

int pair_hist_occupancy(const Plan &plan)
{
  const bool cubic = plan.L[0] == plan.L[1] && plan.L[1] == plan.L[2];
  int nblk = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, cubic ? hist_fn<true>() : hist_fn<false>(), CTA_THREADS, pair_hist_smem(plan)) != cudaSuccess) return 1;
  return nblk < 1 ? 1 : nblk;
}

// the pairCorrelation estimator `e` of the plan through the fixed-point kernel (hist = the replicas [grid][hist_stride])
cudaError_t launch_pair_hist(int grid, cudaStream_t st, const Plan &plan, int e, const WalkerBuf &buf, const int *parent, const Ctl *ctl, double *hist)
{
  const Estimator &E = plan.est[e];
  const long double two64 = 18446744073709551616.0L;
  HistParams H;
  H.cubic = (plan.L[0] == plan.L[1] && plan.L[1] == plan.L[2]) ? 1 : 0;
  for (int d = 0; d < 3; ++d) H.s[d] = (double)((long double)plan.L[d] / two64);
  const long double s = H.cubic ? (long double)plan.L[0] / two64 : 1.0L;      /* length unit of the pair loop */
  H.minx = (double)((long double)E.minx / s);
  H.maxx = (double)((long double)E.maxx / s);
  H.dxinv = (double)((long double)E.dxinv * s);
  H.fxs = (double)(68719476736.0L * ((long double)E.maxx / s) * ((long double)E.maxx / s));      /* 2^36 at r = maxx */
  H.direct = (double)((long double)E.norm / (s * s));
  H.flush = (double)((long double)E.norm / (s * s) / (long double)H.fxs);
  H.bins = E.bins;
  H.offset = E.offset;
  H.stride = plan.hist_stride;
  const WalkerBuf b = buf;
  void *args[] = {(void *)&plan, (void *)&H, (void *)&b, (void *)&parent, (void *)&ctl, (void *)&hist};
  return cudaLaunchKernel(H.cubic ? hist_fn<true>() : hist_fn<false>(), dim3(grid), dim3(CTA_THREADS), args, pair_hist_smem(plan), st);
}

cudaError_t configure_kernels_pair(int max_smem_optin)
{
  for (int cub = 0; cub < 2; ++cub)
    {
      cudaError_t e = cudaFuncSetAttribute(cub ? hist_fn<true>() : hist_fn<false>(), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
      if (e != cudaSuccess) return e;
    }
  for (int vk : {(int)V_STEP, (int)V_NONE})
    for (int occ = 0; occ <= 3; ++occ)
      {
        cudaError_t e = cudaFuncSetAttribute(pair_ptr(vk, occ), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
        if (e != cudaSuccess) return e;
      }
  return cudaSuccess;
}

int pair_variant_threads(int v) { return v == 3 ? 512 : v == 0 ? 256 : 128; }

int pair_move_occupancy(int vk, int occ, size_t smem)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pair_ptr(vk, occ), pair_variant_threads(occ), smem) != cudaSuccess) return 1;
  return n < 1 ? 1 : n;
}

int pair_move_regs(int vk, int occ)
{
  cudaFuncAttributes at;
  if (cudaFuncGetAttributes(&at, pair_ptr(vk, occ)) != cudaSuccess) return -1;
  return at.numRegs;
}

cudaError_t launch_pair_move(int vk, int occ, const LaunchCfg &cfg, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc,
                             const double *dev_tab, const StepArgs &a, double *e_old)
{
  int nw = cfg.nwarps_per_group;
  size_t gb = (cfg.smem - pair_cta_extra_bytes(occ)) / cfg.groups_per_cta;
  /* constants of the pair loop, folded on the host (jastrowGaussGauss.h:19-21: u = u0 exp(-alpha r^2), alpha = 1/(2 r0^2)) */
  const PairPass &PP = plan.p2[0];
  const long double al = PP.J.p[0], ln2 = 0.693147180559945309417232121458176568L;
  const long double two64 = 18446744073709551616.0L;
  LeanParams LP;
  LP.cubic = (plan.L[0] == plan.L[1] && plan.L[1] == plan.L[2]) ? 1 : 0;
  if (const char *e = getenv("QMC4_LEAN_NONCUBIC"))      /* test hook: run a cubic box through the per-dimension path */
    if (atoi(e)) LP.cubic = 0;
  for (int d = 0; d < 3; ++d) LP.s[d] = (double)((long double)plan.L[d] / two64);
  /* cubic box: the pair sums are taken over d' = d / s (box units); every constant absorbs the factor s^2 */
  const long double s2 = LP.cubic ? ((long double)plan.L[0] / two64) * ((long double)plan.L[0] / two64) : 1.0L;
  const long double a2 = al * s2;                    /* exp(-alpha r2) = exp(-a2 r2') */
  LP.k = (double)(-a2 * 256.0L / ln2);
  LP.c = (double)(ln2 / (256.0L * a2));
  LP.c1 = (double)(-a2);
  LP.c2 = (double)(a2 * a2 / 2.0L);
  LP.c3 = (double)(-a2 * a2 * a2 / 6.0L);
  LP.c4 = (double)(a2 * a2 * a2 * a2 / 24.0L);
  LP.sc = -2.0 * PP.J.p[1] * PP.J.p[0];
  for (int d = 0; d < 3; ++d) LP.gs[d] = (double)((long double)LP.sc * ((long double)plan.L[d] / two64));
  LP.u0 = PP.J.p[1];
  LP.lap0 = -12.0 * PP.J.p[0] * PP.J.p[1];
  LP.lap1 = (double)(8.0L * al * al * (long double)PP.J.p[1] * s2);
  LP.v_val = PP.v_val;
  {
    const double thr = (double)((long double)PP.v_thr2 / s2);       /* r2' <= thr  <=>  r2 <= v_thr2 (up to the last bit) */
    memcpy(&LP.thr_bits, &thr, 8);
  }
  LP.tab = dev_tab;
  void *args[] = {(void *)&plan, (void *)&dmc, (void *)&LP, (void *)&a, (void *)&e_old, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(pair_ptr(vk, occ), dim3(grid), dim3(cfg.block), args, cfg.smem, st);
}

} // namespace qmc
