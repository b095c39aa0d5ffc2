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
//
// Reference semantics: dmcDriver::update (dmcDriver.cpp:56-84) = driftDiffusionFirstOrder::move (dmcMoves.cpp:12-30) +
// tableDistances::update (tableDistances.cpp:81-108) + jastrowTwoBodyWavefunctionUndistinguishable::accumulateDerivatives
// (jastrowWavefunctionTwoBody.h:124-164) + potentialTwoBody (potentialTwoBody.h:23-33) + energy (energy.cpp:6-35) +
// transitionProbabilityRatio (dmcMoves.cpp:32-57) + metropolis::acceptLog (metropolis.cpp:5-29).
#include <cmath>
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
};

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
  const double e = tab[n & 255] * q;
  const int hi = __double2hiint(e) + ((n >> 8) << 20);
  return __hiloint2double(hi, __double2loint(e));
}

struct LeanSums {
  double s0, s1;      // sum e, sum e*r2 over this thread's pairs (e = exp(-alpha r2))
  int cnt;            // pairs inside the step potential
};

constexpr int STRIP = 192;     // doubles per warp: x[64] y[64] z[64], every column twice

// One tile: rows = block rb (lane l owns row particle rb*32+l), columns = block cb.  DIAG: rb == cb, cyclic offsets 1..16
// (offset 16 half-populated).  MASKED: the row and/or column block is the ragged last block of the species.
// Two independent pairs per lane and step: stream 0 meets column (l + OFF0 + k), stream 1 column (l + OFF1 + k).
template <int VK, bool DIAG, bool MASKED>
QMC_DEVICE void lean_tile(const LeanParams &LP, const double *tab, const Box &box, const double *xs, const int n, const int rb,
                          const int cb, double *T, const int lane, double *acc, LeanSums &S)
{
  constexpr int IT = DIAG ? 8 : 16, OFF0 = DIAG ? 1 : 0, OFF1 = DIAG ? 9 : 16;
  const int ci = cb * 32 + lane;
  const bool cok = !MASKED || ci < n;
  const double cx = cok ? xs[ci] : 0.0, cy = cok ? xs[n + ci] : 0.0, cz = cok ? xs[2 * n + ci] : 0.0;
  __syncwarp();                      /* the strip may still be read by a slower lane of the previous tile */
  T[lane] = cx;
  T[32 + lane] = cx;
  T[64 + lane] = cy;
  T[96 + lane] = cy;
  T[128 + lane] = cz;
  T[160 + lane] = cz;
  const int ri = rb * 32 + lane;
  const bool rok = !MASKED || ri < n;
  double xi = cx, yi = cy, zi = cz;
  if (!DIAG)
    {
      xi = rok ? xs[ri] : 0.0;
      yi = rok ? xs[n + ri] : 0.0;
      zi = rok ? xs[2 * n + ri] : 0.0;
    }
  __syncwarp();
  double ax = 0, ay = 0, az = 0;
  double v0x = 0, v0y = 0, v0z = 0, v1x = 0, v1y = 0, v1z = 0;
  const int src = (lane + 1) & 31;
  const double *P = T + lane;
  int ncol = MASKED ? n - cb * 32 : 32;       /* valid columns of the block (>= 32: all) */
#pragma unroll 1
  for (int k = 0; k < IT; ++k)
    {
      const double x0 = P[OFF0], y0 = P[64 + OFF0], z0 = P[128 + OFF0];
      const double x1 = P[OFF1], y1 = P[64 + OFF1], z1 = P[128 + OFF1];
      ++P;
      const double dx0 = min_image(xi - x0, box.Lx, box.ix);
      const double dy0 = min_image(yi - y0, box.Ly, box.iy);
      const double dz0 = min_image(zi - z0, box.Lz, box.iz);
      const double dx1 = min_image(xi - x1, box.Lx, box.ix);
      const double dy1 = min_image(yi - y1, box.Ly, box.iy);
      const double dz1 = min_image(zi - z1, box.Lz, box.iz);
      const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
      const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
      double e0 = exp_tab(r20, LP, tab);
      double e1 = exp_tab(r21, LP, tab);
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
  const double sc = LP.sc;
  if (rok)
    {
      acc[ri] = fma(sc, ax, acc[ri]);
      acc[n + ri] = fma(sc, ay, acc[n + ri]);
      acc[2 * n + ri] = fma(sc, az, acc[2 * n + ri]);
    }
  __syncwarp();
  {
    /* after IT rotations the stream-0 visitor of lane l belongs to column l + OFF0 + IT, the stream-1 visitor to l + OFF1 + IT */
    const int q0 = (lane + OFF0 + IT) & 31, q1 = (lane + OFF1 + IT) & 31;
    const int c0 = cb * 32 + q0, c1 = cb * 32 + q1;
    if (!MASKED || c0 < n)
      {
        acc[c0] = fma(sc, v0x, acc[c0]);
        acc[n + c0] = fma(sc, v0y, acc[n + c0]);
        acc[2 * n + c0] = fma(sc, v0z, acc[2 * n + c0]);
      }
    __syncwarp();
    if (!MASKED || c1 < n)
      {
        acc[c1] = fma(sc, v1x, acc[c1]);
        acc[n + c1] = fma(sc, v1y, acc[n + c1]);
        acc[2 * n + c1] = fma(sc, v1z, acc[2 * n + c1]);
      }
    __syncwarp();
  }
}

// Two row blocks against one column block ("2 x 1 register tile"): lane l owns the row particles r0 = 64 m + l and r1 = r0 + 32
// and meets column (l + k) mod 32 with BOTH, so one column load and one visitor rotation serve two pairs (3 LDS + 6 SHFL per
// two pairs; lean_tile: 6 + 12).  The row gradients stay in registers over both column blocks of the 64-block m'.
// Full blocks only (the caller falls back to lean_tile on ragged edges).
template <int VK>
QMC_DEVICE void lean_pairblock(const LeanParams &LP, const double *tab, const Box &box, const double *xs, const int n, const int m, const int mp,
                               double *T, const int lane, double *acc, LeanSums &S)
{
  const int r0 = m * 64 + lane, r1 = r0 + 32;
  const double x0 = xs[r0], y0 = xs[n + r0], z0 = xs[2 * n + r0];
  const double x1 = xs[r1], y1 = xs[n + r1], z1 = xs[2 * n + r1];
  double a0x = 0, a0y = 0, a0z = 0, a1x = 0, a1y = 0, a1z = 0;
  const int src = (lane + 1) & 31;
  const double sc = LP.sc;
#pragma unroll 1
  for (int h = 0; h < 2; ++h)
    {
      const int cb = mp * 2 + h;
      const int ci = cb * 32 + lane;
      const double cx = xs[ci], cy = xs[n + ci], cz = xs[2 * n + ci];
      __syncwarp();
      T[lane] = cx;
      T[32 + lane] = cx;
      T[64 + lane] = cy;
      T[96 + lane] = cy;
      T[128 + lane] = cz;
      T[160 + lane] = cz;
      __syncwarp();
      double vx = 0, vy = 0, vz = 0;
      const double *P = T + lane;
#pragma unroll 1
      for (int k = 0; k < 32; ++k)
        {
          const double xc = P[0], yc = P[64], zc = P[128];
          ++P;
          const double dx0 = min_image(x0 - xc, box.Lx, box.ix);
          const double dy0 = min_image(y0 - yc, box.Ly, box.iy);
          const double dz0 = min_image(z0 - zc, box.Lz, box.iz);
          const double dx1 = min_image(x1 - xc, box.Lx, box.ix);
          const double dy1 = min_image(y1 - yc, box.Ly, box.iy);
          const double dz1 = min_image(z1 - zc, box.Lz, box.iz);
          const double r20 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
          const double r21 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
          const double e0 = exp_tab(r20, LP, tab);
          const double e1 = exp_tab(r21, LP, tab);
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
      acc[ci] = fma(sc, vx, acc[ci]);
      acc[n + ci] = fma(sc, vy, acc[n + ci]);
      acc[2 * n + ci] = fma(sc, vz, acc[2 * n + ci]);
    }
  acc[r0] = fma(sc, a0x, acc[r0]);
  acc[n + r0] = fma(sc, a0y, acc[n + r0]);
  acc[2 * n + r0] = fma(sc, a0z, acc[2 * n + r0]);
  acc[r1] = fma(sc, a1x, acc[r1]);
  acc[n + r1] = fma(sc, a1y, acc[n + r1]);
  acc[2 * n + r1] = fma(sc, a1z, acc[2 * n + r1]);
  __syncwarp();
}

struct LeanScratch {
  double *R;        // [dofs] configuration being evaluated
  double *acc;      // [dofs] gradient accumulator (single copy)
  double *red;      // [nwarps*4 + 8]
  double *strip;    // this warp's column strip
};

QMC_DEVICE LeanScratch make_lean_scratch(const Plan &plan, unsigned char *smem, int group_index, int nwarps, int gwarp, size_t group_bytes)
{
  LeanScratch S;
  double *base = reinterpret_cast<double *>(smem + (size_t)group_index * group_bytes);
  S.R = base;
  S.acc = base + plan.dofs;
  S.red = S.acc + plan.dofs;
  S.strip = S.red + (nwarps * 4 + 8) + (size_t)gwarp * STRIP;
  return S;
}

// log psi, laplacian, V of the configuration in S.R; S.acc = grad log psi on return (all threads of the group).
// Schedule: the particles are cut in 64-blocks; the pairs of DIFFERENT 64-blocks are visited in the rounds of a round-robin
// tournament (in one round every 64-block belongs to at most one warp, which therefore owns its gradient rows), then every
// 64-block is paired with itself (two diagonal 32-tiles + one 32 x 32 tile).
template <int VK>
QMC_DEVICE EvalResult lean_evaluate(const Plan &plan, const LeanParams &LP, const double *tab, const LeanScratch &S, const Group &G)
{
  const int n = plan.N[0], dofs = plan.dofs;
  const Box box = {plan.L[0], plan.L[1], plan.L[2], plan.Linv[0], plan.Linv[1], plan.Linv[2]};
  for (int k = G.gtid; k < dofs; k += G.gsize) S.acc[k] = 0.0;
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
          if (M * 64 + 64 <= n) lean_pairblock<VK>(LP, tab, box, S.R, n, M, Mp, S.strip, G.lane, S.acc, sums);
          else
            { /* ragged last 64-block: its one or two 32-blocks against the two (full) 32-blocks of Mp */
              for (int rb = 2 * M; rb < nb; ++rb)
                for (int cb = 2 * Mp; cb < 2 * Mp + 2; ++cb)
                  {
                    if (rb * 32 + 32 <= n) lean_tile<VK, false, false>(LP, tab, box, S.R, n, rb, cb, S.strip, G.lane, S.acc, sums);
                    else lean_tile<VK, false, true>(LP, tab, box, S.R, n, rb, cb, S.strip, G.lane, S.acc, sums);
                  }
            }
        }
      G.sync();
    }
  for (int d = G.gwarp; d < nB; d += G.nwarps)
    {
      const int b0 = 2 * d, b1 = 2 * d + 1;
      if (b0 * 32 + 32 <= n) lean_tile<VK, true, false>(LP, tab, box, S.R, n, b0, b0, S.strip, G.lane, S.acc, sums);
      else lean_tile<VK, true, true>(LP, tab, box, S.R, n, b0, b0, S.strip, G.lane, S.acc, sums);
      if (b1 < nb)
        {
          if (b1 * 32 + 32 <= n)
            {
              lean_tile<VK, true, false>(LP, tab, box, S.R, n, b1, b1, S.strip, G.lane, S.acc, sums);
              lean_tile<VK, false, false>(LP, tab, box, S.R, n, b1, b0, S.strip, G.lane, S.acc, sums);
            }
          else
            {
              lean_tile<VK, true, true>(LP, tab, box, S.R, n, b1, b1, S.strip, G.lane, S.acc, sums);
              lean_tile<VK, false, true>(LP, tab, box, S.R, n, b1, b0, S.strip, G.lane, S.acc, sums);
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

template <int VK, int MINB>
__global__ void __launch_bounds__(CTA_THREADS, MINB)
pair_move_kernel(const __grid_constant__ Plan plan, const __grid_constant__ DmcParams dmc, const __grid_constant__ LeanParams LP, const StepArgs a,
                 double *e_old, int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (CTA_THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  const LeanScratch S = make_lean_scratch(plan, smem, g, nwarps_per_group, G.gwarp, group_bytes);
  double *tab = reinterpret_cast<double *>(smem + (size_t)groups * group_bytes);        /* 2^(j/256), shared by the CTA */
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = LP.tab[i];
  __syncthreads();
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
      const EvalResult r = lean_evaluate<VK>(plan, LP, tab, S, G);
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
template <int VK, int MINB> void pair_tramp(void **p)
{
  pair_move_kernel<VK, MINB>(*(Plan *)p[0], *(DmcParams *)p[1], *(LeanParams *)p[2], *(StepArgs *)p[3], *(double **)p[4], *(int *)p[5], *(size_t *)p[6]);
}
template <int VK, int MINB> const void *pair_fn() { return (const void *)&pair_tramp<VK, MINB>; }
#else
template <int VK, int MINB> const void *pair_fn() { return (const void *)pair_move_kernel<VK, MINB>; }
#endif

const void *pair_ptr(int vk, int occ)
{
  if (vk == V_STEP) return occ >= 3 ? pair_fn<V_STEP, 3>() : pair_fn<V_STEP, 2>();
  return occ >= 3 ? pair_fn<V_NONE, 3>() : pair_fn<V_NONE, 2>();
}

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
  size_t d = (size_t)2 * plan.dofs + (nwarps * 4 + 8) + (size_t)nwarps * STRIP;
  d = (d + 1) & ~(size_t)1;
  return d * 8;
}
size_t pair_cta_extra_bytes() { return 256 * sizeof(double); }     /* the exp table */

// the 256 table values 2^(j/256) (host, correctly rounded from long double)
void pair_exp_table(double *tab256)
{
  for (int j = 0; j < 256; ++j) tab256[j] = (double)exp2l((long double)j / 256.0L);
}

cudaError_t configure_kernels_pair(int max_smem_optin)
{
  for (int vk : {(int)V_STEP, (int)V_NONE})
    for (int occ = 2; occ <= 3; ++occ)
      {
        cudaError_t e = cudaFuncSetAttribute(pair_ptr(vk, occ), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
        if (e != cudaSuccess) return e;
      }
  return cudaSuccess;
}

int pair_move_occupancy(int vk, int occ, size_t smem)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pair_ptr(vk, occ), CTA_THREADS, smem) != cudaSuccess) return 1;
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
  size_t gb = (cfg.smem - pair_cta_extra_bytes()) / cfg.groups_per_cta;
  /* constants of the pair loop, folded on the host (jastrowGaussGauss.h:19-21: u = u0 exp(-alpha r^2), alpha = 1/(2 r0^2)) */
  const PairPass &PP = plan.p2[0];
  const long double al = PP.J.p[0], u0 = PP.J.p[1], ln2 = 0.693147180559945309417232121458176568L;
  LeanParams LP;
  LP.k = (double)(-al * 256.0L / ln2);
  LP.c = (double)(ln2 / (256.0L * al));
  LP.c1 = (double)(-al);
  LP.c2 = (double)(al * al / 2.0L);
  LP.c3 = (double)(-al * al * al / 6.0L);
  LP.c4 = (double)(al * al * al * al / 24.0L);
  LP.sc = -2.0 * PP.J.p[1] * PP.J.p[0];
  LP.u0 = PP.J.p[1];
  LP.lap0 = -12.0 * PP.J.p[0] * PP.J.p[1];
  LP.lap1 = 8.0 * PP.J.p[0] * PP.J.p[0] * PP.J.p[1];
  LP.v_val = PP.v_val;
  memcpy(&LP.thr_bits, &PP.v_thr2, 8);
  LP.tab = dev_tab;
  void *args[] = {(void *)&plan, (void *)&dmc, (void *)&LP, (void *)&a, (void *)&e_old, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(pair_ptr(vk, occ), dim3(grid), dim3(cfg.block), args, cfg.smem, st);
}

} // namespace qmc
