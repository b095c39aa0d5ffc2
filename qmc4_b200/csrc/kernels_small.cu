// kernels_small.cu — the move step of pipeline 2 for SMALL systems (one species, N <= 16 particles, one- and two-body
// terms: BASELINE config 1, the harmonic trap of tst/sampleInputs): SEVERAL WALKERS PER WARP.
//
// One warp per walker (the general path) keeps 10 of 32 lanes busy at N = 10 and spends its time in warp-wide barriers and
// shared-memory reductions that a 45-pair system does not need.  Here a walker owns a power-of-two SEGMENT of a warp
// (16 lanes for 9..16 particles, 8 for 5..8, 4 below: 2, 4 or 8 walkers per warp), lane i of the segment owns particle i:
//   * proposal: R' = R + tau F + sqrt(tau) xi with exactly the Philox draws of the general kernels (counter = {slot, dof/4,
//     step, rank}), so trajectories do not depend on which kernel ran (dmcMoves.cpp:12-30, vmcMoves.cpp:11-35);
//   * evaluation: cyclic pair schedule inside the segment (lane i meets particle i+k, k = 1..N/2, every unordered pair once);
//     the partner's share of the gradient travels back with one segmented shuffle per component, nothing is scattered and
//     no barrier is wider than the segment
//     (jastrowWavefunctionTwoBody.h:124-164, jastrowWavefunctionOneBody.h:53-85, potential.cpp:21-32, potentialTwoBody.h:23-33);
//   * reductions: segmented xor-shuffles (the width argument of shfl), Green ratio (dmcMoves.cpp:32-57), Metropolis
//     (metropolis.cpp:5-29), commit or carry the parent through (dmcDriver.cpp:56-84).
// The step is latency-bound at BASELINE size (1000 walkers = 63 warps on 148 SMs); packing matters once the population
// fills the machine.  Roofline for this shape: HBM / latency, see DESIGN.md.
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

constexpr int SMALL_THREADS = 256;

// sum over the aligned power-of-two segment of SW lanes: lane ^ m stays inside the segment for m < SW
template <int SW> QMC_DEVICE double seg_sum(double v)
{
#pragma unroll
  for (int m = SW / 2; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}

template <int SW>
__global__ void __launch_bounds__(SMALL_THREADS, 4)
small_move_kernel(const __grid_constant__ Plan plan, const __grid_constant__ DmcParams dmc, const StepArgs a, double *e_old)
{
  constexpr int WPC = SMALL_THREADS / SW;                 /* walkers per CTA */
  __shared__ double sR[WPC][3 * SW];
  __shared__ double sZ[WPC][3 * SW];
  const int n = plan.N[0], dofs = plan.dofs;
  const int lane = threadIdx.x & 31, i = lane % SW, ws = threadIdx.x / SW, segbase = lane - i;
  const Ctl *ctl = a.ctl;
  const int j0 = a.count < 0 ? 0 : (a.first < 0 ? ctl->n_cur + a.first + 1 : a.first);
  const int j1 = a.count < 0 ? ctl->n_cur : j0 + a.count;
  const unsigned long long step = ctl->step;
  const uint32_t step_lo = (uint32_t)step, step_hi = (uint32_t)(step >> 32) | ((uint32_t)dmc.rank << 24);
  const double Lx = plan.L[0], Ly = plan.L[1], Lz = plan.L[2], ix = plan.Linv[0], iy = plan.Linv[1], iz = plan.Linv[2];
  double *R = sR[ws], *Z = sZ[ws];
  const int K = n >> 1;                                   /* cyclic offsets 1..K; offset n/2 is half-populated when n is even */

  /* the segments of a warp run the same number of turns (full-warp shuffles stay legal); a segment without a walker idles */
  for (int jb = j0 + blockIdx.x * WPC; jb < j1; jb += gridDim.x * WPC)
    {
      const int j = min(jb + ws, j1 - 1);
      const bool live = jb + ws < j1;
      const bool act = live && i < n;
      const int src = a.parent[j];
      const double lp_old = a.in.logpsi[src], Eold = a.in.E[src];
      const double *Rold = a.in.R + (size_t)src * dofs;
      const double *Fold = a.in.F + (size_t)src * dofs;
      double ro[3] = {0, 0, 0}, fo[3] = {0, 0, 0}, rn[3] = {0, 0, 0};
      if (act)
        {
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              ro[d] = Rold[d * n + i];
              fo[d] = Fold[d * n + i];
            }
        }
      /* deviates: lane c of the segment draws Philox block c = coordinates 4c..4c+3 (3n <= 4 SW; the stream of the general kernels) */
      if (live && i * 4 < dofs)
        {
          const u32x4 rnd = philox4x32_10({(uint32_t)j, (uint32_t)i, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
          double z[4];
          normal_pair(rnd.x, rnd.y, z[0], z[1]);
          normal_pair(rnd.z, rnd.w, z[2], z[3]);
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (i * 4 + m < dofs) Z[i * 4 + m] = z[m];
        }
      __syncwarp();
      if (act)
        {
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              const int k = d * n + i;
              const double z = Z[k];
              rn[d] = fma(dmc.sqrt_tau, z, fma(dmc.tau, fo[d], ro[d]));
              R[d * SW + i] = rn[d];
              if (a.trace.xi) a.trace.xi[(size_t)j * dofs + k] = z;
            }
        }
      __syncwarp();
      // ---- evaluation of the proposal
      double lp = 0, lap = 0, V = 0, gx = 0, gy = 0, gz = 0;
      if (act)
        for (int t = 0; t < plan.n1b; ++t)
          { /* jastrowOneBodyWavefunction (laplacian has no factor 2) + harmonicPotential; "distance from the origin" is min-imaged */
            const OneBodyTerm &T = plan.t1[t];
            const double dx = min_image(rn[0], Lx, ix), dy = min_image(rn[1], Ly, iy), dz = min_image(rn[2], Lz, iz);
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            if (T.J.kind != J_NONE)
              {
                double d0, d1, d2;
                eval_radial(T.J, a.coef, r, d0, d1, d2);
                lap += d2 + 2 * d1 / r;
                lp += d0;
                gx += dx / r * d1;
                gy += dy / r * d1;
                gz += dz / r * d1;
              }
            if (T.has_pot) V += T.half_omega2 * r * r;
          }
      /* pairs: cyclic schedule inside the segment, each unordered pair once: lane i meets particle i+k and hands the partner's
         share of the gradient to it with one segmented shuffle (the lane that met ME at this offset is i-k) */
      for (int k = 1; k <= K; ++k)
        {
          int q = i + k;
          if (q >= n) q -= n;
          int from = i - k;
          if (from < 0) from += n;
          const bool pa = act && !((2 * k == n) && (i >= K));
          const int qq = pa ? q : 0;
          const double dx = min_image(rn[0] - R[qq], Lx, ix);
          const double dy = min_image(rn[1] - R[SW + qq], Ly, iy);
          const double dz = min_image(rn[2] - R[2 * SW + qq], Lz, iz);
          const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
          double gsum = 0;
          if (pa)
            {
              double r = 0;
              bool have_r = false;
              for (int p = 0; p < plan.n2b; ++p)
                {
                  const PairPass &P = plan.p2[p];
                  if (P.J.kind == J_GAUSSIAN)
                    { /* u = -a r^2, u'/r = -2a, u'' = -2a */
                      gsum += -2.0 * P.J.p[0];
                      lp = fma(-P.J.p[0], r2, lp);
                      lap += -12.0 * P.J.p[0];
                    }
                  else if (P.J.kind == J_GAUSSGAUSS)
                    {
                      const double e = exp_gg(r2, P.J.p);
                      const double al = P.J.p[0], u0 = P.J.p[1];
                      gsum += -2.0 * u0 * al * e;
                      lp = fma(u0, e, lp);
                      lap += e * (8.0 * al * al * u0 * r2 - 12.0 * al * u0);
                    }
                  else if (P.J.kind != J_NONE)
                    {
                      if (!have_r)
                        {
                          r = sqrt(r2);
                          have_r = true;
                        }
                      double d0, d1, d2;
                      eval_radial(P.J, a.coef, r, d0, d1, d2);
                      const double gr = d1 / r;
                      gsum += gr;
                      lp += d0;
                      lap += 2 * (d2 + 2 * gr);      /* jastrowWavefunctionTwoBody.h:150, D = 3 */
                    }
                  if (P.vkind == V_STEP) V += (r2 <= P.v_thr2) ? P.v_val : 0.0;
                  else if (P.vkind != V_NONE)
                    {
                      if (!have_r)
                        {
                          r = sqrt(r2);
                          have_r = true;
                        }
                      V += eval_pair_potential(P.vkind, P.v_val, r);
                    }
                }
            }
          const double cx = gsum * dx, cy = gsum * dy, cz = gsum * dz;      /* 0 on idle lanes */
          gx += cx;
          gy += cy;
          gz += cz;
          const int fl = segbase + (act ? from : i);
          gx -= shfl_d(cx, fl);
          gy -= shfl_d(cy, fl);
          gz -= shfl_d(cz, fl);
        }
      // ---- Green ratio and |grad|^2 on this lane's three coordinates, then the walker's sums over its segment
      double qv = 0, f2 = 0;
      {
        const double fn[3] = {gx, gy, gz};
#pragma unroll
        for (int d = 0; d < 3; ++d)
          {
            const double aa = 0.5 * dmc.tau * (fo[d] - fn[d]) - rn[d] + ro[d];      /* dmcMoves.cpp:46-51 */
            qv = fma(aa, fo[d] + fn[d], qv);
            f2 = fma(fn[d], fn[d], f2);
          }
      }
      lp = seg_sum<SW>(lp);
      lap = seg_sum<SW>(lap);
      V = seg_sum<SW>(V);
      qv = seg_sum<SW>(qv);
      f2 = seg_sum<SW>(f2);
      const double Enew = V - 0.5 * (lap + f2);                  /* energy.cpp:6-35 */
      // ---- Metropolis (dmcDriver.cpp:27-38, metropolis.cpp:5-29); every lane of the segment computes the same draw
      const u32x4 rd = philox4x32_10({(uint32_t)j, 0xffffffffu, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
      const double u_m = u01_53(rd.x, rd.y);
      const double A = 2.0 * (lp - lp_old) + (dmc.vmc ? 0.0 : qv);
      const bool accepted = A > 0.0 || A > log(u_m);
      if (act)
        {
          double *Rout = a.out.R + (size_t)j * dofs;
          double *Fout = a.out.F + (size_t)j * dofs;
          const double fn[3] = {gx, gy, gz};
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              const int k = d * n + i;
              Rout[k] = accepted ? rn[d] : ro[d];          /* reject: wNew = wOld (dmcDriver.cpp:73-76) */
              Fout[k] = accepted ? fn[d] : fo[d];
              if (a.trace.Rprop)
                {
                  a.trace.Rprop[(size_t)j * dofs + k] = rn[d];
                  a.trace.Fprop[(size_t)j * dofs + k] = fn[d];
                }
            }
        }
      if (live && i == 0)
        {
          a.out.logpsi[j] = accepted ? lp : lp_old;
          a.out.E[j] = accepted ? Enew : Eold;
          e_old[j] = Eold;
          a.flags[j] = accepted ? 1 : 0;
          if (a.trace.Rprop)
            {
              a.trace.logpsi[j] = lp;
              a.trace.lap[j] = lap;
              a.trace.V[j] = V;
              a.trace.Eprop[j] = Enew;
              a.trace.q[j] = qv;
              a.trace.u_m[j] = u_m;
              a.trace.accepted[j] = accepted;
              a.trace.valid[j] = 1;
              a.trace.parent[j] = src;
            }
        }
      __syncwarp();          /* the strip is rewritten by the next walker of this segment */
    }
}

#ifdef QMC_EMULATE
template <int SW> void small_tramp(void **p) { small_move_kernel<SW>(*(Plan *)p[0], *(DmcParams *)p[1], *(StepArgs *)p[2], *(double **)p[3]); }
template <int SW> const void *small_fn() { return (const void *)&small_tramp<SW>; }
#else
template <int SW> const void *small_fn() { return (const void *)small_move_kernel<SW>; }
#endif
const void *small_ptr(int sw) { return sw == 16 ? small_fn<16>() : sw == 8 ? small_fn<8>() : small_fn<4>(); }

} // namespace

int small_segment_width(int n) { return n > 8 ? 16 : n > 4 ? 8 : 4; }
int small_walkers_per_cta(int n) { return SMALL_THREADS / small_segment_width(n); }

int small_move_occupancy(int n)
{
  int b = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, small_ptr(small_segment_width(n)), SMALL_THREADS, 0) != cudaSuccess) return 1;
  return b < 1 ? 1 : b;
}
int small_move_regs(int n)
{
  cudaFuncAttributes at;
  if (cudaFuncGetAttributes(&at, small_ptr(small_segment_width(n))) != cudaSuccess) return -1;
  return at.numRegs;
}

cudaError_t launch_small_move(int n, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc, const StepArgs &a, double *e_old)
{
  void *args[] = {(void *)&plan, (void *)&dmc, (void *)&a, (void *)&e_old};
  return cudaLaunchKernel(small_ptr(small_segment_width(n)), dim3(grid), dim3(SMALL_THREADS), args, 0, st);
}

} // namespace qmc
