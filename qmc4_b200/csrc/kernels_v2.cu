// kernels_v2.cu — the production step pipeline ("pipeline 2"; pipeline 1 of kernels.cu is kept for the traced step and
// for runs that carry genealogical records).  Bitwise identical trajectories to pipeline 1 (tests/test_pipeline2.py).
//
//     move_kernel      one DMC move per walker, needs neither E_T nor the estimators: gather the parent (branching copy/kill
//                      fused into the load) -> drift + diffusion (dmcMoves.cpp:12-30) -> evaluation -> Green ratio
//                      (dmcMoves.cpp:32-57) -> Metropolis (metropolis.cpp:5-29) -> commit or carry the parent through
//                      (dmcDriver.cpp:56-84); writes E' , E and the accepted flag per walker
//  -> scan_v2_kernel   nDescendants = int(exp(-tau((E'+E)/2 - E_T)) + u) (branching.cpp:11-14) with the Philox draw of the
//                      walker, prefix sum -> parent map, sums, E_T inputs                 [waits for all-reduce k-1]
//  -> all-reduce k     of {sum E, W, populations, free slots} on the communication stream, overlapped with
//  -> estimator_kernel (post-branching population with unit weight, exactly what dmcDriver::accumulate iterates,
//                      dmcDriver.cpp:257-264) and with move_kernel of step k+1;
//  migration (ptools.cpp:271-401, one step late like dmcDriver.cpp:112,150-157): pack_tail_kernel -> ncclSend/Recv ->
//  unpack_append_kernel on the communication stream while move_kernel(k+1) runs, then move_kernel on the arrivals only.
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

template <int JK, int VK>
__global__ void __launch_bounds__(CTA_THREADS, QMC_MIN_BLOCKS)
move_kernel(const __grid_constant__ Plan plan, const __grid_constant__ DmcParams dmc, const StepArgs a, double *e_old,
            int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (CTA_THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  const Scratch S = make_scratch(plan, smem, g, nwarps_per_group, group_bytes);
  const int dofs = plan.dofs;
  const Ctl *ctl = a.ctl;
  /* walkers [j0, j1): the whole population, or the walkers that arrived by migration during this step */
  const int j0 = a.count < 0 ? 0 : (a.first < 0 ? ctl->n_cur + a.first + 1 : a.first);      /* first = -1 - d: n_cur - d */
  const int j1 = a.count < 0 ? ctl->n_cur : j0 + a.count;
  const unsigned long long step = ctl->step;
  const uint32_t step_lo = (uint32_t)step, step_hi = (uint32_t)(step >> 32) | ((uint32_t)dmc.rank << 24);

  for (int j = j0 + blockIdx.x * groups + g; j < j1; j += gridDim.x * groups)
    {
      const int src = a.parent[j];
      const double *Rold = a.in.R + (size_t)src * dofs;
      const double *Fold = a.in.F + (size_t)src * dofs;
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
      HistCtx H = {nullptr, 1.0, nullptr};
      const EvalResult r = evaluate_walker<JK, VK, false>(plan, a.coef, S, G, H);
      double q = 0, f2 = 0;
      for (int k = G.gtid; k < dofs; k += G.gsize)
        {
          const double fo = Fold[k], ro = Rold[k], fn = S.acc[k], rn = S.R[k];
          const double aa = 0.5 * dmc.tau * (fo - fn) - rn + ro;
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
      const double Enew = r.V - 0.5 * (r.lap + f2s);
      const double lp_old = a.in.logpsi[src];
      const double Eold = a.in.E[src];
      const u32x4 rd = philox4x32_10({(uint32_t)j, 0xffffffffu, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
      const double u_m = u01_53(rd.x, rd.y);
      const double A = 2.0 * (r.logpsi - lp_old) + (dmc.vmc ? 0.0 : qs);     /* vmc: vmcDriver.cpp:68 */
      const bool accepted = r.valid && (A > 0.0 || A > log(u_m));
      const double Efin = accepted ? Enew : Eold;
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
              a.trace.valid[j] = r.valid;
              a.trace.parent[j] = src;
            }
        }
      G.sync();
    }
}

// Estimators over the post-branching population with unit weight per walker (dmcDriver::accumulate,
// dmcDriver.cpp:257-264): slot j of the new population is walker parent[j] of `buf`.
// MINB: resident CTAs per SM the register allocation is bounded for (2: up to 128 registers, 3: 80, 4: 64) — the
// occupancy experiment of DESIGN.md §4.3 (QMC4_EST_OCC), the histogram pass being latency-bound at 16 warps/SM.
template <int MINB>
__global__ void __launch_bounds__(CTA_THREADS, MINB)
estimator_kernel(const __grid_constant__ Plan plan, const WalkerBuf buf, const int *parent, const Ctl *ctl, double *hist,
                 int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (CTA_THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  const Scratch S = make_scratch(plan, smem, g, nwarps_per_group, group_bytes);
  const int dofs = plan.dofs;
  const int nW = ctl->n_cur;
  HistCtx H;
  H.base = hist + (size_t)blockIdx.x * plan.hist_stride;
  H.w = 1.0;
  H.fx = reinterpret_cast<unsigned *>(smem + (size_t)groups * group_bytes);
  hist_fx_zero(H.fx, plan.hist_stride);
  __syncthreads();
  for (int j = blockIdx.x * groups + g; j < nW; j += gridDim.x * groups)
    {
      const double *src = buf.R + (size_t)parent[j] * dofs;
      for (int k = G.gtid; k < dofs; k += G.gsize) S.R[k] = src[k];
      G.sync();
      histogram_pass(plan, S, G, H);
      post_estimators(plan, S, G, H);
      G.sync();
    }
  __syncthreads();
  hist_fx_flush(plan, H.fx, H.base);
}


// Migration (walkerDistribution, ptools.cpp:271-401): senders ship their TAIL walkers (:297-303), receivers append (:307-324).
// The kernels index through the control block, so the host only needs the plan (who sends how many to whom), which it derives
// from the all-reduced population vector.  Slab record = {R[dofs], F[dofs], log psi, E}.
__global__ void pack_tail_kernel(WalkerBuf buf, const int *parent, const Ctl *ctl, int dofs, int n_send, double *slab)
{
  const int first = ctl->n_cur - n_send;
  for (int w = blockIdx.x; w < n_send; w += gridDim.x)
    {
      double *dst = slab + (size_t)w * (2 * dofs + 2);
      const int p = parent[first + w];
      const size_t s = (size_t)p * dofs;
      for (int k = threadIdx.x; k < dofs; k += blockDim.x)
        {
          dst[k] = buf.R[s + k];
          dst[dofs + k] = buf.F[s + k];
        }
      if (threadIdx.x == 0)
        {
          dst[2 * dofs] = buf.logpsi[p];
          dst[2 * dofs + 1] = buf.E[p];
        }
    }
}
// arrivals land in the free physical slots [n_prev, n_prev + n_recv) and are appended to what remains of the population after
// this rank's own departures: a rank the plan first drains and then tops up both sends and receives (ptools.cpp:93-161)
__global__ void unpack_append_kernel(WalkerBuf buf, int *parent, const Ctl *ctl, int dofs, int n_send, int n_recv, const double *slab)
{
  const int slot0 = ctl->n_prev, idx0 = ctl->n_cur - n_send;
  for (int w = blockIdx.x; w < n_recv; w += gridDim.x)
    {
      const double *src = slab + (size_t)w * (2 * dofs + 2);
      const size_t s = (size_t)(slot0 + w) * dofs;
      for (int k = threadIdx.x; k < dofs; k += blockDim.x)
        {
          buf.R[s + k] = src[k];
          buf.F[s + k] = src[dofs + k];
        }
      if (threadIdx.x == 0)
        {
          buf.logpsi[slot0 + w] = src[2 * dofs];
          buf.E[slot0 + w] = src[2 * dofs + 1];
          parent[idx0 + w] = slot0 + w;
        }
    }
}
// In-step migration (qmc_step): the tail walkers travel AFTER their move of step k and BEFORE its branching — the sender has
// already advanced them, the receiver appends the moved state to its own output population and its scan branches them.
// Nothing is moved twice and no second launch of the move kernel is needed for the arrivals (walkers in flight "skip
// nothing", dmcDriver.cpp:133-136).  Slab record = {R'[dofs], F'[dofs], log psi', E', E of the parent, accepted}.
__global__ void pack_moved_kernel(WalkerBuf out, const double *e_old, const int *flags, const Ctl *ctl, int dofs, int n_send, double *slab)
{
  const int first = ctl->n_cur - n_send;
  for (int w = blockIdx.x; w < n_send; w += gridDim.x)
    {
      double *dst = slab + (size_t)w * (2 * dofs + 4);
      const int j = first + w;
      const size_t s = (size_t)j * dofs;
      for (int k = threadIdx.x; k < dofs; k += blockDim.x)
        {
          dst[k] = out.R[s + k];
          dst[dofs + k] = out.F[s + k];
        }
      if (threadIdx.x == 0)
        {
          dst[2 * dofs] = out.logpsi[j];
          dst[2 * dofs + 1] = out.E[j];
          dst[2 * dofs + 2] = e_old[j];
          dst[2 * dofs + 3] = (double)flags[j];
        }
    }
}
__global__ void unpack_moved_kernel(WalkerBuf out, double *e_old, int *flags, const Ctl *ctl, int dofs, int n_send, int n_recv, const double *slab)
{
  const int idx0 = ctl->n_cur - n_send;          /* behind the walkers that stay */
  for (int w = blockIdx.x; w < n_recv; w += gridDim.x)
    {
      const double *src = slab + (size_t)w * (2 * dofs + 4);
      const int j = idx0 + w;
      const size_t s = (size_t)j * dofs;
      for (int k = threadIdx.x; k < dofs; k += blockDim.x)
        {
          out.R[s + k] = src[k];
          out.F[s + k] = src[dofs + k];
        }
      if (threadIdx.x == 0)
        {
          out.logpsi[j] = src[2 * dofs];
          out.E[j] = src[2 * dofs + 1];
          e_old[j] = src[2 * dofs + 2];
          flags[j] = (int)src[2 * dofs + 3];
        }
    }
}
// order-independent fingerprint of the current population (conservation checks of the migration): sum over the walkers of the
// 64-bit patterns of (log psi, E) mixed by an odd multiplier, wrapping arithmetic
__global__ void population_checksum_kernel(WalkerBuf buf, const int *parent, const Ctl *ctl, unsigned long long *out)
{
  unsigned long long s = 0;
  const int n = ctl->n_cur;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
      const int p = parent[j];
      const unsigned long long a = (unsigned long long)__double_as_longlong(buf.logpsi[p]);
      const unsigned long long b = (unsigned long long)__double_as_longlong(buf.E[p]);
      s += (a * 0x9E3779B97F4A7C15ull) ^ (b + 0xD1B54A32D192ED03ull);
    }
  if (s) atomicAdd(out, s);
}

#ifdef QMC_EMULATE   /* tests/emu: cudaLaunchKernel gets a trampoline that unpacks the argument array */
template <int JK, int VK> void prop2_tramp(void **p)
{
  move_kernel<JK, VK>(*(Plan *)p[0], *(DmcParams *)p[1], *(StepArgs *)p[2], *(double **)p[3], *(int *)p[4], *(size_t *)p[5]);
}
template <int JK, int VK> const void *prop2() { return (const void *)&prop2_tramp<JK, VK>; }
#else
template <int JK, int VK> const void *prop2() { return (const void *)move_kernel<JK, VK>; }
#endif

#ifdef QMC_EMULATE
template <int MINB> void est_tramp(void **p)
{
  estimator_kernel<MINB>(*(Plan *)p[0], *(WalkerBuf *)p[1], *(const int **)p[2], *(const Ctl **)p[3], *(double **)p[4], *(int *)p[5],
                         *(size_t *)p[6]);
}
template <int MINB> const void *est_fn() { return (const void *)&est_tramp<MINB>; }
#else
template <int MINB> const void *est_fn() { return (const void *)estimator_kernel<MINB>; }
#endif
const void *est_ptr(int occ) { return occ >= 4 ? est_fn<4>() : occ == 3 ? est_fn<3>() : est_fn<2>(); }

const void *prop2_ptr(int ksel)
{
  switch (ksel)
    {
    case KSEL_GG_STEP: return prop2<J_GAUSSGAUSS, V_STEP>();
    case KSEL_GG_NONE: return prop2<J_GAUSSGAUSS, V_NONE>();
    case KSEL_GAUSS_NONE: return prop2<J_GAUSSIAN, V_NONE>();
    default: return prop2<J_GENERIC, V_GENERIC>();
    }
}

} // namespace

cudaError_t configure_kernels_v2(int max_smem_optin)
{
  for (int k = 0; k < KSEL_COUNT; ++k)
    {
      cudaError_t e = cudaFuncSetAttribute(prop2_ptr(k), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
      if (e != cudaSuccess) return e;
    }
  for (int occ = 2; occ <= 4; ++occ)
    {
      cudaError_t e = cudaFuncSetAttribute(est_ptr(occ), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
      if (e != cudaSuccess) return e;
    }
  return cudaSuccess;
}

int estimator_occupancy(int occ, size_t smem)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, est_ptr(occ), CTA_THREADS, smem) != cudaSuccess) return 1;
  return n < 1 ? 1 : n;
}

int propagate_v2_occupancy(int ksel, size_t smem)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, prop2_ptr(ksel), CTA_THREADS, smem) != cudaSuccess) return 1;
  return n < 1 ? 1 : n;
}

cudaError_t launch_propagate_v2(int ksel, const LaunchCfg &cfg, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc,
                                const StepArgs &a, double *e_old)
{
  int nw = cfg.nwarps_per_group;
  size_t gb = (cfg.smem - (size_t)plan.hist_stride * 8) / cfg.groups_per_cta;
  void *args[] = {(void *)&plan, (void *)&dmc, (void *)&a, (void *)&e_old, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(prop2_ptr(ksel), dim3(grid), dim3(cfg.block), args, cfg.smem, st);
}

cudaError_t launch_estimators(const LaunchCfg &cfg, cudaStream_t st, const Plan &plan, const WalkerBuf &buf, const int *parent,
                              const Ctl *ctl, double *hist, int occ, int grid)
{
  int nw = cfg.nwarps_per_group;
  size_t gb = (cfg.smem - (size_t)plan.hist_stride * 8) / cfg.groups_per_cta;
  void *args[] = {(void *)&plan, (void *)&buf, (void *)&parent, (void *)&ctl, (void *)&hist, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(est_ptr(occ), dim3(grid), dim3(cfg.block), args, cfg.smem, st);
}

cudaError_t launch_pack_tail(cudaStream_t st, const WalkerBuf &buf, const int *parent, const Ctl *ctl, int dofs, int n_send, double *slab)
{
  if (n_send <= 0) return cudaSuccess;
  QMC_LAUNCH(pack_tail_kernel, n_send < 16 ? n_send : 16, 256, 0, st, buf, parent, ctl, dofs, n_send, slab);      /* fits the SMs left to the comm stream */
  return cudaGetLastError();
}
cudaError_t launch_unpack_append(cudaStream_t st, const WalkerBuf &buf, int *parent, const Ctl *ctl, int dofs, int n_send, int n_recv,
                                 const double *slab)
{
  if (n_recv <= 0) return cudaSuccess;
  QMC_LAUNCH(unpack_append_kernel, n_recv < 16 ? n_recv : 16, 256, 0, st, buf, parent, ctl, dofs, n_send, n_recv, slab);
  return cudaGetLastError();
}
cudaError_t launch_pack_moved(cudaStream_t st, const WalkerBuf &out, const double *e_old, const int *flags, const Ctl *ctl, int dofs, int n_send,
                              double *slab)
{
  if (n_send <= 0) return cudaSuccess;
  QMC_LAUNCH(pack_moved_kernel, n_send < 16 ? n_send : 16, 256, 0, st, out, e_old, flags, ctl, dofs, n_send, slab);
  return cudaGetLastError();
}
cudaError_t launch_unpack_moved(cudaStream_t st, const WalkerBuf &out, double *e_old, int *flags, const Ctl *ctl, int dofs, int n_send, int n_recv,
                                const double *slab)
{
  if (n_recv <= 0) return cudaSuccess;
  QMC_LAUNCH(unpack_moved_kernel, n_recv < 16 ? n_recv : 16, 256, 0, st, out, e_old, flags, ctl, dofs, n_send, n_recv, slab);
  return cudaGetLastError();
}
cudaError_t launch_population_checksum(cudaStream_t st, const WalkerBuf &buf, const int *parent, const Ctl *ctl, unsigned long long *out)
{
  QMC_LAUNCH(population_checksum_kernel, 64, 256, 0, st, buf, parent, ctl, out);
  return cudaGetLastError();
}

} // namespace qmc
