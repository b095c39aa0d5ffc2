// kernels.cu — the sm_100a kernels of the DMC walker-propagation path.
//
//   propagate_kernel : one DMC move per walker, fully fused —
//        gather parent walker (branching copy/kill fused into the load, branching.cpp:34-61)
//        -> drift + diffusion with Philox normals (dmcMoves.cpp:12-30, vmcMoves.cpp:11-35)
//        -> evaluate log psi / grad / laplacian / V / E_L in one pair pass (evaluate.cuh)
//        -> log Green ratio (dmcMoves.cpp:32-57) + Metropolis (metropolis.cpp:5-29, dmcDriver.cpp:27-38)
//        -> accept / reject (dmcDriver.cpp:56-84) -> nDescendants (branching.cpp:11-14)
//        -> estimators on sampling steps.
//   (the branching scan shared by both pipelines lives in kernels_scan.cu)
//   eval_kernel      : evaluation of given configurations (parity entry, initial population,
//                      estimator accumulation over given configurations).
#include <algorithm>
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

// ---------------------------------------------------------------------------------------------
template <int JK, int VK, bool HIST>
__global__ void __launch_bounds__(CTA_THREADS, QMC_MIN_BLOCKS)
eval_kernel(const __grid_constant__ Plan plan, const EvalArgs a, int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (CTA_THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  const Scratch S = make_scratch(plan, smem, g, nwarps_per_group, group_bytes);
  const int dofs = plan.dofs;
  HistCtx H;
  H.base = (HIST && a.hist) ? a.hist + (size_t)blockIdx.x * plan.hist_stride : nullptr;
  H.w = 1.0;
  H.fx = (HIST && a.hist) ? reinterpret_cast<unsigned *>(smem + (size_t)groups * group_bytes) : nullptr;
  if (HIST && H.fx)
    {
      hist_fx_zero(H.fx, plan.hist_stride);
      __syncthreads();
    }
  for (long j = (long)blockIdx.x * groups + g; j < a.nW; j += (long)gridDim.x * groups)
    {
      const double *src = a.R + j * dofs;
      for (int k = G.gtid; k < dofs; k += G.gsize) S.R[k] = src[k];
      G.sync();
      EvalResult r = evaluate_walker<JK, VK, HIST>(plan, a.coef, S, G, H);
      if (HIST && H.base && r.valid)
        {
          histogram_pass(plan, S, G, H);
          post_estimators(plan, S, G, H);
        }
      double f2 = 0;
      for (int k = G.gtid; k < dofs; k += G.gsize)
        {
          const double f = S.acc[k];
          f2 = fma(f, f, f2);
          if (a.F) a.F[j * dofs + k] = r.valid ? f : 0.0;
        }
      f2 = warp_sum(f2);
      if (G.lane == 0) S.red[G.nwarps * 4 + G.gwarp] = f2;
      G.sync();
      if (G.gtid == 0)
        {
          double s = 0;
          for (int w = 0; w < G.nwarps; ++w) s += S.red[G.nwarps * 4 + w];
          const double E = r.V - 0.5 * (r.lap + s);     /* energy.cpp:6-35 */
          if (a.logpsi) a.logpsi[j] = r.valid ? r.logpsi : 0.0;
          if (a.lap) a.lap[j] = r.valid ? r.lap : 0.0;
          if (a.V) a.V[j] = r.valid ? r.V : 0.0;
          if (a.E) a.E[j] = r.valid ? E : 0.0;
          if (a.valid) a.valid[j] = r.valid;
          if (!r.valid && a.flag_invalid) a.ctl->invalid_init = 1;
        }
      G.sync();
    }
  if (HIST && H.fx)
    {
      __syncthreads();
      hist_fx_flush(plan, H.fx, H.base);
    }
}

// ---------------------------------------------------------------------------------------------
template <int JK, int VK, bool HIST>
__global__ void __launch_bounds__(CTA_THREADS, QMC_MIN_BLOCKS)
propagate_kernel(const __grid_constant__ Plan plan, const __grid_constant__ DmcParams dmc, const StepArgs a,
                 int nwarps_per_group, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(nwarps_per_group);
  const int groups = (CTA_THREADS / 32) / nwarps_per_group;
  const int g = (threadIdx.x >> 5) / nwarps_per_group;
  const Scratch S = make_scratch(plan, smem, g, nwarps_per_group, group_bytes);
  const int dofs = plan.dofs;
  const Ctl *ctl = a.ctl;
  const int nW = ctl->n_cur;
  /* E_T from the (all-reduced) sums of the population this step starts from (branching.cpp:65-97) */
  const double eT = dmc.branching_on ? branch_shift(dmc, ctl->gsum[0], ctl->gsum[1]) : 0.0;
  const unsigned long long step = ctl->step;
  const uint32_t step_lo = (uint32_t)step, step_hi = (uint32_t)(step >> 32) | ((uint32_t)dmc.rank << 24);
  double *hist_replica = (HIST && a.hist) ? a.hist + (size_t)blockIdx.x * plan.hist_stride : nullptr;
  unsigned *hist_fx = (HIST && a.hist) ? reinterpret_cast<unsigned *>(smem + (size_t)groups * group_bytes) : nullptr;
  double n_moves = 0, n_acc = 0;
  if (HIST && hist_fx)
    {
      hist_fx_zero(hist_fx, plan.hist_stride);
      __syncthreads();
    }

  for (int j = blockIdx.x * groups + g; j < nW; j += gridDim.x * groups)
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
                  const double rn = fma(dmc.sqrt_tau, z[m], fma(dmc.tau, Fold[k], Rold[k]));
                  S.R[k] = rn;
                  if (a.trace.xi) a.trace.xi[(size_t)j * dofs + k] = z[m];
                }
            }
        }
      G.sync();
      // ---- evaluate the proposal
      HistCtx H;
      H.base = hist_replica;
      H.w = 1.0;
      H.fx = hist_fx;
      const EvalResult r = evaluate_walker<JK, VK, HIST>(plan, a.coef, S, G, H);
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
      const double u_b = u01_53(rd.z, rd.w);
      const double A = 2.0 * (r.logpsi - lp_old) + (dmc.vmc ? 0.0 : qs);     /* vmc: vmcDriver.cpp:68 */
      const bool accepted = r.valid && (A > 0.0 || A > log(u_m));
      const double Efin = accepted ? Enew : Eold;
      int n = 1;
      if (dmc.branching_on) n = (int)(exp(-dmc.tau * (0.5 * (Efin + Eold) - eT)) + u_b);   /* branching.cpp:13 */
      n = max(0, min(n, a.cap));
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
          a.n_desc[j] = n;
          n_moves += 1;
          n_acc += accepted ? 1 : 0;
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
              a.trace.u_b[j] = u_b;
              a.trace.accepted[j] = accepted;
              a.trace.valid[j] = r.valid;
              a.trace.n_desc[j] = n;
              a.trace.parent[j] = src;
            }
        }
      // ---- estimators (sampling steps): the reference accumulates over the POST-branching
      //      population (dmcDriver.cpp:257-264), i.e. the final configuration with weight n.
      if (HIST && hist_replica && n > 0)
        { // S.R must hold the final configuration: the proposal if accepted, the parent otherwise
          G.sync();
          if (!accepted)
            {
              for (int k = G.gtid; k < dofs; k += G.gsize) S.R[k] = Rold[k];
              G.sync();
            }
          HistCtx Hn = {hist_replica, (double)n, hist_fx};
          histogram_pass(plan, S, G, Hn);
          post_estimators(plan, S, G, Hn);
        }
      G.sync();
    }
  if (G.gtid == 0)
    {
      double *p = a.partials + ((size_t)blockIdx.x * groups + g) * 2;
      p[0] = n_moves;
      p[1] = n_acc;
    }
  if (HIST && hist_fx)
    {
      __syncthreads();
      hist_fx_flush(plan, hist_fx, hist_replica);
    }
}

__global__ void finalize_shift_kernel(const __grid_constant__ DmcParams dmc, Ctl *ctl)
{
  if (threadIdx.x == 0 && blockIdx.x == 0 && dmc.branching_on) ctl->eshift = branch_shift(dmc, ctl->gsum[0], ctl->gsum[1]);
}

__global__ void init_identity_kernel(int *parent, int *n_desc, int n, Ctl *ctl)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    {
      parent[i] = i;
      n_desc[i] = 1;
    }
  if (i == 0) ctl->n_cur = n;
}

__global__ void reduce_hist_kernel(const double *hist, int replicas, int stride, double *out)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= stride) return;
  double s = 0;
  for (int r = 0; r < replicas; ++r) s += hist[(size_t)r * stride + b];
  out[b] = s;
}

// jastrow<T>::print (jastrow.h:42-63): u, u', u'' of one radial functor at the n bin centres of [minx, maxx] — the
// device's own evaluation of the functor, written by the host driver to wave<i>.dat (main.cpp:220-227)
__global__ void tabulate_kernel(const RadialFn J, const double *coef, double minx, double deltax, long n, double *out)
{
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
      /* jastrow.h:51 rounds the product before the sum (host code, no contraction): no FMA here */
      const double x = __dadd_rn(minx, __dmul_rn((double)i + 0.5, deltax));
      double d0, d1, d2;
      eval_radial(J, coef, x, d0, d1, d2);
      out[4 * i + 0] = x;
      out[4 * i + 1] = d0;
      out[4 * i + 2] = d1;
      out[4 * i + 3] = d2;
    }
}

// walker (un)packing for migration: slab = count x (2*dofs + 2) doubles  {R, F, logpsi, E}
__global__ void pack_kernel(WalkerBuf buf, const int *parent, int dofs, int first, int count, double *slab)
{
  const int w = blockIdx.x;
  if (w >= count) return;
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
__global__ void unpack_kernel(WalkerBuf buf, int dofs, int first, int count, const double *slab)
{
  const int w = blockIdx.x;
  if (w >= count) return;
  const double *src = slab + (size_t)w * (2 * dofs + 2);
  const size_t s = (size_t)(first + w) * dofs;
  for (int k = threadIdx.x; k < dofs; k += blockDim.x)
    {
      buf.R[s + k] = src[k];
      buf.F[s + k] = src[dofs + k];
    }
  if (threadIdx.x == 0)
    {
      buf.logpsi[first + w] = src[2 * dofs];
      buf.E[first + w] = src[2 * dofs + 1];
    }
}

// FP64 pipe peak: 8 independent DFMA chains per thread, enough CTAs to fill every SM
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        {
          x0 = fma(x0, a, b);
          x1 = fma(x1, a, b);
          x2 = fma(x2, a, b);
          x3 = fma(x3, a, b);
          x4 = fma(x4, a, b);
          x5 = fma(x5, a, b);
          x6 = fma(x6, a, b);
          x7 = fma(x7, a, b);
        }
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

#ifdef QMC_EMULATE   /* tests/emu: cudaLaunchKernel gets a trampoline that unpacks the argument array */
template <int JK, int VK, bool HIST> struct KernelPtrs {
  static void eval_tramp(void **p)
  {
    eval_kernel<JK, VK, HIST>(*(Plan *)p[0], *(EvalArgs *)p[1], *(int *)p[2], *(size_t *)p[3]);
  }
  static void prop_tramp(void **p)
  {
    propagate_kernel<JK, VK, HIST>(*(Plan *)p[0], *(DmcParams *)p[1], *(StepArgs *)p[2], *(int *)p[3], *(size_t *)p[4]);
  }
  static const void *eval() { return (const void *)&eval_tramp; }
  static const void *prop() { return (const void *)&prop_tramp; }
};
#else
template <int JK, int VK, bool HIST> struct KernelPtrs {
  static const void *eval() { return (const void *)eval_kernel<JK, VK, HIST>; }
  static const void *prop() { return (const void *)propagate_kernel<JK, VK, HIST>; }
};
#endif

const void *eval_ptr(int ksel, bool hist)
{
  switch (ksel * 2 + (hist ? 1 : 0))
    {
    case 0: return KernelPtrs<J_GAUSSGAUSS, V_STEP, false>::eval();
    case 1: return KernelPtrs<J_GAUSSGAUSS, V_STEP, true>::eval();
    case 2: return KernelPtrs<J_GAUSSGAUSS, V_NONE, false>::eval();
    case 3: return KernelPtrs<J_GAUSSGAUSS, V_NONE, true>::eval();
    case 4: return KernelPtrs<J_GAUSSIAN, V_NONE, false>::eval();
    case 5: return KernelPtrs<J_GAUSSIAN, V_NONE, true>::eval();
    case 6: return KernelPtrs<J_GENERIC, V_GENERIC, false>::eval();
    default: return KernelPtrs<J_GENERIC, V_GENERIC, true>::eval();
    }
}
const void *prop_ptr(int ksel, bool hist)
{
  switch (ksel * 2 + (hist ? 1 : 0))
    {
    case 0: return KernelPtrs<J_GAUSSGAUSS, V_STEP, false>::prop();
    case 1: return KernelPtrs<J_GAUSSGAUSS, V_STEP, true>::prop();
    case 2: return KernelPtrs<J_GAUSSGAUSS, V_NONE, false>::prop();
    case 3: return KernelPtrs<J_GAUSSGAUSS, V_NONE, true>::prop();
    case 4: return KernelPtrs<J_GAUSSIAN, V_NONE, false>::prop();
    case 5: return KernelPtrs<J_GAUSSIAN, V_NONE, true>::prop();
    case 6: return KernelPtrs<J_GENERIC, V_GENERIC, false>::prop();
    default: return KernelPtrs<J_GENERIC, V_GENERIC, true>::prop();
    }
}

} // namespace

cudaError_t configure_kernels(int max_smem_optin)
{
  for (int k = 0; k < KSEL_COUNT; ++k)
    for (int h = 0; h < 2; ++h)
      {
        cudaError_t e = cudaFuncSetAttribute(eval_ptr(k, h), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(prop_ptr(k, h), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
        if (e != cudaSuccess) return e;
      }
  return cudaSuccess;
}

int propagate_regs(int ksel, bool hist)
{
  cudaFuncAttributes at;
  if (cudaFuncGetAttributes(&at, prop_ptr(ksel, hist)) != cudaSuccess) return -1;
  return at.numRegs;
}
int propagate_occupancy(int ksel, bool hist, size_t smem)
{
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, prop_ptr(ksel, hist), CTA_THREADS, smem) != cudaSuccess) return 1;
  return n < 1 ? 1 : n;
}
int eval_regs(int ksel, bool hist)
{
  cudaFuncAttributes at;
  if (cudaFuncGetAttributes(&at, eval_ptr(ksel, hist)) != cudaSuccess) return -1;
  return at.numRegs;
}

cudaError_t launch_eval(int ksel, bool hist, const LaunchCfg &cfg, cudaStream_t st, const Plan &plan, const EvalArgs &a)
{
  int nw = cfg.nwarps_per_group;
  size_t gb = (cfg.smem - (size_t)plan.hist_stride * 8) / cfg.groups_per_cta;
  void *args[] = {(void *)&plan, (void *)&a, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(eval_ptr(ksel, hist), dim3(cfg.grid), dim3(cfg.block), args, cfg.smem, st);
}

cudaError_t launch_propagate(int ksel, bool hist, const LaunchCfg &cfg, cudaStream_t st, const Plan &plan,
                             const DmcParams &dmc, const StepArgs &a)
{
  int nw = cfg.nwarps_per_group;
  size_t gb = (cfg.smem - (size_t)plan.hist_stride * 8) / cfg.groups_per_cta;
  void *args[] = {(void *)&plan, (void *)&dmc, (void *)&a, (void *)&nw, (void *)&gb};
  return cudaLaunchKernel(prop_ptr(ksel, hist), dim3(cfg.grid), dim3(cfg.block), args, cfg.smem, st);
}

cudaError_t launch_finalize_shift(cudaStream_t st, const DmcParams &dmc, Ctl *ctl)
{
  QMC_LAUNCH(finalize_shift_kernel, 1, 32, 0, st, dmc, ctl);
  return cudaGetLastError();
}

cudaError_t launch_init_identity(cudaStream_t st, int *parent, int *n_desc, int n, Ctl *ctl)
{
  QMC_LAUNCH(init_identity_kernel, (n + 255) / 256 + 1, 256, 0, st, parent, n_desc, n, ctl);
  return cudaGetLastError();
}

cudaError_t launch_reduce_hist(cudaStream_t st, const double *hist, int replicas, int stride, double *out)
{
  QMC_LAUNCH(reduce_hist_kernel, (stride + 127) / 128, 128, 0, st, hist, replicas, stride, out);
  return cudaGetLastError();
}

cudaError_t launch_tabulate(cudaStream_t st, const RadialFn &J, const double *coef, double minx, double deltax, long n, double *out)
{
  if (n <= 0) return cudaSuccess;
  QMC_LAUNCH(tabulate_kernel, (int)std::min<long>((n + 127) / 128, 1184), 128, 0, st, J, coef, minx, deltax, n, out);
  return cudaGetLastError();
}

cudaError_t launch_pack(cudaStream_t st, const WalkerBuf &buf, const int *parent, int dofs, int first, int count, double *slab)
{
  if (count <= 0) return cudaSuccess;
  QMC_LAUNCH(pack_kernel, count, 128, 0, st, buf, parent, dofs, first, count, slab);
  return cudaGetLastError();
}
cudaError_t launch_unpack(cudaStream_t st, const WalkerBuf &buf, int dofs, int first, int count, const double *slab)
{
  if (count <= 0) return cudaSuccess;
  QMC_LAUNCH(unpack_kernel, count, 128, 0, st, buf, dofs, first, count, slab);
  return cudaGetLastError();
}

double measure_fp64_peak_tflops(cudaStream_t st, float *ms_out)
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, iters = 4096;
  double *out = nullptr;
  if (cudaMalloc(&out, (size_t)grid * 256 * sizeof(double)) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
    {
      cudaEventRecord(e0, st);
      QMC_LAUNCH(fp64_peak_kernel, grid, 256, 0, st, out, iters, 0.999999, 1e-9);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (ms_out) *ms_out = best;
  const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)grid;
  return flops / (best * 1e-3) * 1e-12;
}

} // namespace qmc
