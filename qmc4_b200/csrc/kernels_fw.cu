// kernels_fw.cu — forward walking (pure estimators) on the device: SURVEY.md §8(f)1.
//
// Reference (dmc/correlationEstimator.cpp): every walker owns a ring of `recordSteps` records of an
// observable (realHistogramStorer::store, :87-107: the walker's own histogram, sums/weight with
// weight 1); the ring is deep-copied whenever the walker is cloned by branching (walkers.cpp:139-179)
// and swapped along on accept/reject (dmcDriver.cpp:77-81), so record (i - fw - 1) of a walker is the
// observable of its ANCESTOR fw sampling events ago; realHistogramForwardWalkingEstimator::accumulate
// (:128-152) adds that record for every walker of the current population once the ring has wrapped.
//
// B200 design: nothing is copied.  Each sampling event stores ONE histogram row per surviving slot in
// a ring of max(fw)+1 events and one int per slot linking it to its ancestor slot of the previous
// event (the per-step parent maps composed between events); the estimator counts, per ancestor slot,
// how many current walkers descend from it (fw link hops per walker) and takes a count-weighted
// column sum of the rows in a fixed order (bitwise reproducible).  HBM traffic per event:
// stride*8 B per surviving slot written once + read once per fw value, against the reference's
// recordSteps*bins*8 B copied per cloned walker per step.
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

constexpr int FW_CHUNKS = 64;

// One CTA (8 warps = one walker group) per slot: the slot's own histograms with unit weight into
// rows[slot][hist_stride], through the CTA's fixed-point shared-memory histogram (evaluate.cuh).
__global__ void __launch_bounds__(CTA_THREADS)
fw_record_kernel(const __grid_constant__ Plan plan, const WalkerBuf buf, const int *n_desc, const Ctl *ctl, double *rows,
                 int *nslots_out, size_t group_bytes)
{
  QMC_DYNAMIC_SMEM(smem);
  const Group G = make_group(CTA_THREADS / 32);
  const Scratch S = make_scratch(plan, smem, 0, CTA_THREADS / 32, group_bytes);
  const int dofs = plan.dofs, stride = plan.hist_stride;
  const int n = ctl->n_prev;
  if (blockIdx.x == 0 && threadIdx.x == 0) *nslots_out = n;
  HistCtx H;
  H.w = 1.0;
  H.fx = reinterpret_cast<unsigned *>(smem + group_bytes);
  hist_fx_zero(H.fx, stride);
  __syncthreads();
  for (int j = blockIdx.x; j < n; j += gridDim.x)
    {
      if (n_desc[j] <= 0) continue;                 /* no descendant will ever look this record up */
      double *row = rows + (size_t)j * stride;
      for (int k = threadIdx.x; k < stride; k += CTA_THREADS) row[k] = 0.0;
      const double *src = buf.R + (size_t)j * dofs;
      for (int k = threadIdx.x; k < dofs; k += CTA_THREADS) S.R[k] = src[k];
      __syncthreads();
      H.base = row;
      histogram_pass(plan, S, G, H);
      post_estimators(plan, S, G, H);
      __syncthreads();
      hist_fx_flush(plan, H.fx, row);               /* adds the fixed-point totals to the row and zeroes fx */
      __syncthreads();
    }
}

// comp maps a walker index of the step about to run to its ancestor slot at the latest sampling event
__global__ void fw_compose_kernel(const int *parent, const int *comp_in, int *comp_out, const Ctl *ctl, int fresh)
{
  const int n = ctl->n_cur;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    comp_out[i] = fresh ? parent[i] : comp_in[parent[i]];
}

__global__ void fw_copy_links_kernel(const int *comp, int *link, const Ctl *ctl)
{
  const int n = ctl->n_prev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) link[i] = comp[i];
}

// cnt[a] = number of walkers of the current (post-branching) population whose ancestor `fw` events ago is slot a
__global__ void fw_count_kernel(const int *parent, const int *links, int depth, int cap, long event, int fw, const Ctl *ctl, int *cnt)
{
  const int n = ctl->n_cur;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      int a = parent[i];
      for (int h = 0; h < fw; ++h) a = links[(size_t)((event - h) % depth) * cap + a];
      atomicAdd(cnt + a, 1);
    }
}

// partial[chunk][b] = sum over the chunk's slots (ascending) of cnt[a] * rows[a][offset+b]
__global__ void fw_colsum_kernel(const double *rows, const int *cnt, const int *nslots, int stride, int offset, int bins, double *partial)
{
  const int b = blockIdx.y * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  const int n = *nslots, len = (n + FW_CHUNKS - 1) / FW_CHUNKS;
  const int lo = blockIdx.x * len, hi = min(n, lo + len);
  double s = 0.0;
  for (int a = lo; a < hi; ++a)
    {
      const int m = cnt[a];
      if (m) s += (double)m * rows[(size_t)a * stride + offset + b];
    }
  partial[(size_t)blockIdx.x * bins + b] = s;
}

__global__ void fw_finish_kernel(const double *partial, int bins, double *sums, double *weight, const Ctl *ctl)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *weight += (double)ctl->n_cur;        /* acc.weight() += 1 per walker (correlationEstimator.cpp:150) */
  if (b >= bins) return;
  double s = 0.0;
  for (int c = 0; c < FW_CHUNKS; ++c) s += partial[(size_t)c * bins + b];
  sums[b] += s;
}

} // namespace

size_t fw_record_group_bytes(const Plan &plan, int max_sf_bins)
{
  const int wpw = CTA_THREADS / 32;
  size_t d = (size_t)plan.dofs * (1 + wpw) + (wpw * 4 + 8) + plan.needs_pair_table + (size_t)max_sf_bins * 2 * wpw;
  d = (d + 1) & ~(size_t)1;
  return d * 8;
}

cudaError_t configure_kernels_fw(int max_smem_optin)
{
#ifdef QMC_EMULATE
  (void)max_smem_optin;
  return cudaSuccess;
#else
  return cudaFuncSetAttribute((const void *)fw_record_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
#endif
}

cudaError_t launch_fw_record(cudaStream_t st, int grid, size_t group_bytes, const Plan &plan, const WalkerBuf &buf, const int *n_desc,
                             const Ctl *ctl, double *rows, int *nslots_out)
{
  const size_t smem = group_bytes + (size_t)plan.hist_stride * 8;
  QMC_LAUNCH(fw_record_kernel, grid, CTA_THREADS, smem, st, plan, buf, n_desc, ctl, rows, nslots_out, group_bytes);
  return cudaGetLastError();
}

cudaError_t launch_fw_compose(cudaStream_t st, int grid, const int *parent, const int *comp_in, int *comp_out, const Ctl *ctl, int fresh)
{
  QMC_LAUNCH(fw_compose_kernel, grid, 256, 0, st, parent, comp_in, comp_out, ctl, fresh);
  return cudaGetLastError();
}

cudaError_t launch_fw_copy_links(cudaStream_t st, int grid, const int *comp, int *link, const Ctl *ctl)
{
  QMC_LAUNCH(fw_copy_links_kernel, grid, 256, 0, st, comp, link, ctl);
  return cudaGetLastError();
}

cudaError_t launch_fw_accumulate(cudaStream_t st, int grid, const int *parent, const int *links, int depth, int cap, long event, int fw,
                                 const Ctl *ctl, int *cnt, const double *rows, const int *nslots, int stride, int offset, int bins,
                                 double *partial, double *sums, double *weight)
{
  cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)cap * sizeof(int), st);
  if (e != cudaSuccess) return e;
  QMC_LAUNCH(fw_count_kernel, grid, 256, 0, st, parent, links, depth, cap, event, fw, ctl, cnt);
  QMC_LAUNCH(fw_colsum_kernel, dim3(FW_CHUNKS, (bins + 127) / 128), 128, 0, st, rows, cnt, nslots, stride, offset, bins, partial);
  QMC_LAUNCH(fw_finish_kernel, (bins + 127) / 128, 128, 0, st, partial, bins, sums, weight, ctl);
  return cudaGetLastError();
}

int fw_chunks() { return FW_CHUNKS; }

} // namespace qmc
