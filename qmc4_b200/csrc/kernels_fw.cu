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
__global__ void fw_compose_kernel(const int *parent, const int *comp_in, int *comp_out, const Ctl *ctl, int fresh, int *level_slots)
{
  const int n = ctl->n_cur;
  if (level_slots && blockIdx.x == 0 && threadIdx.x == 0) *level_slots = ctl->n_prev;   /* slots this event's records occupy */
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    comp_out[i] = fresh ? parent[i] : comp_in[parent[i]];
}

__global__ void fw_chain_kernel(const int *comp, const int *links, int depth, int cap, long e_last, int levels, int first, int count, int *chain)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  int a = comp[first + k];
  for (int l = 0; l < levels; ++l)
    {
      chain[(size_t)k * levels + l] = a;
      if (l + 1 < levels) a = links[(size_t)((e_last - l) % depth) * cap + a];
    }
}

__global__ void fw_records_kernel(int unpack, double *rows, int row_depth, int row_levels, int stride, double *rec, int width, int *links,
                                  int depth, int levels, int cap, long e_last, const int *chain, int base, int *comp, int first, int count,
                                  double *msg)
{
  const int k = blockIdx.x;
  if (k >= count) return;
  const size_t msz = (size_t)row_levels * stride + (size_t)(rec ? levels : 0) * width;
  double *m = msg + (size_t)k * msz;
  for (int l = 0; l < row_levels; ++l)
    {
      const int slot = unpack ? base + k : chain[(size_t)k * levels + l];
      double *row = rows + ((size_t)((e_last - l) % row_depth) * cap + slot) * stride;
      for (int b = threadIdx.x; b < stride; b += blockDim.x)
        {
          if (unpack) row[b] = m[(size_t)l * stride + b];
          else m[(size_t)l * stride + b] = row[b];
        }
    }
  if (rec)
    for (int t = threadIdx.x; t < levels * width; t += blockDim.x)
      {
        const int l = t / width, w = t % width;
        const int slot = unpack ? base + k : chain[(size_t)k * levels + l];
        double *r = rec + ((size_t)((e_last - l) % depth) * cap + slot) * width + w;
        double *q = m + (size_t)row_levels * stride + t;
        if (unpack) *r = *q;
        else *q = *r;
      }
  if (unpack)
    {
      for (int l = threadIdx.x; l < levels; l += blockDim.x) links[(size_t)((e_last - l) % depth) * cap + base + k] = base + k;
      if (threadIdx.x == 0) comp[first + k] = base + k;
    }
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


// ---- superfluidFraction (correlationEstimator.cpp:187-345) -------------------------------------------------------
// Reference: centerOfMassStorer keeps, per walker, a ring of `recordSteps` records {sum_i x_i^d of every stored set}
// (unwrapped coordinates, :222-246), deep-copied with the walker; superfluidFractionEstimator::accumulate (:322-357)
// adds, for the record j written last and every record i <= j of the SAME turn of the ring,
// W = sum_d (XA_i - XA_j)(XB_i - XB_j) to bin j-i with weight 1.
// Here: one record per surviving slot and sampling event (3*nSpecies doubles) in the same event ring as the ancestor
// links; a slot with n copies in the post-branching population follows its j links once and contributes n*W.

constexpr int SFL_THREADS = 128, SFL_BATCH = 8;

// rec[slot][s*3+d] = sum_i R[slot][off_s + d*N_s + i], summed in particle order like centerOfMassStorer::computeCenterOfMass
__global__ void sfl_record_kernel(const __grid_constant__ Plan plan, const WalkerBuf buf, const int *n_desc, const Ctl *ctl, double *rec)
{
  const int n = ctl->n_prev, width = plan.nSpecies * 3;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)n * width; t += (long)gridDim.x * blockDim.x)
    {
      const int j = (int)(t / width), k = (int)(t % width), s = k / 3, d = k % 3;
      if (n_desc[j] <= 0) continue;
      const double *x = buf.R + (size_t)j * plan.dofs + plan.off[s] + d * plan.N[s];
      double acc = 0.0;
      for (int i = 0; i < plan.N[s]; ++i) acc += x[i];
      rec[(size_t)j * width + k] = acc;
    }
}

// partial[chunk][lag] = sum over the chunk's slots (fixed tree order) of n_desc[slot] * W(lag), lag = 0..jmax
__global__ void __launch_bounds__(SFL_THREADS)
sfl_lag_kernel(const double *rec, const int *links, const int *n_desc, const Ctl *ctl, int depth, int cap, int width, long event, int jmax,
              int offA, int offB, int lag_stride, double *partial)
{
  __shared__ double sh[SFL_BATCH][SFL_THREADS];
  const int n = ctl->n_prev;
  const int slot = blockIdx.x * SFL_THREADS + threadIdx.x;
  if (blockIdx.x * SFL_THREADS >= n) return;                       /* uniform per CTA; the finish kernel skips these chunks too */
  const int m = slot < n ? max(n_desc[slot], 0) : 0;
  double a0[3] = {0, 0, 0}, b0[3] = {0, 0, 0};
  if (m)
    {
      const double *r0 = rec + ((size_t)(event % depth) * cap + slot) * width;
      for (int d = 0; d < 3; ++d) { a0[d] = r0[offA + d]; b0[d] = r0[offB + d]; }
    }
  int a = slot;
  for (int base = 0; base <= jmax; base += SFL_BATCH)
    {
      const int nb = min(SFL_BATCH, jmax + 1 - base);
      for (int q = 0; q < nb; ++q)
        {
          double W = 0.0;
          if (m)
            {
              const long ev = event - (base + q);
              const double *r = rec + ((size_t)(ev % depth) * cap + a) * width;
              for (int d = 0; d < 3; ++d) W += (r[offA + d] - a0[d]) * (r[offB + d] - b0[d]);      /* :343-350 */
              W *= (double)m;
              if (base + q < jmax) a = links[(size_t)(ev % depth) * cap + a];   /* ancestor slot one event earlier */
            }
          sh[q][threadIdx.x] = W;
        }
      __syncthreads();
      for (int half = SFL_THREADS / 2; half > 0; half >>= 1)
        {
          if ((int)threadIdx.x < half)
            for (int q = 0; q < nb; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + half];
          __syncthreads();
        }
      if ((int)threadIdx.x < nb) partial[(size_t)blockIdx.x * lag_stride + base + threadIdx.x] = sh[threadIdx.x][0];
      __syncthreads();
    }
}

__global__ void sfl_finish_kernel(const double *partial, int lag_stride, int jmax, double *sums, double *weights, const Ctl *ctl)
{
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > jmax) return;
  const int chunks = (ctl->n_prev + SFL_THREADS - 1) / SFL_THREADS;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * lag_stride + l];
  sums[l] += s;
  weights[l] += (double)ctl->n_cur;                               /* acc.accumulate(W, j-i): weight 1 per walker (:352) */
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

cudaError_t launch_fw_compose(cudaStream_t st, int grid, const int *parent, const int *comp_in, int *comp_out, const Ctl *ctl, int fresh,
                              int *level_slots)
{
  QMC_LAUNCH(fw_compose_kernel, grid, 256, 0, st, parent, comp_in, comp_out, ctl, fresh, level_slots);
  return cudaGetLastError();
}

// ---- migration of the genealogical records with their walker (qmc_rebalance; the reference ships the walker's whole
// ring, walkers.cpp:181-233 pack/unpack + ptools.cpp:271-401) ---------------------------------------------------------
// sender: chain[k][l] = slot, at event e_last - l, of the ancestor of the k-th tail walker (population index first + k)
cudaError_t launch_fw_chain(cudaStream_t st, const int *comp, const int *links, int depth, int cap, long e_last, int levels, int first, int count,
                            int *chain)
{
  QMC_LAUNCH(fw_chain_kernel, (count + 127) / 128, 128, 0, st, comp, links, depth, cap, e_last, levels, first, count, chain);
  return cudaGetLastError();
}
// sender: msg[k] = { rows of the last row_levels events, centre-of-mass records of the last `levels` events } along the chain;
// receiver (unpack = 1): the same message written to slot base + k of every level, links[level][base + k] = base + k,
// comp[first + k] = base + k
cudaError_t launch_fw_records(cudaStream_t st, int unpack, double *rows, int row_depth, int row_levels, int stride, double *rec, int width,
                              int *links, int depth, int levels, int cap, long e_last, const int *chain, int base, int *comp, int first,
                              int count, double *msg)
{
  QMC_LAUNCH(fw_records_kernel, count, 128, 0, st, unpack, rows, row_depth, row_levels, stride, rec, width, links, depth, levels, cap, e_last,
             chain, base, comp, first, count, msg);
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

int sfl_chunk_slots() { return SFL_THREADS; }

cudaError_t launch_sfl_record(cudaStream_t st, int grid, const Plan &plan, const WalkerBuf &buf, const int *n_desc, const Ctl *ctl, double *rec)
{
  QMC_LAUNCH(sfl_record_kernel, grid, 256, 0, st, plan, buf, n_desc, ctl, rec);
  return cudaGetLastError();
}

// links of event e-l are followed only when lag l+1 is needed, so the unwritten links of event 0 are never read
cudaError_t launch_sfl_accumulate(cudaStream_t st, const double *rec, const int *links, const int *n_desc, const Ctl *ctl, int depth, int cap,
                                 int width, long event, int jmax, int offA, int offB, int lag_stride, double *partial, double *sums,
                                 double *weights)
{
  const int chunks = (cap + SFL_THREADS - 1) / SFL_THREADS;
  QMC_LAUNCH(sfl_lag_kernel, chunks, SFL_THREADS, 0, st, rec, links, n_desc, ctl, depth, cap, width, event, jmax, offA, offB, lag_stride, partial);
  QMC_LAUNCH(sfl_finish_kernel, (jmax + 128) / 128, 128, 0, st, partial, lag_stride, jmax, sums, weights, ctl);
  return cudaGetLastError();
}

int fw_chunks() { return FW_CHUNKS; }

} // namespace qmc
