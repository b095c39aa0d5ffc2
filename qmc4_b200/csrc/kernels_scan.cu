// kernels_scan.cu — branching bookkeeping of one DMC step, shared by both step pipelines:
//   nDescendants per walker (pipeline 2 decides it here: branchingControl::nDescendants, branching.cpp:11-14; pipeline 1's
//   propagate kernel already wrote it) -> exclusive prefix sum -> parent map of the next step (branchingControl::branch,
//   branching.cpp:16-63: the kill / clone loop becomes "slot j of the new population is a copy of slot parent[j]") ->
//   population, sum of the local energies of the NEW population, Metropolis counters -> the E_T inputs
//   (branchingControl::setEnergyShift, branching.cpp:65-97).
//
// Several CTAs (1024 walkers each) with a decoupled look-back: a CTA takes a ticket (its logical index, so every
// predecessor has started), publishes its total, sums the totals of its predecessors and fills its share of the parent map.
// The CTA that finishes last adds the per-CTA partial sums in index order (bitwise reproducible) and writes the control
// block.  Round 1 ran this on ONE CTA (34 us at 16k walkers, 10 % of a config-2 step).
#include "kernels.cuh"
#include "evaluate.cuh"
#include "kernel_common.cuh"

namespace qmc {

namespace {

constexpr int SCAN_T = 256;              // threads per CTA
constexpr int SCAN_ITEMS = 4;            // walkers per thread
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

template <bool DECIDE>
__global__ void __launch_bounds__(SCAN_T)
scan_multi_kernel(const __grid_constant__ DmcParams dmc, const StepArgs a, const double *e_old, int *parent_next, int n_partials, int multi_rank,
                  int init_mode)
{
  const ScanWorkspace ws = a.ws;
  __shared__ int s_warp[SCAN_T / 32];
  __shared__ double s_red[SCAN_T / 32];
  __shared__ int s_bid, s_prefix, s_tot, s_last;
  Ctl *ctl = a.ctl;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int W = ctl->n_cur + (DECIDE ? a.n_adjust : 0);     /* pipeline 2: migration removed the tail / appended the arrivals */
  const int nblk = (int)gridDim.x;
  const bool solo = nblk == 1;                               /* small populations: one CTA, no inter-CTA protocol at all */
  if (!solo)
    {
      if (t == 0) s_bid = atomicAdd(ws.ticket, 1);          /* logical index in arrival order: predecessors are running */
      __syncthreads();
    }
  const int bid = solo ? 0 : s_bid;
  const unsigned long long step = ctl->step;
  const uint32_t step_lo = (uint32_t)step, step_hi = (uint32_t)(step >> 32) | ((uint32_t)dmc.rank << 24);
  const double eT = (DECIDE && dmc.branching_on) ? branch_shift(dmc, ctl->gsum[0], ctl->gsum[1]) : 0.0;

  // ---- this CTA's walkers: tiles bid, bid + nblk, ... (the launch has enough CTAs for one tile each in the common case)
  double esum = 0.0;
  int nacc = 0, dropped = 0;
  // each CTA owns ONE contiguous range of walkers [lo, hi)
  const int per = (W + nblk - 1) / nblk;
  const int span = ((per + SCAN_TILE - 1) / SCAN_TILE) * SCAN_TILE;      /* multiple of the tile */
  const int lo = min(W, bid * span), hi = min(W, lo + span);
  // pass 1: total descendants of the range
  int mine_total = 0;
  for (int base = lo; base < hi; base += SCAN_TILE)
    {
      const int j0 = base + SCAN_ITEMS * t;
#pragma unroll
      for (int u = 0; u < SCAN_ITEMS; ++u)
        {
          const int j = j0 + u;
          if (j < hi)
            {
              int nn;
              if (DECIDE)
                {
                  nn = 1;
                  const double e = a.out.E[j];
                  if (dmc.branching_on || a.trace.u_b)
                    {
                      const u32x4 rd = philox4x32_10({(uint32_t)j, 0xffffffffu, step_lo, step_hi}, dmc.seed_lo, dmc.seed_hi);
                      const double u_b = u01_53(rd.z, rd.w);
                      if (dmc.branching_on) nn = (int)(exp(-dmc.tau * (0.5 * (e + e_old[j]) - eT)) + u_b);      /* branching.cpp:13 */
                      if (a.trace.u_b) a.trace.u_b[j] = u_b;
                    }
                  nn = max(0, min(nn, a.cap));
                  a.n_desc[j] = nn;
                  if (a.trace.n_desc) a.trace.n_desc[j] = nn;
                  nacc += a.flags[j] & 1;
                }
              else nn = a.n_desc[j];
              mine_total += nn;
            }
        }
    }
  int tot = warp_sum_i(mine_total);
  if (lane == 0) s_warp[warp] = tot;
  __syncthreads();
  if (warp == 0)
    {
      int s = 0;
      for (int w = 0; w < SCAN_T / 32; ++w) s += s_warp[w];
      if (lane == 0)
        { /* publish the total of this range in ONE word (bit 31 = valid) ... */
          s_tot = s;
          if (!solo) atomicExch(ws.flag + bid, 0x80000000u | (unsigned)s);
        }
      /* ... and look back: the lanes wait for different predecessors */
      int prefix = 0;
      for (int i = lane; i < bid; i += 32)
        {
          unsigned v;
          while (((v = atomicAdd(ws.flag + i, 0u)) & 0x80000000u) == 0u) { }
          prefix += (int)(v & 0x7fffffffu);
        }
      prefix = warp_sum_i(prefix);
      if (lane == 0) s_prefix = prefix;
    }
  __syncthreads();
  // pass 2: offsets inside the range and the parent map
  int running = s_prefix;
  for (int base = lo; base < hi; base += SCAN_TILE)
    {
      const int j0 = base + SCAN_ITEMS * t;
      int n[SCAN_ITEMS];
#pragma unroll
      for (int u = 0; u < SCAN_ITEMS; ++u) n[u] = (j0 + u < hi) ? a.n_desc[j0 + u] : 0;
      const int mine = n[0] + n[1] + n[2] + n[3];
      int inc = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
        {
          const int v = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += v;
        }
      __syncthreads();                               /* s_warp is reused */
      if (lane == 31) s_warp[warp] = inc;
      __syncthreads();
      int woff = 0, tile_total = 0;
      for (int w = 0; w < SCAN_T / 32; ++w)
        {
          const int v = s_warp[w];
          if (w < warp) woff += v;
          tile_total += v;
        }
      int off = running + woff + (inc - mine);
#pragma unroll
      for (int u = 0; u < SCAN_ITEMS; ++u)
        {
          const int j = j0 + u;
          if (n[u] > 0)
            {
              const double e = a.out.E[j];
              for (int c = 0; c < n[u]; ++c)
                {
                  if (off < a.cap)
                    {
                      parent_next[off] = j;
                      esum += e;
                    }
                  else dropped = 1;
                  ++off;
                }
            }
        }
      running += tile_total;
    }
  // ---- per-CTA partial sums in a fixed order: lanes -> warps -> thread 0
  double es = warp_sum(esum);
  int na = warp_sum_i(nacc);
  int dr = warp_sum_i(dropped);
  __syncthreads();
  if (lane == 0)
    {
      s_red[warp] = es;
      s_warp[warp] = na | (dr ? (1 << 30) : 0);
    }
  __syncthreads();
  if (t == 0)
    {
      double e = 0;
      int acc = 0, drp = 0;
      for (int w = 0; w < SCAN_T / 32; ++w)
        {
          e += s_red[w];
          acc += s_warp[w] & ((1 << 30) - 1);
          drp |= s_warp[w] >> 30;
        }
      if (drp) ctl->overflow = 1;
      if (solo)
        {
          s_red[0] = e;
          s_warp[0] = acc;
          s_last = 1;
        }
      else
        {
          ws.esum[bid] = e;
          ws.nacc[bid] = acc;
          __threadfence();
          s_last = (atomicAdd(ws.done, 1) == nblk - 1) ? 1 : 0;
        }
    }
  __syncthreads();
  if (!s_last) return;
  // ---- the CTA that finished last: totals in index order, control block, workspace reset for the next launch
  if (!solo) __threadfence();
  if (t == 0)
    {
      int total = 0, acc = 0;
      double e = 0.0;
      if (solo)
        {
          total = s_tot;
          e = s_red[0];
          acc = s_warp[0];
        }
      else
        for (int i = 0; i < nblk; ++i)
          {
            total += (int)(*((volatile unsigned *)(ws.flag + i)) & 0x7fffffffu);
            e += *((volatile double *)(ws.esum + i));
            acc += *((volatile int *)(ws.nacc + i));
          }
      if (total == 0 && W > 0)
        { /* always keep at least a walker (branching.cpp:29-32) */
          a.n_desc[0] = 1;
          parent_next[0] = 0;
          e = a.out.E[0];
          total = 1;
        }
      const int newW = min(total, a.cap);
      double moves = (double)W, accd = (double)acc;
      if (!DECIDE)
        { /* pipeline 1: the propagate kernel counted per walker group */
          moves = 0.0;
          accd = 0.0;
          for (int i = 0; i < n_partials; ++i)
            {
              moves += a.partials[2 * i];
              accd += a.partials[2 * i + 1];
            }
        }
      ctl->n_prev = W;
      ctl->n_cur = newW;
      if (!init_mode)
        {
          ctl->step += 1;
          ctl->blk_moves += moves;
          ctl->blk_acc += accd;
          ctl->blk_steps += 1;
        }
      ctl->lsum[0] = e;
      ctl->lsum[1] = (double)newW;
      for (int r = 0; r < MAX_RANKS; ++r)
        {
          ctl->lsum[2 + r] = (r == dmc.rank) ? (double)newW : 0.0;
          ctl->lsum[2 + MAX_RANKS + r] = (r == dmc.rank) ? (double)(a.cap - max(W, newW)) : 0.0;
        }
      if (a.accumulate)
        {
          ctl->blk_esum += e;
          ctl->blk_ecount += (double)newW;
          ctl->est_weight += (double)newW;
        }
      if (!multi_rank)
        {
          for (int r = 0; r < NSUM; ++r) ctl->gsum[r] = ctl->lsum[r];
          if (dmc.branching_on) ctl->eshift = branch_shift(dmc, e, (double)newW);
        }
      if (!solo)
        {
          *ws.ticket = 0;
          *ws.done = 0;
        }
    }
  if (!solo)
    {
      __syncthreads();      /* thread 0 has summed the published totals: only now may the other threads clear them for the next launch */
      for (int i = t; i < nblk; i += SCAN_T) ws.flag[i] = 0u;
    }
}

} // namespace

int scan_max_blocks() { return 128; }

cudaError_t launch_scan_multi(cudaStream_t st, bool decide, const DmcParams &dmc, const StepArgs &a, const double *e_old, int n_partials,
                              int multi_rank, int init_mode)
{
  const int64_t capacity_hint = a.cap;
  /* one tile of 1024 walkers per CTA up to scan_max_blocks() CTAs (all co-resident: 256 threads, no shared-memory pressure) */
  int64_t g = (capacity_hint + SCAN_TILE - 1) / SCAN_TILE;
  if (g < 1) g = 1;
  if (g > scan_max_blocks()) g = scan_max_blocks();
  if (decide)
    QMC_LAUNCH((scan_multi_kernel<true>), (int)g, SCAN_T, 0, st, dmc, a, e_old, a.parent, n_partials, multi_rank, init_mode);
  else
    QMC_LAUNCH((scan_multi_kernel<false>), (int)g, SCAN_T, 0, st, dmc, a, e_old, a.parent, n_partials, multi_rank, init_mode);
  return cudaGetLastError();
}

cudaError_t launch_scan(cudaStream_t st, const DmcParams &dmc, const StepArgs &a, int n_partials, int multi_rank, int init_mode)
{
  return launch_scan_multi(st, false, dmc, a, nullptr, n_partials, multi_rank, init_mode);
}
cudaError_t launch_scan_v2(cudaStream_t st, const DmcParams &dmc, const StepArgs &a, const double *e_old, int multi_rank)
{
  return launch_scan_multi(st, true, dmc, a, e_old, 0, multi_rank, 0);
}

} // namespace qmc
