// kernel_common.cuh — helpers shared by the kernel translation units (kernels.cu, kernels_v2.cu):
// CTA geometry, walker-group construction, shared-memory carve-up, E_T formula.
#pragma once
#include "kernels.cuh"
#include "evaluate.cuh"

namespace qmc {
namespace {

constexpr int CTA_THREADS = 256;
#ifndef QMC_MIN_BLOCKS
#define QMC_MIN_BLOCKS 2
#endif

QMC_DEVICE Group make_group(int nwarps_per_group)
{
  Group G;
  const int warp = threadIdx.x >> 5;
  const int g = warp / nwarps_per_group;
  G.nwarps = nwarps_per_group;
  G.gwarp = warp - g * nwarps_per_group;
  G.lane = threadIdx.x & 31;
  G.gtid = G.gwarp * 32 + G.lane;
  G.gsize = nwarps_per_group * 32;
  G.bar_id = 1 + g;
  return G;
}

QMC_DEVICE Scratch make_scratch(const Plan &plan, unsigned char *smem, int group_index, int nwarps_per_group, size_t group_bytes)
{
  Scratch S;
  double *base = reinterpret_cast<double *>(smem + (size_t)group_index * group_bytes);
  S.R = base;
  S.acc = base + plan.dofs;
  S.red = S.acc + (size_t)nwarps_per_group * plan.dofs;
  S.tab = S.red + (nwarps_per_group * 4 + 8);
  S.aux = S.tab + plan.needs_pair_table;       /* needs_pair_table = doubles reserved for the pair table */
  return S;
}

QMC_DEVICE double branch_shift(const DmcParams &dmc, double sumE, double pop)
{ /* branchingControl::setEnergyShift (branching.cpp:65-97) */
  double e = sumE / pop;
  if (pop > dmc.mean_walkers + dmc.delta_walkers) e += dmc.log_up;
  else if (pop < dmc.mean_walkers - dmc.delta_walkers) e += dmc.log_down;
  return e;
}


} // namespace
} // namespace qmc
